for L in 1 2 3 4; do python bench.py --n 131072 --steps 5 --warmup 3 --no-cpu --lanes $L 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes', d['config']['lanes_per_gpu'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'stage sum', round(sum(d['stage_ms'].values()),2), {k: round(v,1) for k,v in d['stage_ms'].items()})
"; done
