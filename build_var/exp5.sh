for v in build_var/lib_l16s32.so build_var/lib_l16s16.so build_var/lib_l8s16.so build_var/lib_l12s32.so build_var/lib_l8s32.so; do
echo "== $v"
BLSGPU_SO=${v:+$PWD/$v} python -m pytest tests -x -q -m gpu -k "r1cs" 2>&1 | tail -1
BLSGPU_SO=${v:+$PWD/$v} timeout 400 python bench_configs.py --cfg 5r --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms','witnesses_per_sec','assignments_per_sec')})
    elif 'rror' in l: print(l.strip()[:300])
"
done
