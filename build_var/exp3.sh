python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in build_var/lib_r1m6.so ""; do
  echo "== r1cs $v"
  BLSGPU_SO=${v:+$PWD/$v} timeout 400 python bench_configs.py --cfg 5,5r --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms','witnesses_per_sec','assignments_per_sec')})
    elif 'rror' in l: print(l.strip()[:300])
"
done
python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
