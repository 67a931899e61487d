python tests/devcheck/run_devcheck.py 2>&1 | grep -v " ok$" | tail -5
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for v in "" build_var/lib_pf.so; do
echo "== $v"
BLSGPU_SO=${v:+$PWD/$v} timeout 400 python bench_configs.py --cfg 5,5r --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms','witnesses_per_sec','assignments_per_sec')})
    elif 'rror' in l: print(l.strip()[:300])
"
done
BLSGPU_SO=$PWD/build_var/lib_pf.so python -m pytest tests -x -q -m gpu -k "r1cs" 2>&1 | tail -3
