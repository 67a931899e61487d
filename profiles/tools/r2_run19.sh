set -x
echo "== default (narrow)"; python bench_configs.py --cfg 5r --steps 3 --scale 0.5 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms'], d['value'])"
for v in wide wide4; do echo "== r1_$v"; BLSGPU_SO=build_var/r1_$v.so python bench_configs.py --cfg 5r --steps 3 --scale 0.5 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms'], d['value'])"; done
BLSGPU_SO=build_var/r1_wide.so python -m pytest tests -x -q -m gpu -k "r1cs" 2>&1 | tail -2
