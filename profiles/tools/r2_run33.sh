set -x
BLSGPU_SO=build_var/pack_l8.so python -m pytest tests -m gpu -x -q -k "r1cs or witness" 2>&1 | tail -3
for v in base l8 l4 l2 pack_l16 pack_l8 pack_l4; do
  echo "== $v"; export BLSGPU_SO=build_var/$v.so
  python bench_configs.py --cfg 5r --steps 2 --scale 0.5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms'))"
done
