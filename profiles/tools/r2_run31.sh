set -x
for v in longrows; do BLSGPU_SO=build_var/$v.so python -m pytest tests -m gpu -x -q -k "r1cs or witness" 2>&1 | tail -3; done
for v in default longrows longrows4; do
  echo "== $v"; if [ $v = default ]; then unset BLSGPU_SO; else export BLSGPU_SO=build_var/$v.so; fi
  python bench_configs.py --cfg 5r --steps 2 --scale 0.5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms'), d.get('config','')[:80])"
done
