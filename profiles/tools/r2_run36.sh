set -x
BLSGPU_SO=build_var/pow2.so timeout 900 python -m pytest tests -m gpu -x -q -k "witness or r1cs" 2>&1 | tail -2
for v in base pow2; do echo "== $v"; export BLSGPU_SO=build_var/$v.so; timeout 300 python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid|matches|rror" | tail -4
  python bench_configs.py --cfg 5r --steps 2 --scale 0.5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms'))"
done
