set -x
BLSGPU_SO=build_var/hsplit.so python -m pytest tests -m gpu -x -q -k "verify or pairing or gt or hash" 2>&1 | tail -3
for v in base fstep hsplit hsplit_b; do
  echo "== $v"; export BLSGPU_SO=build_var/$v.so
  python bench.py --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
