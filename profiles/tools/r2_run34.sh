set -x
BLSGPU_SO=build_var/asmem.so python -m pytest tests -m gpu -x -q -k "split or gt_bytes or verify_batch" 2>&1 | tail -2
for v in base asmem; do
  echo "== $v"; export BLSGPU_SO=build_var/$v.so
  python bench.py --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
