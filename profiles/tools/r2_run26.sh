set -x
for v in default minb3 minb4; do
  echo "== $v"; if [ $v = default ]; then unset BLSGPU_SO; else export BLSGPU_SO=build_var/$v.so; fi
  python bench.py --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
