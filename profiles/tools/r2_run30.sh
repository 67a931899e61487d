set -x
for v in default v_minb v_it16 v_it4; do
  echo "== $v"; if [ $v = default ]; then unset BLSGPU_SO; else export BLSGPU_SO=build_var/$v.so; fi
  python bench.py --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done
unset BLSGPU_SO
for ln in 1 2 3 4; do echo "== n=131072 lanes=$ln"; python bench.py --n 131072 --lanes $ln --skip-extra --no-cpu --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"; done
