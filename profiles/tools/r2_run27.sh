set -x
BLSGPU_SO=build_var/lines2.so python -m pytest tests -m gpu -x -q -k "verify or pairing or gt" 2>&1 | tail -3
for v in default lines2 lines3; do for n in 1048576 131072; do
  echo "== $v n=$n"; if [ $v = default ]; then unset BLSGPU_SO; else export BLSGPU_SO=build_var/$v.so; fi
  python bench.py --n $n --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done; done
