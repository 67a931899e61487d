set -x
python -m pytest tests -m gpu -x -q -k "verify or pairing or gt" 2>&1 | tail -3
for n in 131072; do for sp in 2 1; do for ln in 2 4; do
  echo "== n=$n split=$sp lanes=$ln"; python bench.py --n $n --split $sp --lanes $ln --skip-extra --no-cpu --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"
done; done; done
for sp in 2 1; do echo "== n=2^20 split=$sp"; python bench.py --split $sp --skip-extra --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['stage_ms'], d['gpu_launches'])"; done
