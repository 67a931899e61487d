set -x
for ln in 1 2 3 4; do echo "== n=131072 lanes=$ln"; python bench.py --n 131072 --lanes $ln --skip-extra --no-cpu --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"; done
for ln in 1 4; do echo "== n=2^20 lanes=$ln"; python bench.py --lanes $ln --skip-extra --no-cpu --steps 3 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])"; done
