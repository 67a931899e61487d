set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "hash or api or sign" 2>&1 | tail -2
python bench_configs.py --cfg 4 --steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('value'), d.get('ms'))"
