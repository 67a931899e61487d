set -x
python -m pytest tests -x -q -m gpu -k "cooperative or uncompressed or aggregate_verify" 2>&1 | tail -4
python bench.py --skip-extra --no-cpu --steps 3 --coop 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('coop0', d['value'], d['stage_ms'])"
python bench.py --skip-extra --no-cpu --steps 3 --coop 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('coop1', d['value'], d['stage_ms'])"
