set -x
python -m pytest tests -x -q -m gpu -k "device_primitives or gt_bytes or verify_fixtures or verify_differential or committees" 2>&1 | tail -4
python bench.py --skip-extra --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('pingpong', d['value'], d['stage_ms'])"
BLSGPU_SO=build_var/nopp.so python bench.py --skip-extra --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('rolled', d['value'], d['stage_ms'])"
