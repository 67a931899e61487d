set -x
python -m pytest tests -x -q -m gpu -k "any_message_length or device_primitives" 2>&1 | tail -4
BLSGPU_SO=build_var/wit_trace.so python profiles/tools/wit_trace.py 512 2>&1 | tail -24
for v in a b c d e f; do echo "== r1_$v"; BLSGPU_SO=build_var/r1_$v.so python bench_configs.py --cfg 5r --steps 3 --scale 0.5 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(d['ms'], d['value'])"; done
