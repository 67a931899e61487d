set -x
python -m pytest tests -x -q -m gpu -k "witness" 2>&1 | tail -6
python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "split|grid|rror|matches" | tail -6
python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "grid|rror|matches" | tail -2
BLSGPU_SO=build_var/wit_trace.so python profiles/tools/wit_trace.py 512 2>&1 | tail -16
