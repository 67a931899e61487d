set -x
python -m pytest tests -x -q -m gpu -k "witness" 2>&1 | tail -3
python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "split|grid|rror|matches" | tail -6
