set -x
python -m pytest tests -x -q -m gpu -k "witness" 2>&1 | tail -3
BLSGPU_SO=build_var/wit_pf8.so python profiles/tools/wit_trace.py 512 2>&1 | tail -24
for v in pf4 pf8 pf16; do echo "== wit_$v"; BLSGPU_SO=build_var/wit_$v.so python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid" | tail -3; done
BLSGPU_SO=build_var/wit_pf8.so python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "grid" | tail -1
