set -x
BLSGPU_SO=build_var/wsort.so timeout 600 python -m pytest tests -m gpu -x -q -k "witness" 2>&1 | tail -2
for v in base wsort wsort_b3 wsort_b4; do echo "== $v"; BLSGPU_SO=build_var/$v.so timeout 300 python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid|matches|rror" | tail -4; done
