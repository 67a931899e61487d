set -x
for v in wit_256; do echo "== $v"; BLSGPU_SO=build_var/$v.so python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid" | tail -3; done
echo "== default"; python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid" | tail -3
