set -x
for v in wit_nobar wit_nowork; do echo "== $v"; BLSGPU_SO=build_var/$v.so timeout 300 python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen.*grid" | tail -2; done
echo "== default"; timeout 300 python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen.*grid" | tail -2
