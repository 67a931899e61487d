set -x
for v in wit_tree wit_tree32; do echo "== $v"; BLSGPU_SO=build_var/$v.so timeout 240 python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid|matches|Error|error" | tail -5; done
echo "== default"; timeout 240 python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid|matches" | tail -4
