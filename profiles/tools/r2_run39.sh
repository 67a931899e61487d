set -x
BLSGPU_SO=build_var/wit_trace.so timeout 400 python profiles/tools/wit_trace.py 512 2>&1 | tail -40
for v in wit_nobar wit_nowork; do echo "== $v"; BLSGPU_SO=build_var/$v.so timeout 300 python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen" | tail -2; done
