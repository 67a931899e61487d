python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash profiles/tools/cmp.sh 1048576 $PWD/build_var/lib_r1m6.so $PWD/bls_verify_gadget_b200/libblsgpu.so
python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
