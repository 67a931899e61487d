set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1
( time python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err ) 2>&1 | tail -4; tail -5 gpurun_out/r2_bench_a.err; cut -c1-1500 gpurun_out/r2_bench_a.json
ncu --set full --clock-control none --import-source on -k regex:"^k_decode_g1|^k_decode_g2|^k_hash_to_g2|^k_miller$|^k_final_exp" -c 14 -o gpurun_out/r2_stage_a python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > gpurun_out/r2_ncu_stage.log 2>&1; tail -3 gpurun_out/r2_ncu_stage.log
