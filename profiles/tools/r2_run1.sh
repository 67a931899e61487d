set -x
python -m pytest tests -x -q -m gpu -k "r1cs or witness" 2>&1 | tail -5
python bench_r1cs.py --steps 3 --warmup 3 > gpurun_out/r2_bench_r1cs_a.json 2> gpurun_out/r2_bench_r1cs_a.err; tail -3 gpurun_out/r2_bench_r1cs_a.err; cut -c1-600 gpurun_out/r2_bench_r1cs_a.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_r1cs -c 80 --csv --log-file gpurun_out/r2_l5r_a.csv python bench_configs.py --cfg 5r --steps 1 --scale 0.25 > gpurun_out/r2_l5r_a.log 2>&1; tail -2 gpurun_out/r2_l5r_a.log
python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
