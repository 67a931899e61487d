python -m pytest tests -x -q -m gpu -k "witness or r1cs" 2>&1 | tail -2
python bench_r1cs.py --steps 3 --warmup 3 > gpurun_out/bench_r1cs_1gpu.json 2> gpurun_out/bench_r1cs_1gpu.err; tail -c 1500 gpurun_out/bench_r1cs_1gpu.json; tail -3 gpurun_out/bench_r1cs_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_r1cs -c 60 --csv --log-file gpurun_out/l5r.csv python bench_configs.py --cfg 5r --steps 1 --scale 0.25 > gpurun_out/l5r.log 2>&1
