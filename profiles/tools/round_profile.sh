#!/bin/bash
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err || exit 1
python bench_configs.py --steps 2 > gpurun_out/bench_configs_r01c.jsonl 2> gpurun_out/bench_configs_r01c.err
python bench.py --n 65536 --steps 1 --warmup 3 --no-cpu --lanes 1 > gpurun_out/plain_r01c.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --n 65536 --steps 1 --warmup 3 --no-cpu --lanes 1 > gpurun_out/ncu_l.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_miller$|k_final_exp" -c 2 --launch-skip 6 -o gpurun_out/prof_r01c -f python bench.py --n 65536 --steps 1 --warmup 3 --no-cpu --lanes 1 > gpurun_out/ncu_f.log 2>&1
cat gpurun_out/bench_r01c.json | cut -c1-300
