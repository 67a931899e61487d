set -x
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2>/dev/null; cut -c1-200 gpurun_out/bench_reference_arm_r02.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --skip-extra --no-cpu > gpurun_out/launches_r02.log 2>&1; tail -1 gpurun_out/launches_r02.log | cut -c1-120
