python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash profiles/tools/round_profile.sh > gpurun_out/round_profile.log 2>&1
python bench_r1cs.py --steps 3 --warmup 3 > gpurun_out/bench_r1cs_1gpu.json 2> gpurun_out/bench_r1cs_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_r1cs -c 60 --csv --log-file gpurun_out/l5r.csv python bench_configs.py --cfg 5r --steps 1 --scale 0.25 > gpurun_out/l5r.log 2>&1
python profiles/tools/wit_bench.py 512 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
cut -c1-200 gpurun_out/bench_r01c.json; cut -c1-200 gpurun_out/bench_r1cs_1gpu.json
