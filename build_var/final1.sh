python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash profiles/tools/round_profile.sh > gpurun_out/round_profile.log 2>&1
python bench_r1cs.py --steps 3 --warmup 3 > gpurun_out/bench_r1cs_1gpu.json 2> gpurun_out/bench_r1cs_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-400 gpurun_out/bench_r1cs_1gpu.json; cut -c1-300 gpurun_out/bench_ref.json
