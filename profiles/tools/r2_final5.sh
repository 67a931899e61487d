# last confirmation of round 2 on one GPU: all GPU tests, smoke, the driver's bench line, the secondary configs on their own
set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python bench_configs.py --cfg 2r,3a,3b,4 --steps 2 > gpurun_out/bench_configs_r02_verify.jsonl 2>/dev/null; cut -c1-160 gpurun_out/bench_configs_r02_verify.jsonl
