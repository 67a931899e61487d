# confirmation after the small-pass default (six-lane final exponentiation below 4,096 items): all GPU tests, smoke, bench line, latency table
set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-200 gpurun_out/bench_r02.json
