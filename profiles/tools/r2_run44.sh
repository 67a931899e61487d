set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1
python profiles/tools/latency.py 2>&1 | grep -E "'n': (1|32|1024|8192)," | grep "True, 'six_lane_small_pass_kernels': True"
