set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "verify or split or cooperative or gt_bytes or api or fast_aggregate or committees" 2>&1 | tail -2
python profiles/tools/latency_stages.py 2>&1 | tail -3
python profiles/tools/latency.py 2>&1 | grep -E "'n': (1|1024|8192|37888)," 
