set -x
python -m pytest tests -x -q -m gpu -k "r1cs or witness or file or device_primitives" 2>&1 | tail -4
python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "witness_|rror|matches" | tail -9
python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "witness_|rror|matches" | tail -5
python bench_r1cs.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('r1cs', d['value'], d['ms_per_step'], 'e2e', d['e2e']['assignments_per_sec'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_r1cs -c 80 --csv --log-file gpurun_out/r2_l5r_d.csv python bench_configs.py --cfg 5r --steps 1 --scale 0.25 > gpurun_out/r2_l5r_d.log 2>&1; tail -1 gpurun_out/r2_l5r_d.log | cut -c1-200
