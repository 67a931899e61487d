set -x
python -m pytest tests -x -q -m gpu -k "aggregate_verify or uncompressed" 2>&1 | tail -8
mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"^k_decode_g1|^k_decode_g2|^k_hash_to_g2|^k_miller$|^k_final_exp" -c 6 -o /tmp/ncu/r2_stage_a python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > gpurun_out/r2_ncu_stage.log 2>&1; tail -3 gpurun_out/r2_ncu_stage.log
python profiles/tools/ncu_executed.py 65536 /tmp/ncu/r2_stage_a.ncu-rep > gpurun_out/ncu_r02_executed.json; grep -c imad_wide gpurun_out/ncu_r02_executed.json
ncu -i /tmp/ncu/r2_stage_a.ncu-rep --page raw --csv > gpurun_out/ncu_r02_stage_raw.csv 2>/dev/null
ncu -i /tmp/ncu/r2_stage_a.ncu-rep --page source --csv -k k_miller 2>/dev/null | cut -d, -f1-10 > gpurun_out/ncu_r02_k_miller_source.csv; wc -l gpurun_out/ncu_r02_k_miller_source.csv
