# round-2 final measurements of the verify path on one GPU after the stage kernels were split (the commands behind profiles/*_r02*)
set -x
# (pytest -m gpu and smoke: run separately, 43 passed)

mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"^k_decode_g1|^k_decode_g2|^k_hash_field|^k_hash_map|^k_hash_clear|^k_miller_lines|^k_miller_accum|^k_final_squarings|^k_final_step" -c 32 -o /tmp/ncu/stage python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > /dev/null 2>&1
python profiles/tools/ncu_executed.py --stages 65536 /tmp/ncu/stage.ncu-rep > gpurun_out/ncu_r02_executed.json 2> gpurun_out/ncu_r02_executed.err; grep -c imad_wide gpurun_out/ncu_r02_executed.json
cp gpurun_out/ncu_r02_executed.json profiles/ncu_r02_executed.json
ncu -i /tmp/ncu/stage.ncu-rep --page raw --csv > /tmp/ncu/stage_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("/tmp/ncu/stage_raw.csv")))
h=rows[0]
want=("Kernel Name","dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum","inst_executed","l1tex__t_sector_hit_rate.pct","launch__block_size","launch__grid_size","launch__registers_per_thread","lts__t_sector_hit_rate.pct","lts__throughput.avg.pct_of_peak_sustained_elapsed","sass__inst_executed_local_loads","sass__inst_executed_local_stores","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","thread_inst_executed")
keep=[i for i,c in enumerate(h) if c in want or ("pcsamp_warps_issue_stalled" in c and "not_issued" not in c)]
w=csv.writer(open("gpurun_out/ncu_r02_stage_raw_excerpt.csv","w"))
for r in rows: w.writerow([r[i][:60] for i in keep])
PY
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python bench.py --split 0 --skip-extra --no-cpu > gpurun_out/bench_r02_one_launch.json 2>/dev/null; cut -c1-200 gpurun_out/bench_r02_one_launch.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --skip-extra --no-cpu > gpurun_out/launches_r02.log 2>&1; tail -1 gpurun_out/launches_r02.log | cut -c1-120
python bench_configs.py --cfg 2r,3a,3b,4 --steps 2 > gpurun_out/bench_configs_r02.jsonl 2>/dev/null; cut -c1-200 gpurun_out/bench_configs_r02.jsonl
ls -la gpurun_out | head -30
