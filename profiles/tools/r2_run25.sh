set -x
mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"k_miller_part|k_final_s" -c 15 -o /tmp/ncu/split python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > /dev/null 2>&1
ncu -i /tmp/ncu/split.ncu-rep --page raw --csv > /tmp/ncu/split_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("/tmp/ncu/split_raw.csv")))
h=rows[0]
keep=[i for i,c in enumerate(h) if c in ("Kernel Name","gpu__time_duration.sum","inst_executed","launch__registers_per_thread","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","sass__inst_executed_local_loads","sass__inst_executed_local_stores","l1tex__t_sector_hit_rate.pct") or "pcsamp_warps_issue_stalled" in c]
w=csv.writer(open("gpurun_out/r2_split_raw_excerpt.csv","w"))
for r in rows: w.writerow([r[i][:60] for i in keep])
PY
python profiles/tools/ncu_executed.py 65536 /tmp/ncu/split.ncu-rep > gpurun_out/r2_split_exec.json 2> gpurun_out/r2_split_exec.err
