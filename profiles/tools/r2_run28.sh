set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p /tmp/ncu
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_lines.csv python bench.py --n 262144 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_miller_lines|k_miller_accum" -s 2 -c 2 -o /tmp/ncu/lines python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > /dev/null 2>&1
ncu -i /tmp/ncu/lines.ncu-rep --page raw --csv > /tmp/ncu/lines_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("/tmp/ncu/lines_raw.csv")))
h=rows[0]
keep=[i for i,c in enumerate(h) if c in ("Kernel Name","gpu__time_duration.sum","inst_executed","launch__registers_per_thread","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","sass__inst_executed_local_loads","sass__inst_executed_local_stores","l1tex__t_sector_hit_rate.pct") or "pcsamp_warps_issue_stalled" in c]
w=csv.writer(open("gpurun_out/r2_lines_raw_excerpt.csv","w"))
for r in rows: w.writerow([r[i][:60] for i in keep])
PY
python profiles/tools/ncu_executed.py 65536 /tmp/ncu/lines.ncu-rep > gpurun_out/r2_lines_exec.json 2> gpurun_out/r2_lines_exec.err
