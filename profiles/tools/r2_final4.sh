# final one-GPU run of round 2: all GPU tests, smoke, the driver's bench line, the reference arm, R1CS launch list + ncu excerpt, secondary configs, witness timing
set -x
( time python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -2 gpurun_out/bench_r02.err; cut -c1-300 gpurun_out/bench_r02.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2>/dev/null; cut -c1-200 gpurun_out/bench_reference_arm_r02.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_r1cs|k_witness" -c 200 --csv --log-file gpurun_out/launches_r1cs_r02.csv python bench_r1cs.py --steps 1 --warmup 1 --per-gpu 128 --no-cpu > gpurun_out/launches_r1cs_r02.log 2>&1; tail -1 gpurun_out/launches_r1cs_r02.log | cut -c1-120
mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"k_r1cs_long_rows|k_r1cs_transpose|k_r1cs_lut$|k_r1cs_rows_list|k_witness_light|k_witness_levels" -c 8 -o /tmp/ncu/r1 python bench_r1cs.py --steps 1 --warmup 1 --per-gpu 64 --no-cpu > /dev/null 2>&1
ncu -i /tmp/ncu/r1.ncu-rep --page raw --csv > /tmp/ncu/r1_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("/tmp/ncu/r1_raw.csv")))
h=rows[0]
want=("Kernel Name","dram__bytes_read.sum","dram__bytes_write.sum","gpu__time_duration.sum","inst_executed","l1tex__t_sector_hit_rate.pct","launch__block_size","launch__grid_size","launch__registers_per_thread","lts__t_sector_hit_rate.pct","lts__throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","thread_inst_executed")
keep=[i for i,c in enumerate(h) if c in want or ("pcsamp_warps_issue_stalled" in c and "not_issued" not in c)]
w=csv.writer(open("gpurun_out/ncu_r02_r1cs_raw_excerpt.csv","w"))
for r in rows: w.writerow([r[i][:60] for i in keep])
PY
python bench_configs.py --cfg 5,5r --steps 2 > gpurun_out/bench_configs_r02_r1cs.jsonl 2>/dev/null; cut -c1-200 gpurun_out/bench_configs_r02_r1cs.jsonl
python profiles/tools/wit_bench.py 512 fused 2>&1 | grep -E "grid|matches" | tail -4
python profiles/tools/wit_bench.py 2048 2>&1 | grep -E "grid|matches" | tail -2
ls -la gpurun_out | head -30
