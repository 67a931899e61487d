set -x
mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on -k regex:"^k_miller$|^k_final_exp" -c 2 -o /tmp/ncu/mil python bench.py --n 65536 --lanes 1 --steps 1 --warmup 3 --skip-extra --no-cpu > /dev/null 2>&1
ncu -i /tmp/ncu/mil.ncu-rep --page source --csv -k k_miller 2>/dev/null | gzip > gpurun_out/r2_k_miller_source.csv.gz
ncu -i /tmp/ncu/mil.ncu-rep --page source --csv -k k_final_exp 2>/dev/null | gzip > gpurun_out/r2_k_final_exp_source.csv.gz
ls -la gpurun_out/*.gz
