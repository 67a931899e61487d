python -m pytest tests -x -q -m gpu -k "witness or r1cs" 2>&1 | tail -2
timeout 400 python bench_configs.py --cfg 5,5r --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms','witnesses_per_sec','assignments_per_sec')})
    elif 'rror' in l: print(l.strip()[:300])
"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_r1cs -c 60 --csv --log-file gpurun_out/l5r.csv python bench_configs.py --cfg 5r --steps 1 --scale 0.25 > gpurun_out/l5r.log 2>&1
python - <<EOF2
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/l5r.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]: d[r[ki].split("(")[0]].append(float(r[vi].replace(",","")))
for k,v in d.items(): print(k, len(v), "avg us", round(sum(v)/len(v)/1e3,1))
EOF2
