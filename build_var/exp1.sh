for v in "" build_var/lib_r1m6.so build_var/lib_r1m8.so; do
  echo "== r1cs $v"
  BLSGPU_SO=${v:+$PWD/$v} timeout 400 python bench_configs.py --cfg 5r --steps 3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms','witnesses_per_sec','assignments_per_sec')})
    elif 'rror' in l: print(l.strip()[:300])
"
done
for v in "" build_var/lib_w256.so build_var/lib_w128x4.so; do
  for n in 512 2048; do
  echo "== wit $v n=$n"
  BLSGPU_SO=${v:+$PWD/$v} timeout 400 python profiles/tools/wit_bench.py $n 2>&1 | grep -E "witness_gen|rror|matches" | tail -3
  done
done
