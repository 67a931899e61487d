#!/bin/bash
# usage: cmp.sh N lib1 lib2 ...   -- stage times per library at batch size N
N=$1; shift
for so in "$@"; do
  echo "== $so"
  BLSGPU_SO=$so timeout 300 python bench.py --steps 2 --warmup 3 --n $N --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v,1) for k,v in d['stage_ms'].items()}, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
