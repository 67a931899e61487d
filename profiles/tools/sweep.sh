#!/bin/bash
for so in build/lib_*.so; do
  echo "== $so"
  BLSGPU_SO=$so timeout 300 python bench.py --steps 1 --warmup 3 --n $1 --no-cpu 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k: round(v,1) for k,v in d['stage_ms'].items()}, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
    elif 'rror' in l: print(l.strip()[:200])
"
done
