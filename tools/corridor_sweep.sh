#!/bin/bash
# traced single pairs over corridor margins (TSTA_PSA_CORRIDOR): ms per step and corridors needed
mkdir -p gpurun_out
{
for w in pair200k_trace pair100k_trace; do
  for c in "" 1536 3072 4096 5120 6144 8192; do
    if [ -z "$c" ]; then unset TSTA_PSA_CORRIDOR; else export TSTA_PSA_CORRIDOR=$c; fi
    echo "== $w corridor ${c:-default}"
    python bench.py --workload $w --steps 3 --warmup 2 --no-suite --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms %.3f e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k:v for k,v in d.items() if 'trace' in k or 'corridor' in k or 'tiles' in k})"
  done
done
} > gpurun_out/corridor.log 2>&1
