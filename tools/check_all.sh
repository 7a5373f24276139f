#!/bin/bash
# quick GPU regression pass: CLI + batch tests, batch e2e with pinned buffers
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cli.py tests/test_psa_batch_gpu.py -x -q 2>&1 | tail -5 ) 
for w in batch250 batch10k; do timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], 'value %.0f e2e %.0f GCUPS ms %.1f e2e_ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"; done
