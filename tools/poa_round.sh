#!/bin/bash
# POA evidence after the byte-key rewrite: tests, ncu capture of a late-round fill, the msa16x50k bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_poa_gpu.py -x -q 2>&1 | tail -2
bash tools/prof_poa.sh | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --workload msa16x50k 2> gpurun_out/msa_bench.err | grep '^{' > gpurun_out/msa_bench.json
python -c "
import json
d=json.load(open('gpurun_out/msa_bench.json')); print('msa value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d.get('driver_loop_s_per_run'), d.get('cpu_baseline'))"
