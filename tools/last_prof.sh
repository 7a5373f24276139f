#!/bin/bash
mkdir -p gpurun_out
python tools/batch_time.py 200000 250 x -4 > gpurun_out/lp_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:psa_batch16_dual -s 1 -c 1 -f -o gpurun_out/r01_psa_batch16_dual_250bp \
    python tools/batch_time.py 200000 250 x -4 > gpurun_out/lp_ncu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --workload msa16x50k 2> gpurun_out/lp_msa.err | grep '^{' > gpurun_out/lp_msa.json
python -c "
import json
d=json.load(open('gpurun_out/lp_msa.json')); print('msa value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d.get('driver_loop_s_per_run'), d.get('cpu_baseline'))"
tail -n 2 gpurun_out/lp_ncu.log
