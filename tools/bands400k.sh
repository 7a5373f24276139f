#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --workload pair400k_bands > gpurun_out/bands400k_1.json 2> gpurun_out/bands400k_1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --workload pair400k_bands > gpurun_out/bands400k_$N.json 2> gpurun_out/bands400k_$N.err
for n in 1 $N; do python -c "
import json
l=[x for x in open('gpurun_out/bands400k_$n.json') if x.startswith('{')]
d=json.loads(l[-1]); print($n, 'GPUs: value %.0f e2e %.0f GCUPS ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['result'], d['config'].get('cols_per_lane'), d['config'].get('bands_per_gpu'))" || tail -n 5 gpurun_out/bands400k_$n.err; done
