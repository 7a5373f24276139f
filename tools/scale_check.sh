#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): sliced-pair parity, weak scaling of the default bench, column slices
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( timeout 600 python -m pytest tests/test_psa_multigpu.py -q 2>&1 | tail -3 )
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pair_$n.json 2> gpurun_out/scale_pair_$n.err
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --workload pair100k_bands > gpurun_out/scale_bands_$n.json 2> gpurun_out/scale_bands_$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_pair_$n.json 2> gpurun_out/scale_pair_$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 5 --warmup 3 --workload pair100k_bands > gpurun_out/scale_bands_$n.json 2> gpurun_out/scale_bands_$n.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 3 --warmup 3 --workload batch10k > gpurun_out/scale_batch10k_$n.json 2> gpurun_out/scale_batch10k_$n.err
  fi
  for w in pair bands batch10k; do f=gpurun_out/scale_${w}_$n.json; [ -s $f ] && python -c "
import json,sys
d=json.load(open('$f')); print('$w', d['n_gpus'], 'GPUs: value %.0f e2e %.0f GCUPS  ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; done
done
