mkdir -p gpurun_out
for w in batch250 batch10k batch250_trace; do
  echo "== $w"; timeout 300 python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | cut -c1-300; tail -n 3 gpurun_out/bench_$w.err
done
echo "== bands x2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload pair100k_bands 2> gpurun_out/bench_bands2.err | tee gpurun_out/bench_bands2.json | cut -c1-1200; tail -n 3 gpurun_out/bench_bands2.err
echo "== pair x2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_pair2.err | tee gpurun_out/bench_pair2.json | cut -c1-400; tail -n 3 gpurun_out/bench_pair2.err
