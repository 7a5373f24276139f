mkdir -p gpurun_out
N=${1:-2}
echo "== bands x$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload pair100k_bands 2> gpurun_out/bench_bands$N.err | tee gpurun_out/bench_bands$N.json | cut -c1-1800; grep -i "error" gpurun_out/bench_bands$N.err | tail -n 3
