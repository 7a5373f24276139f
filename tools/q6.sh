mkdir -p gpurun_out
for w in pair100k pair100k_trace batch250 batch10k msa16x10k; do
  echo "== $w"; timeout 600 python bench.py --steps 3 --warmup 3 --workload $w 2> gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json | cut -c1-1500; tail -3 gpurun_out/bench_$w.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600
