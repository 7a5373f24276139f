#!/bin/bash
# every BASELINE workload once, with the CPU baseline (for profiles/r01_bench_all.jsonl)
mkdir -p gpurun_out; : > gpurun_out/bench_all.jsonl
for w in pair100k pair100k_trace pair200k_trace batch250 batch10k batch250_trace batch10k_trace msa16x50k; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w 2> gpurun_out/bench_all_$w.err | grep '^{' >> gpurun_out/bench_all.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l); cb=d.get('cpu_baseline') or {}
    print('%-60s value %8.1f e2e %8.1f cpu %s' % (d['config']['workload'][:60], d['value'], d['e2e']['value'], cb.get('value')))
PY
