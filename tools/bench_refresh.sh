#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/bench_refresh.jsonl
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tail -8
for w in batch250 batch250_trace batch10k_trace msa16x50k; do
  timeout 900 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline 2> gpurun_out/bench_refresh_$w.err | grep '^{' >> gpurun_out/bench_refresh.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_refresh.jsonl'):
    d=json.loads(l)
    print('%-60s value %8.1f e2e %8.1f' % (d['config']['workload'][:60], d['value'], d['e2e']['value']), d.get('driver_loop_s_per_run'), d.get('process_wall_s_per_run'))
PY
