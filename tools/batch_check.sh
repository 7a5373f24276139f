timeout 900 python -m pytest tests/test_psa_batch_gpu.py tests/test_cli.py -x -q 2>&1 | tail -3
for w in batch250 batch250_trace batch10k; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w --no-cpu-baseline 2> gpurun_out/bc_$w.err | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-50s value %8.1f e2e %8.1f' % (d['config']['workload'][:50], d['value'], d['e2e']['value']))"
done
