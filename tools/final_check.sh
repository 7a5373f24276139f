#!/bin/bash
# final single-GPU pass of the round: full GPU suite, smoke, default bench (both arms)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/final_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
cat gpurun_out/final_pytest.log; tail -n 1 gpurun_out/final_smoke.log; cut -c1-400 gpurun_out/final_bench.json; cut -c1-200 gpurun_out/final_bench_reference.json
