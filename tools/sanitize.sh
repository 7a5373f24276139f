#!/bin/bash
# memcheck of the kernels changed late in the round (plain run first, then the same command under compute-sanitizer)
mkdir -p gpurun_out
CMD="python -m pytest tests/test_poa_gpu.py tests/test_psa_batch_gpu.py -x -q -k random_dags_or_quirk_or_chunked"
CMD="python -m pytest tests/test_poa_gpu.py tests/test_psa_batch_gpu.py -x -q -k 'random_dags or quirk or chunked'"
eval timeout 600 $CMD 2>&1 | tail -2 || exit 1
eval timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 $CMD > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize.log | tail -3
