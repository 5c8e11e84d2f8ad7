#!/bin/bash
OUT=gpurun_out/r02_call5
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_batched.py -k "first_conv or attention or image" -x -q > $OUT/pytest_first.log 2>&1; echo "first rc=$?"; tail -5 $OUT/pytest_first.log
timeout 300 python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
timeout 300 python tools/profile_single.py 2>&1 | tee $OUT/single_latency.txt
bash tools/gpu_suite.sh r02_call5
bash tools/ncu_kernels.sh r02_call5
