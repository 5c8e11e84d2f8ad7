#!/bin/bash
# call 20: layer1 on the plane-layout implicit GEMM -- unit tests, suite, bench, launch list
OUT=gpurun_out/r02_call20
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "plane_layout" > $OUT/pytest_p8.log 2>&1; echo "p8 rc=$?"; tail -15 $OUT/pytest_p8.log
bash tools/gpu_suite.sh r02_call20 pytest
