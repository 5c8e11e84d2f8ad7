#!/bin/bash
OUT=gpurun_out/r02_call22
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "dense_grid or first_conv" > $OUT/pytest_dense.log 2>&1; echo "dense rc=$?"; tail -15 $OUT/pytest_dense.log
bash tools/gpu_suite.sh r02_call22 pytest
