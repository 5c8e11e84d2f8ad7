#!/bin/bash
# conv1 as the direct fp32 kernel (k_cf_direct): parity tests of the conv file, then the suite
OUT=gpurun_out/r02_call35
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "first_conv or neighbour_tables" 2>&1 | tail -5 | tee $OUT/pytest_conv1.txt
bash tools/gpu_suite.sh r02_call35 pytest
