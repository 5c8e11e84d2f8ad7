#!/bin/bash
# call 12: fused stem kernel + parity scan -- unit tests first, then the suite, bench and a launch list
OUT=gpurun_out/r02_call12
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "fused_stem" > $OUT/pytest_stem.log 2>&1; echo "stem rc=$?"; tail -12 $OUT/pytest_stem.log
bash tools/gpu_suite.sh r02_call12 pytest
