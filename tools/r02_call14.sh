#!/bin/bash
OUT=gpurun_out/r02_call14
mkdir -p $OUT
python tools/stem_bench.py 2>&1 | tee $OUT/stem_bench.txt
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "fused_stem" 2>&1 | tail -3
bash tools/gpu_suite.sh r02_call14 pytest
