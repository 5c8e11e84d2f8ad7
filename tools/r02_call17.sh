#!/bin/bash
OUT=gpurun_out/r02_call17
mkdir -p $OUT
python tools/stem_bench.py 2>&1 | tee $OUT/stem_bench.txt
bash tools/gpu_suite.sh r02_call17 pytest
