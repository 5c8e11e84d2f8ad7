#!/bin/bash
# vectorised LayerNorm rows, coalesced K / V pack: module timings + the suite
OUT=gpurun_out/r02_call48
mkdir -p $OUT
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
bash tools/gpu_suite.sh r02_call48 pytest
