#!/bin/bash
OUT=gpurun_out/r02_call30
mkdir -p $OUT
timeout 600 python tools/conv_g4_check.py --reps 30 --modes 0 2>&1 | cut -c1-120 | tee $OUT/conv_g4_check.txt
for C in 64 32; do
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --flags 0,7 2>&1 | tail -2 | tee $OUT/conv_g4_bench_${C}_batched.txt
done
bash tools/gpu_suite.sh r02_call30 pytest
