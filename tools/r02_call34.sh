#!/bin/bash
# residual sub-tiles through the free ring slots; index staging in shared memory again for KC = 32
OUT=gpurun_out/r02_call34
mkdir -p $OUT
timeout 600 python tools/conv_g4_check.py --reps 30 --modes 0 2>&1 | cut -c1-120 | tee $OUT/conv_g4_check.txt
for C in 64 32; do
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --flags 0,7 2>&1 | tail -2 | tee $OUT/conv_g4_bench_${C}_batched.txt
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --residual --flags 0,32 2>&1 | tail -2 | tee $OUT/conv_g4_bench_${C}_batched_residual.txt
done
bash tools/gpu_suite.sh r02_call34 pytest
