#!/bin/bash
# evidence at the final state of the round: suite (pytest -m gpu, bench, ncu launch list), conv microbench (skeleton share) at the batched
# size, bit-reproducibility over 100 launches, ncu --set full per kernel family
OUT=gpurun_out/r02_call43
mkdir -p $OUT
bash tools/gpu_suite.sh r02_call43 pytest
for C in 64 32; do
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --flags 0,1,2,4,7 2>&1 | tee $OUT/conv_g4_bench_${C}_batched.txt
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --residual --flags 0,32 2>&1 | tee $OUT/conv_g4_bench_${C}_batched_residual.txt
done
timeout 600 python tools/conv_g4_check.py --reps 100 --modes 0 2>&1 | cut -c1-400 | tee $OUT/conv_g4_check_100.txt
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
bash tools/ncu_kernels.sh r02_call43
