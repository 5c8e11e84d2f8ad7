#!/bin/bash
# final evidence: skeleton share of the conv kernel at the batched size, bit-reproducibility over 100 launches, ncu --set full per kernel family
OUT=gpurun_out/r02_call28
mkdir -p $OUT
for C in 64 32; do
  timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --flags 0,1,2,4,7 2>&1 | tee $OUT/conv_g4_bench_${C}_batched.txt
done
timeout 600 python tools/conv_g4_check.py --reps 100 --modes 0 2>&1 | tee $OUT/conv_g4_check_100.txt
python tools/image_conv_bench.py 2>&1 | tee $OUT/image_conv_bench.txt
python tools/stem_bench.py 2>&1 | tee $OUT/stem_bench.txt
bash tools/ncu_kernels.sh r02_call28
