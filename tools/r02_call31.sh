#!/bin/bash
OUT=gpurun_out/r02_call31
mkdir -p $OUT
for C in 64 32; do
  for F in 0 7; do
    timeout 300 python tools/conv_g4_bench.py --frags 10 --cin $C --cout $C --reps 5 --flags $F --trace 2>&1 | tee $OUT/trace_${C}_flags$F.txt | tail -60
  done
done
