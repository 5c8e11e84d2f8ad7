#!/bin/bash
OUT=gpurun_out/r02_call8
mkdir -p $OUT
for P in 2 3; do
  timeout 600 python bench.py --steps 20 --plans $P > $OUT/bench_plans$P.json 2> $OUT/bench_plans$P.err; tail -2 $OUT/bench_plans$P.err
  python tools/show_bench.py $OUT/bench_plans$P.json 2>/dev/null | head -3
done
bash tools/ncu_kernels.sh r02_call8
