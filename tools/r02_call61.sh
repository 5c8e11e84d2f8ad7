#!/bin/bash
# the force-rebuilt final library: eight bench runs in a row, the streaming regression test twice
OUT=gpurun_out/r02_call61
mkdir -p $OUT
for i in 1 2 3 4 5 6 7 8; do
  S=$SECONDS
  timeout 300 python bench.py --steps 30 > $OUT/bench_$i.json 2> $OUT/bench_$i.err; RC=$?
  echo "bench $i rc=$RC $((SECONDS - S)) s $(python tools/show_bench.py $OUT/bench_$i.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a $OUT/stress.txt
done
for i in 1 2; do timeout 300 python -m pytest tests/test_gpu_batched.py -m gpu -q -x -k "bench_shape" 2>&1 | tail -1 | tee -a $OUT/stress.txt; done
