#!/bin/bash
# bench with the streaming warm-up: six runs (scatter of `value`), one C4 run
OUT=gpurun_out/r02_call62
mkdir -p $OUT
for i in 1 2 3 4 5 6; do
  timeout 300 python bench.py --steps 30 > $OUT/bench_$i.json 2> $OUT/bench_$i.err; RC=$?
  echo "bench $i rc=$RC $(python tools/show_bench.py $OUT/bench_$i.json 2>/dev/null | head -1 | cut -c1-90)" | tee -a $OUT/stress.txt
done
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 2>$OUT/c4.err | cut -c1-200
