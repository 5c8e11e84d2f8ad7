#!/bin/bash
OUT=gpurun_out/r02_call15
mkdir -p $OUT
python tools/stem_bench.py 2>&1 | tee $OUT/stem_bench.txt
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "fused_stem" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python tools/show_bench.py $OUT/bench.json | head -3
