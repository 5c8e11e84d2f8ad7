#!/bin/bash
# call 10: fused tail kernel -- its unit test first, then the whole -m gpu suite, then bench
OUT=gpurun_out/r02_call10
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "fused_tail" > $OUT/pytest_tail.log 2>&1; echo "tail rc=$?"; tail -12 $OUT/pytest_tail.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python tools/show_bench.py $OUT/bench.json | tee $OUT/bench_summary.txt
