#!/bin/bash
# call 9: full -m gpu suite (incl. the rescaled-checkpoint range tests) + bench at 3 plans in flight
OUT=gpurun_out/r02_call9
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python tools/show_bench.py $OUT/bench.json | tee $OUT/bench_summary.txt
