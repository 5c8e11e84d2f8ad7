#!/bin/bash
OUT=gpurun_out/r02_call18
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 20 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python tools/show_bench.py $OUT/bench.json | head -4
timeout 600 python bench.py --steps 20 --plans 4 > $OUT/bench_p4.json 2> $OUT/bench_p4.err; python tools/show_bench.py $OUT/bench_p4.json | head -2
timeout 600 python bench.py --steps 10 --config C4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo "c4 rc=$?"; python tools/show_bench.py $OUT/bench_c4.json | head -2
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
