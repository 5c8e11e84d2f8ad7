#!/bin/bash
# Round 2, first GPU call: suite with the un-gated batched tests on the collapsed (former variant y) build, bench in batched mode,
# ncu launch list of the batched step, batched-size conv micro-benchmarks.
OUT=gpurun_out/r02_call1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --batched 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench.json; tail -3 $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_batched.csv python bench.py --profile --batched 10 --steps 1 --warmup 1 > $OUT/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches_batched.csv > $OUT/launches_batched_summary.txt 2>&1; head -60 $OUT/launches_batched_summary.txt
for F in 1 10; do
  timeout 300 python tools/conv_g4_bench.py --frags $F --cin 64 --cout 64 --flags 0,2,4,6,7 > $OUT/conv64_frags$F.txt 2>&1; cat $OUT/conv64_frags$F.txt
  timeout 300 python tools/conv_g4_bench.py --frags $F --cin 32 --cout 32 --flags 0,2,4,6,7 > $OUT/conv32_frags$F.txt 2>&1; cat $OUT/conv32_frags$F.txt
done
timeout 300 python tools/conv_g4_bench.py --frags 10 --n 14107 --cin 64 --cout 64 --flags 0,7 2>&1 | tee $OUT/conv64_l2.txt
timeout 300 python tools/conv_g4_bench.py --frags 10 --n 3765 --cin 128 --cout 128 --flags 0,7 2>&1 | tee $OUT/conv128_l4.txt
timeout 300 python tools/conv_g4_bench.py --frags 10 --n 1085 --cin 256 --cout 256 --flags 0,7 2>&1 | tee $OUT/conv256_l8.txt
ls -la $OUT
