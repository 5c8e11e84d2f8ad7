#!/bin/bash
# One GPU-box call: parity tests, bench, step profile, conv microbench, ncu launch list, ncu full capture of the top kernel.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 2500 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; echo "reference rc=$?"; cat $OUT/bench_reference.json | cut -c1-300
for k in 1 6 10; do timeout 200 python bench.py --steps 20 --streams $k --batched 0 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print(\"streams\",d[\"config\"][\"streams\"],d[\"value\"],d[\"ms_per_step\"],d[\"e2e\"][\"value\"])"; done | tee $OUT/streams_sweep.txt
timeout 300 python tools/profile_step.py > $OUT/profile_step.txt 2>&1; cat $OUT/profile_step.txt
timeout 300 python tools/conv_g4_bench.py --flags 0,3,6,7 --old > $OUT/conv_g4_bench_64.txt 2>&1; cat $OUT/conv_g4_bench_64.txt
timeout 300 python tools/conv_g4_bench.py --cin 32 --cout 32 > $OUT/conv_g4_bench_32.txt 2>&1; cat $OUT/conv_g4_bench_32.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --profile --steps 2 --warmup 1 --streams 1 > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sparse_conv_g4 -s 4 -c 2 -o $OUT/prof_conv_g4 \
    python tools/conv_g4_bench.py --reps 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
