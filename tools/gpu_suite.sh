#!/bin/bash
# One GPU call: the -m gpu suite, bench.py (our arm), an ncu launch list of the timed steps only, summaries under gpurun_out/<tag>/.
#   bash tools/gpu_suite.sh <tag> [pytest|nopytest] [extra bench args]
TAG=${1:-suite}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ "${2:-pytest}" = "pytest" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 20 ${@:3} > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python tools/show_bench.py $OUT/bench.json | tee $OUT/bench_summary.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv python bench.py --profile --steps 2 --warmup 2 ${@:3} > $OUT/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; head -45 $OUT/launches_summary.txt
