#!/bin/bash
# First GPU call of the next round: verify what was written without GPU access at the end of round 1, in ONE gpurun call.
#   1. the regular GPU suite (must stay green), 2. the gated tests of the batched captured plan, 3. bench: default vs --batched,
#   4. the kernel variant library x (lean producer addressing, zero fill skipped for clean rows): conv microbench, parity, GPU tests.
# Usage (repo root on the box): bash tools/gpu_round2_first.sh <tag>
TAG=${1:-r02_first}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
IMFNET_B200_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_batched.py -m gpu -x -q -s > $OUT/pytest_batched.log 2>&1; echo "batched rc=$?" | tee -a $OUT/pytest_batched.log; tail -15 $OUT/pytest_batched.log
timeout 600 python bench.py --steps 20 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench (auto probe) rc=$?"; cut -c1-400 $OUT/bench_default.json
timeout 300 python bench.py --probe-batched > $OUT/probe.json 2> $OUT/probe.err; echo "probe rc=$?"; cat $OUT/probe.json; tail -5 $OUT/probe.err
for B in 5 10 16; do
  timeout 300 python bench.py --steps 20 --batched $B > $OUT/bench_batched_$B.json 2> $OUT/bench_batched_$B.err; echo "bench --batched $B rc=$?"
  python - $OUT/bench_batched_$B.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    for n, m in d["config"]["execution_modes_timed"].items():
        print("  %-22s %2d fragments/step: %.1f M voxels/s resident, %.1f M end to end" % (n, m["fragments_per_step"], m["voxels_per_s"] / 1e6, m["voxels_per_s_e2e"] / 1e6))
    print("  note:", d["config"]["mode_selection"])
except Exception as e:
    print("  no result:", e)
PY
done
tail -5 $OUT/bench_batched_10.err
# conv-kernel experiment: rebuild with the switch into a scratch copy of the library, run the microbench and the parity checker
timeout 300 python tools/conv_g4_bench.py > $OUT/conv_g4_default.txt 2>&1; cat $OUT/conv_g4_default.txt
timeout 300 python tools/conv_g4_bench.py --cin 32 --cout 32 --trace > $OUT/conv_g4_trace_32.txt 2>&1; head -50 $OUT/conv_g4_trace_32.txt
for V in x z; do     # x: lean producer addressing, vector residual loads, uniform MMA issue; z: x + zero fill skipped for clean rows
  export IMFNET_B200_VARIANT=$V
  timeout 300 python tools/conv_g4_bench.py > $OUT/conv_g4_variant_$V.txt 2>&1; cat $OUT/conv_g4_variant_$V.txt
  timeout 300 python tools/conv_g4_bench.py --cin 32 --cout 32 > $OUT/conv_g4_variant_${V}_32.txt 2>&1; cat $OUT/conv_g4_variant_${V}_32.txt
  timeout 300 python tools/conv_g4_check.py > $OUT/conv_g4_check_variant_$V.txt 2>&1; tail -8 $OUT/conv_g4_check_variant_$V.txt
  timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_forward.py -m gpu -x -q > $OUT/pytest_variant_$V.log 2>&1; echo "gpu tests (variant $V) rc=$?"; tail -3 $OUT/pytest_variant_$V.log
  timeout 300 python bench.py --steps 20 --batched 0 > $OUT/bench_variant_$V.json 2> $OUT/bench_variant_$V.err; echo "bench (variant $V) rc=$?"; cut -c1-200 $OUT/bench_variant_$V.json
done
timeout 300 python tools/flash_bench.py > $OUT/flash_bench_variant_z.txt 2>&1; cat $OUT/flash_bench_variant_z.txt
export IMFNET_B200_VARIANT=y      # x + early hand-off of the MMA warps' turn (protocol change: check reproducibility first)
timeout 300 python tools/conv_g4_check.py > $OUT/conv_g4_check_variant_y.txt 2>&1; tail -8 $OUT/conv_g4_check_variant_y.txt
timeout 300 python tools/conv_g4_bench.py > $OUT/conv_g4_variant_y.txt 2>&1; cat $OUT/conv_g4_variant_y.txt
timeout 300 python tools/conv_g4_bench.py --cin 32 --cout 32 > $OUT/conv_g4_variant_y_32.txt 2>&1; cat $OUT/conv_g4_variant_y_32.txt
unset IMFNET_B200_VARIANT
ls -la $OUT
