#!/bin/bash
# First GPU call of the next round: verify what was written without GPU access at the end of round 1, in ONE gpurun call.
#   1. the regular GPU suite (must stay green)                      2. the gated tests of the batched captured plan
#   3. bench.py in its automatic mode: probes the default library, the variant libraries x / z / y / w and the batched plan
#      (config.mode_probes: step time, 64->64 and 32->32 conv launches, debug-flag sweep, batched verdict per library)
#   4. run-to-run reproducibility + accuracy of the conv kernel for every variant (tools/conv_g4_check.py)
#   5. clock64 trace of the 32->32 launch with the default library (open question)
# Usage (repo root on the box): bash tools/gpu_round2_first.sh <tag>
TAG=${1:-r02_first}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
IMFNET_B200_UNVERIFIED=1 timeout 300 python -m pytest tests/test_gpu_batched.py -m gpu -x -q -s > $OUT/pytest_batched.log 2>&1; echo "batched rc=$?" | tee -a $OUT/pytest_batched.log; tail -15 $OUT/pytest_batched.log
timeout 900 python bench.py --steps 20 > $OUT/bench_auto.json 2> $OUT/bench_auto.err; echo "bench (automatic mode) rc=$?"; tail -3 $OUT/bench_auto.err
python - $OUT/bench_auto.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print("value %.1f M voxels/s, e2e %.1f M, library %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["config"]["library"]))
    for n, m in d["config"]["execution_modes_timed"].items():
        print("  %-22s %2d fragments/step: %.1f M voxels/s resident, %.1f M end to end" % (n, m["fragments_per_step"], m["voxels_per_s"] / 1e6, m["voxels_per_s_e2e"] / 1e6))
    print("  selection:", d["config"]["mode_selection"])
    for n, r in (d["config"]["mode_probes"] or {}).items():
        print("  probe %-8s %s" % (n, json.dumps(r)))
except Exception as e:
    print("  no result:", e)
PY
timeout 300 python bench.py --steps 20 --batched 0 > $OUT/bench_default_only.json 2> $OUT/bench_default_only.err; echo "bench --batched 0 rc=$?"; cut -c1-160 $OUT/bench_default_only.json
for V in x z y w; do
  IMFNET_B200_VARIANT=$V timeout 300 python tools/conv_g4_check.py --modes 0 > $OUT/conv_g4_check_variant_$V.txt 2>&1; echo "conv_g4_check variant $V rc=$?"; tail -6 $OUT/conv_g4_check_variant_$V.txt
done
timeout 300 python tools/conv_g4_bench.py --cin 32 --cout 32 --trace > $OUT/conv_g4_trace_32.txt 2>&1; head -50 $OUT/conv_g4_trace_32.txt
ls -la $OUT
