#!/bin/bash
OUT=gpurun_out/r02_call4
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_batched.py::test_batched_attention_fusion_all_items_one_launch_vs_oracle tests/test_gpu_forward.py::test_attention_fusion_matches_reference_golden tests/test_gpu_forward.py::test_attention_ragged_sizes_vs_oracle tests/test_gpu_forward.py::test_attention_stress_8192x4800_vs_oracle tests/test_gpu_conv.py -k "attention or first_conv" -x -q > $OUT/pytest_first.log 2>&1; echo "first rc=$?"; tail -15 $OUT/pytest_first.log
timeout 300 python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
timeout 300 python tools/profile_single.py 2>&1 | tee $OUT/single_latency.txt
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/single_launches.csv python tools/profile_single.py > /dev/null 2>&1
python tools/summarize_launches.py $OUT/single_launches.csv > $OUT/single_launches_summary.txt 2>&1; head -40 $OUT/single_launches_summary.txt
bash tools/gpu_suite.sh r02_call4
