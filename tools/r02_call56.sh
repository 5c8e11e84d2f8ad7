#!/bin/bash
# final state: suite (pytest -m gpu with the streaming regression test, bench, ncu launch list), attention module timings + ncu --set full
OUT=gpurun_out/r02_call56
mkdir -p $OUT
bash tools/gpu_suite.sh r02_call56 pytest
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:k_flash_fusion -c 3 -o $OUT/flash -f python tools/flash_bench.py --profile > $OUT/ncu_flash.log 2>&1; echo "flash rc=$?"
ncu -i $OUT/flash.ncu-rep --page raw --csv > $OUT/flash.raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/flash.raw.csv 2>&1 | tee $OUT/ncu_flash_summary.txt
rm -f $OUT/flash.ncu-rep
