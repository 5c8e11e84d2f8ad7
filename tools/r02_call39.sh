#!/bin/bash
# attention kernel alone (ncu gpu__time_duration of the three shapes) under the IMF_FF_DBG experiment switches
OUT=gpurun_out/r02_call39
mkdir -p $OUT
for D in 0 3 4 8 12 16 32 64 127; do
  IMF_FF_DBG=$D timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_flash_fusion --csv --log-file $OUT/t_$D.csv python tools/flash_bench.py --profile > /dev/null 2>&1
  echo "IMF_FF_DBG=$D: $(grep k_flash_fusion $OUT/t_$D.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')" | tee -a $OUT/flash_dbg_kernel_times.txt
done
