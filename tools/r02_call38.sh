#!/bin/bash
# attention: what a block's time consists of (IMF_FF_DBG experiment switches; results meaningless except for 0)
OUT=gpurun_out/r02_call38
mkdir -p $OUT
for D in 0 3 7 11 15 19 35 67 127; do
  echo "IMF_FF_DBG=$D" | tee -a $OUT/flash_dbg.txt
  IMF_FF_DBG=$D timeout 300 python tools/flash_bench.py 2>&1 | grep "8192" | tee -a $OUT/flash_dbg.txt
done
