#!/bin/bash
# attention traffic hypothesis: module timings with the V / K loads switched off after the rings are primed (IMF_FF_DBG, results meaningless)
OUT=gpurun_out/r02_call37
mkdir -p $OUT
for D in 0 1 2 3; do
  echo "IMF_FF_DBG=$D" | tee -a $OUT/flash_dbg.txt
  IMF_FF_DBG=$D timeout 300 python tools/flash_bench.py 2>&1 | tee -a $OUT/flash_dbg.txt
done
