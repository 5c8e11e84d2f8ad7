#!/bin/bash
# attention: clock64 timeline of CTA 0 (trace build)
OUT=gpurun_out/r02_call40
mkdir -p $OUT
timeout 300 python tools/flash_trace.py 2>&1 | tee $OUT/flash_trace.txt | head -5
IMF_FF_DBG=127 timeout 300 python tools/flash_trace.py 2>&1 > $OUT/flash_trace_dbg127.txt
