#!/bin/bash
# attention: clock64 timeline of CTA 0 (trace build), single issuer, pass 1 over block pairs
OUT=gpurun_out/r02_call46
mkdir -p $OUT
timeout 300 python tools/flash_trace.py > $OUT/flash_trace.txt 2>&1
tail -3 $OUT/flash_trace.txt
