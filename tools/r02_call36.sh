#!/bin/bash
# attention: P as a tensor-memory A operand (parity of the forward / batched tests, module timings, ncu of the attention kernel)
OUT=gpurun_out/r02_call36
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:k_flash_fusion -c 3 -o $OUT/flash -f python tools/flash_bench.py --profile > $OUT/ncu_flash.log 2>&1; echo "flash rc=$?"
ncu -i $OUT/flash.ncu-rep --page raw --csv > $OUT/flash.raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/flash.raw.csv 2>&1 | tee $OUT/ncu_flash_summary.txt
