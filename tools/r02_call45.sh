#!/bin/bash
# attention: pass 1 over block pairs (N = 128 products): parity, kernel times of the three shapes, module timings, ncu --set full
OUT=gpurun_out/r02_call45
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_gpu_batched.py -m gpu -q -x 2>&1 | tail -3 | tee $OUT/pytest.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:k_flash_fusion --csv --log-file $OUT/t.csv python tools/flash_bench.py --profile > /dev/null 2>&1
echo "k_flash_fusion ns (1x1085, 10x1085, 8192): $(grep k_flash_fusion $OUT/t.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')" | tee $OUT/flash_kernel_times.txt
python tools/flash_bench.py 2>&1 | tee $OUT/flash_bench.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:k_flash_fusion -c 3 -o $OUT/flash -f python tools/flash_bench.py --profile > $OUT/ncu_flash.log 2>&1; echo "flash rc=$?"
ncu -i $OUT/flash.ncu-rep --page raw --csv > $OUT/flash.raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/flash.raw.csv 2>&1 | tee $OUT/ncu_flash_summary.txt
