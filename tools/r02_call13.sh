#!/bin/bash
OUT=gpurun_out/r02_call13
mkdir -p $OUT
python tools/stem_bench.py 2>&1 | tee $OUT/stem_bench.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:k_stem_conv -c 1 -o $OUT/stem -f python tools/stem_bench.py --profile > $OUT/ncu_stem.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/stem.ncu-rep --page raw --csv > $OUT/stem.raw.csv 2>/dev/null
ncu -i $OUT/stem.ncu-rep --page source --csv > $OUT/stem.source.csv 2>/dev/null
ls -la $OUT
timeout 300 python -m pytest tests/test_gpu_coords.py tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
