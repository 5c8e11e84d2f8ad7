#!/bin/bash
OUT=gpurun_out/r02_call21
mkdir -p $OUT
python tools/image_conv_bench.py 2>&1 | tee $OUT/image_conv_bench.txt
timeout 600 ncu --set full --clock-control none --profile-from-start off --import-source on -k regex:k_image_conv3x3_p8 -c 1 -o $OUT/p8 -f python tools/image_conv_bench.py --profile > $OUT/ncu_p8.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/p8.ncu-rep --page raw --csv > $OUT/p8.raw.csv 2>/dev/null
ncu -i $OUT/p8.ncu-rep --page source --csv > $OUT/p8.source.csv 2>/dev/null
rm -f $OUT/p8.ncu-rep
