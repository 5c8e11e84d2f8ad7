#!/bin/bash
# ncu --set full captures of the kernel families of one batched forward (bench.py --profile: only the steps between
# cudaProfilerStart/Stop are seen) and of the attention kernel at the stress shape.  Reports land in gpurun_out/<tag>/ and are
# summarised on the build box with tools/ncu_summary.py.
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
COMMON="--set full --clock-control none --profile-from-start off"
timeout 900 ncu $COMMON -k regex:k_sparse_conv_g4 -c 46 -o $OUT/g4 -f python bench.py --profile --steps 1 --warmup 2 > $OUT/ncu_g4.log 2>&1; echo "g4 rc=$?"
timeout 900 ncu $COMMON -k regex:'k_cf_|k_tail_fused|k_stem_|k_kernel_map_t|k_h2_gemm|k_flash|k_image_|k_layernorm|k_parity|k_stride|k_hash|k_compact|k_flag|k_h2_|k_batch|k_nn_|k_match' -c 64 -o $OUT/others -f python bench.py --profile --steps 1 --warmup 2 > $OUT/ncu_others.log 2>&1; echo "others rc=$?"
timeout 600 ncu $COMMON --import-source on -k regex:k_flash_fusion -c 3 -o $OUT/flash -f python tools/flash_bench.py --profile > $OUT/ncu_flash.log 2>&1; echo "flash rc=$?"
# the raw pages are what gets summarised; the big reports stay on the box (gpurun_out is limited to 64 MiB)
for R in g4 others flash; do
  ncu -i $OUT/$R.ncu-rep --page raw --csv > $OUT/$R.raw.csv 2>/dev/null
done
rm -f $OUT/g4.ncu-rep $OUT/others.ncu-rep
ls -la $OUT
