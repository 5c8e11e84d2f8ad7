#!/bin/bash
OUT=gpurun_out/r02_call6
mkdir -p $OUT
for CPV in 64 512; do
  echo "== cells per voxel $CPV"
  IMF_CF_CELLS_PER_VOXEL=$CPV timeout 600 python bench.py --steps 10 > $OUT/bench_$CPV.json 2> $OUT/bench_$CPV.err; tail -2 $OUT/bench_$CPV.err
  python tools/show_bench.py $OUT/bench_$CPV.json | head -8
  IMF_CF_CELLS_PER_VOXEL=$CPV timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:k_cf_ --csv --log-file $OUT/cf_$CPV.csv python bench.py --profile --steps 1 --warmup 2 > /dev/null 2>&1
  python tools/summarize_launches.py $OUT/cf_$CPV.csv
done
