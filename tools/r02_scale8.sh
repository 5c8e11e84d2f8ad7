#!/bin/bash
# 8-GPU box: C2 at N = 8 and 4, C4 (fragment pairs + matching) at N = 8; one process per GPU under torchrun
OUT=gpurun_out/r02_scale8
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
run() {  # n, tag, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 20 --warmup 5 ${@:3} > $OUT/$2.json 2> $OUT/$2.err
  echo "$2 rc=$?"; python tools/show_bench.py $OUT/$2.json 2>/dev/null | head -1
}
run 8 c2_n8
run 8 c4_n8 --config C4 --steps 10
run 4 c2_n4
timeout 600 python bench.py --steps 20 > $OUT/c2_n1.json 2> $OUT/c2_n1.err; python tools/show_bench.py $OUT/c2_n1.json 2>/dev/null | head -1
