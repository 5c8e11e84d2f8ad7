#!/bin/bash
# safety checks of the final library: smoke(), the reference arm, the N = 2 bench under torchrun, the C4 config
OUT=gpurun_out/r02_call44
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$OUT/ref.err | tee $OUT/bench_reference.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>$OUT/n2.err | tee $OUT/bench_n2.json | cut -c1-400
timeout 600 python bench.py --config C4 --steps 5 --warmup 3 2>$OUT/c4.err | tee $OUT/bench_c4.json | cut -c1-400
