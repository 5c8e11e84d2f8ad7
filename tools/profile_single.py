"""Launch list of ONE single-fragment forward() (model(x, image), the reference's own call form) for `ncu --profile-from-start off`:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_single.py
Without ncu it prints the latency (CUDA events, L2 flushed, median of 10)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

import imfnet_b200.me as ME
from imfnet_b200 import load_model, synthetic

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0), strict=True)
model = model.eval().cuda()
coords, feats, image = synthetic.make_config(cfg, seed=0)
coords, feats, image = coords.cuda(), feats.cuda(), image.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        model(ME.SparseTensor(feats, coordinates=coords), image)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model(ME.SparseTensor(feats, coordinates=coords), image)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{cfg}: single-fragment forward latency median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms")
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(ME.SparseTensor(feats, coordinates=coords), image)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
