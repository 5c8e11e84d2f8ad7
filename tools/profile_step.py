"""Where does one forward go?  Host enqueue time vs device time, per phase (run on the GPU box).

    python tools/profile_step.py [--config C2] [--iters 20]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

import imfnet_b200.me as ME
from imfnet_b200 import load_model, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--impl", default="tc")
args = ap.parse_args()

target, voxel, W, H = synthetic.CONFIGS[args.config]
frags = []
for i in range(8):
    c, _ = synthetic.make_fragment(target, voxel, i)
    frags.append((torch.from_numpy(c).cuda(), torch.ones((len(c), 1), device="cuda"), synthetic.make_image(W, H, i).cuda()))
model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0))
model = model.eval().cuda()

with torch.no_grad():
    for i in range(10):
        c, f, im = frags[i % 8]
        model(ME.SparseTensor(f, coordinates=c), im)
    torch.cuda.synchronize()
    host, dev = [], []
    for i in range(args.iters):
        c, f, im = frags[i % 8]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        model(ME.SparseTensor(f, coordinates=c), im)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3)
        dev.append(e0.elapsed_time(e1))
    print(f"forward: host enqueue {np.median(host):.3f} ms, device {np.median(dev):.3f} ms (median of {args.iters})")
    # image branch alone
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img = model.img_encoder(frags[0][2])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"image encoder alone: {np.median(ts):.3f} ms")
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kv = model.attention_fusion.project_context(img[0].reshape(128, -1), True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"kv projection alone: {np.median(ts):.3f} ms")
    q = torch.randn(1085, 256, device="cuda")
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.attention_fusion.fuse(q, kv)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"attention fuse (M=1085) alone: {np.median(ts):.3f} ms")
    # coordinate phase alone
    from imfnet_b200.sparse import CoordinateManager
    ts = []
    for i in range(10):
        c = frags[i % 8][0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cm = CoordinateManager(c)
        cm.build_pyramid([2, 4, 8])
        for t in (1, 2, 4, 8):
            cm.table(t, t, 3, False)
        for a, b in ((1, 2), (2, 4), (4, 8)):
            cm.table(a, b, 3, False)
            cm.table(b, a, 3, True)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"coordinate phase alone (hash + pyramid + 10 tables): {np.median(ts):.3f} ms")
