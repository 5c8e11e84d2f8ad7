"""Time the fused attention kernel path (imf_attention_fusion_fwd on M point tokens x L image tokens) with CUDA events."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import load_model, synthetic

model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0))
af = model.eval().cuda().attention_fusion
for M, L in ((1085, 4800), (8192, 4800)):
    P = torch.randn(M, 256, device="cuda")
    I = torch.randn(L, 128, device="cuda")
    kv = af.project_context(I, False)
    for _ in range(3):
        af.fuse(P, kv)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        af.fuse(P, kv)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    flops_attn = 4.0 * M * L * 128
    print(f"fuse M={M} L={L}: median {np.median(ts):.1f} us (whole module: LN, to_q, attention, to_out, FFN); "
          f"QK^T+PV = {flops_attn / 1e9:.2f} GFLOP algorithmic", flush=True)
