"""Run-to-run reproducibility and accuracy of imf_sparse_conv_g4_fwd for the MMA issue-assignment modes (run on the GPU box).

    python tools/conv_g4_check.py [--reps 12]

Modes (profiling flags of imf_debug_conv_g4_trace): 0 = stage -> MMA warp by sub-tile (production), 16 = by stage number,
8 = one issuing warp.  For each (rows, channels, mode): bitwise comparison of `reps` launches with the first one and the
row-wise relative error against a float64 torch evaluation of the same sum.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib, synthetic
from imfnet_b200.sparse import CoordinateManager

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--modes", default="0,16,8")
args = ap.parse_args()
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream

for n_req, cin, cout in ((50000, 64, 64), (50000, 32, 32), (50000, 64, 128), (13000, 64, 128), (3000, 128, 128), (700, 128, 128)):
    coords, _ = synthetic.make_fragment(n_req, 0.025, 0)
    cm = CoordinateManager(torch.from_numpy(coords).cuda())
    nbr_t, ld_n, tile_mask = cm.table_t(1, 1, 3, False)
    n = len(coords)
    kci, kco = (64 if cin % 64 == 0 else 32), (64 if cout % 64 == 0 else 32)
    g = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn(n, cin, device="cuda", generator=g)
    W = torch.randn(27, cin, cout, device="cuda", generator=g) / np.sqrt(27 * cin)
    Xh = torch.zeros(n, 2 * cin, dtype=torch.float16, device="cuda")
    _lib.check(L.imf_h2_pack(X.data_ptr(), cin, n, cin, kci, Xh.data_ptr(), 2 * cin, None, s))
    packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(27, cin, cout, kci)), dtype=torch.uint8, device="cuda")
    _lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), 27, cin, cout, kci, 1024.0, packed.data_ptr(), s))
    one, zero = torch.full((cout,), 1.0 / 1024.0, device="cuda"), torch.zeros(cout, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(cout))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    # float64 reference
    ref = torch.zeros(n, cout, dtype=torch.float64, device="cuda")
    Xd, Wd = X.double(), W.double()
    for k in range(27):
        idx = nbr_t[k, :n].long()
        ok = idx >= 0
        ref[ok] += Xd[idx[ok]] @ Wd[k]
    ref = ref.float()
    for mode in [int(m) for m in args.modes.split(",")]:
        L.imf_debug_conv_g4_trace(None, 0, 0, mode)
        outs = []
        for r in range(args.reps):
            Yh = torch.full((n, 2 * cout), float("nan"), dtype=torch.float16, device="cuda")
            _lib.check(L.imf_sparse_conv_g4_fwd(Xh.data_ptr(), 2 * cin, kci, packed.data_ptr(), nbr_t.data_ptr(), ld_n, tile_mask.data_ptr(),
                                                None, n, 27, cin, cout, one.data_ptr(), zero.data_ptr(), None, 0, 0, 0, Yh.data_ptr(),
                                                2 * cout, n, kco, ws.data_ptr(), ws_bytes, err.data_ptr(), s))
            torch.cuda.synchronize()
            outs.append(Yh)
        Y = torch.empty(n, cout, device="cuda")
        _lib.check(L.imf_h2_unpack(outs[0].data_ptr(), 2 * cout, n, cout, kco, Y.data_ptr(), cout, s))
        rel = float(((Y - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-20)).max())
        ndiff = [int((o.view(torch.int16) != outs[0].view(torch.int16)).any(dim=1).sum()) for o in outs[1:]]
        print(f"n={n} {cin}->{cout} mode={mode}: max row rel err {rel:.2e}; rows differing from launch 0 in later launches: {ndiff}; "
              f"err flag {int(err.item())}", flush=True)
L.imf_debug_conv_g4_trace(None, 0, 0, 0)
