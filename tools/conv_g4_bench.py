"""Time the TMA-gather sparse convolution (imf_sparse_conv_g4_fwd) on one layer shape with CUDA events (run on the GPU box).

    python tools/conv_g4_bench.py [--n 50000] [--cin 64] [--cout 64] [--reps 20] [--grid 0] [--trace]
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
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--cin", type=int, default=64)
ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--flush", type=int, default=1)
ap.add_argument("--grid", type=int, default=0)
ap.add_argument("--trace", action="store_true")
ap.add_argument("--npw", type=int, default=0, help="minimum stages per split CTA (profiling hook)")
ap.add_argument("--flags", default="0", help="comma list of profiling flags: 1 no W copies, 2 no gathers, 4 no MMAs")
ap.add_argument("--residual", action="store_true", help="with a residual input (the second convolution of a block); flag 32 = no L2 prefetch of it")
ap.add_argument("--frags", type=int, default=1, help="fragments of --n voxels in one launch (batch index in column 0), as the batched plan runs it")
args = ap.parse_args()

L = _lib.lib()
parts = []
for b in range(args.frags):
    c, _ = synthetic.make_fragment(args.n, 0.025, b)
    c[:, 0] = b
    parts.append(c)
coords = np.concatenate(parts)
cm = CoordinateManager(torch.from_numpy(coords).cuda())
nbr_t, ld_n, tile_mask = cm.table_t(1, 1, 3, False)
n, cin, cout = len(coords), args.cin, args.cout
kci, kco = (64 if cin % 64 == 0 else 32), (64 if cout % 64 == 0 else 32)
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, cin, device="cuda", generator=g)
W = torch.randn(27, cin, cout, device="cuda", generator=g) / np.sqrt(27 * cin)
s = torch.cuda.current_stream().cuda_stream
Xh = torch.zeros(n, 2 * cin, dtype=torch.float16, device="cuda")
_lib.check(L.imf_h2_pack(X.data_ptr(), cin, n, cin, kci, Xh.data_ptr(), 2 * cin, None, s))
packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(27, cin, cout, kci)), dtype=torch.uint8, device="cuda")
_lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), 27, cin, cout, kci, 1024.0, packed.data_ptr(), s))
Yh = torch.zeros(n, 2 * cout, dtype=torch.float16, device="cuda")
one, zero = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
Rh = torch.zeros(n, 2 * cout, dtype=torch.float16, device="cuda")
_lib.check(L.imf_h2_pack(torch.randn(n, cout, device="cuda", generator=g).data_ptr(), cout, n, cout, kco, Rh.data_ptr(), 2 * cout, None, s))
err = torch.zeros(1, dtype=torch.int32, device="cuda")
ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(cout))
ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pairs = int((nbr_t[:, :n] >= 0).sum())
alg = 4 * pairs * cin + 8 * pairs + 4 * 27 * cin * cout + 4 * n * cout


def g4():
    _lib.check(L.imf_sparse_conv_g4_fwd(Xh.data_ptr(), 2 * cin, kci, packed.data_ptr(), nbr_t.data_ptr(), ld_n, tile_mask.data_ptr(), None, n,
                                        27, cin, cout, one.data_ptr(), zero.data_ptr(), Rh.data_ptr() if args.residual else None, 2 * cout, kco, 1, Yh.data_ptr(), 2 * cout, n, kco,
                                        ws.data_ptr(), ws_bytes, err.data_ptr(), s))


def timeit(fn, label):
    ts = []
    for i in range(args.reps + 3):
        if args.flush:
            flushbuf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{label}: n={n} {cin}->{cout} pairs={pairs}: median {np.median(ts):.1f} us  min {np.min(ts):.1f} us  alg {alg / 1e6:.1f} MB -> "
          f"{alg / np.median(ts) / 1e3:.0f} GB/s   {2 * pairs * cin * cout / np.median(ts) / 1e6:.1f} TFLOP/s (algorithmic)", flush=True)


for fl in [int(f) for f in args.flags.split(",")]:
    L.imf_debug_conv_g4_trace(None, args.grid, args.npw, fl)
    timeit(g4, f"g4 grid={args.grid or 148} npw={args.npw or 8} flags={fl}")
    if fl in (0, 32):
        assert int(err.item()) == 0, int(err.item())
    err.zero_()          # (with profiling flags the results are meaningless and may leave the fp16 range)
L.imf_debug_conv_g4_trace(None, args.grid, args.npw, 0)
if args.trace:      # needs a library built with IMFNET_B200_NVCC_FLAGS=-DIMF_G4_TRACE
    trace = torch.zeros(1024, dtype=torch.int64, device="cuda")
    L.imf_debug_conv_g4_trace(trace.data_ptr(), args.grid, args.npw, int(args.flags.split(",")[-1]))
    g4()
    torch.cuda.synchronize()
    L.imf_debug_conv_g4_trace(None, 0, 0, 0)
    t = trace.cpu().numpy()
    life = t[256:256 + 148]
    ends = t[256 + 320:256 + 320 + 148]
    busy = life[life > 0]
    print(f"per-CTA lifetime (cycles): n={len(busy)} min {busy.min()} median {int(np.median(busy))} max {busy.max()}; "
          f"end-time spread (globaltimer ns): {int(ends[ends > 0].max() - ends[ends > 0].min())}")
    names = ["start", "barriers+TMEM", "masks", "epi wait", "acc ready", "epilogue done", "exit"]
    print("CTA0 timeline (cycles):", {nm: int(t[i] - t[0]) for i, nm in enumerate(names)})
    for j in range(8):
        r = t[160 + 8 * j: 166 + 8 * j]
        if r[0] == 0:
            break
        print(f"  epilogue sub-tile {j}: buffer free {int(r[0] - t[0])}  TMEM loaded {int(r[1] - t[0])}  staged {int(r[2] - t[0])}  "
              f"fenced {int(r[3] - t[0])}  all warps staged {int(r[4] - t[0])}  stores issued {int(r[5] - t[0])}")
    for i in range(36):
        r = t[16 + 4 * i: 20 + 4 * i]
        if r[2] == 0:
            break
        print(f"  stage {i:3d}: producer at wait {int(r[0] - t[0]):7d}  slot free {int(r[1] - t[0]):7d}   mma saw full {int(r[2] - t[0]):7d}  committed {int(r[3] - t[0]):7d}")
