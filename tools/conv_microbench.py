"""Time one sparse convolution layer of the h2 tensor-core tier with CUDA events (run on the GPU box).

    python tools/conv_microbench.py [--n 50000] [--cin 64] [--cout 64] [--flags 0,1,2,3,4,7] [--reps 20]

--flags: profiling-hook values (include/imfnet_b200.h::imf_debug_conv_flags) to attribute the time: 1 = no weight copies,
2 = no gathers, 4 = no MMAs.
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
ap.add_argument("--flags", default="0")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--flush", type=int, default=1)
args = ap.parse_args()

L = _lib.lib()
coords, _ = synthetic.make_fragment(args.n, 0.025, 0)
cm = CoordinateManager(torch.from_numpy(coords).cuda())
nbr = cm.table(1, 1, 3, False)
n, cin, cout = len(coords), args.cin, args.cout
kci, kco = (64 if cin % 64 == 0 else 32), (64 if cout % 64 == 0 else 32)
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, cin, device="cuda", generator=g)
W = torch.randn(27, cin, cout, device="cuda", generator=g) / np.sqrt(27 * cin)
Xh = torch.zeros(n, 2 * cin, dtype=torch.float16, device="cuda")
_lib.check(L.imf_h2_pack(X.data_ptr(), cin, n, cin, kci, Xh.data_ptr(), 2 * cin, None, _lib.cur_stream()))
packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(27, cin, cout, kci)), dtype=torch.uint8, device="cuda")
_lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), 27, cin, cout, kci, 1024.0, packed.data_ptr(), _lib.cur_stream()))
Yh = torch.zeros(n, 2 * cout, dtype=torch.float16, device="cuda")
one, zero = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pairs = int((nbr >= 0).sum())
alg = 4 * pairs * cin + 8 * pairs + 4 * 27 * cin * cout + 4 * n * cout
s = torch.cuda.current_stream().cuda_stream
for fl in [int(f) for f in args.flags.split(",")]:
    L.imf_debug_conv_flags(fl)
    ts = []
    for i in range(args.reps + 3):
        if args.flush:
            flushbuf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.imf_sparse_conv_h2_fwd(Xh.data_ptr(), 2 * cin, kci, packed.data_ptr(), nbr.data_ptr(), None, n, 27, cin, cout,
                                            one.data_ptr(), zero.data_ptr(), None, 0, 0, 1, Yh.data_ptr(), 2 * cout, kco, None, 0,
                                            err.data_ptr(), s))
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    L.imf_debug_conv_flags(0)
    print(f"n={n} {cin}->{cout} pairs={pairs} flags={fl}: median {np.median(ts):.1f} us  min {np.min(ts):.1f} us  "
          f"alg {alg / 1e6:.1f} MB -> {alg / np.median(ts) / 1e3:.0f} GB/s", flush=True)

# pipeline timeline of CTA 0 (clock64 stamps, cycles relative to kernel start)
trace = torch.zeros(16 + 4 * 27 * max(1, cin // kci) + 8, dtype=torch.int64, device="cuda")
L.imf_debug_conv_trace(trace.data_ptr())
_lib.check(L.imf_sparse_conv_h2_fwd(Xh.data_ptr(), 2 * cin, kci, packed.data_ptr(), nbr.data_ptr(), None, n, 27, cin, cout,
                                    one.data_ptr(), zero.data_ptr(), None, 0, 0, 1, Yh.data_ptr(), 2 * cout, kco, None, 0,
                                    err.data_ptr(), s))
torch.cuda.synchronize()
L.imf_debug_conv_trace(None)
t = trace.cpu().numpy()
t0 = t[0]
names = ["start", "barriers+TMEM", "nbr staged", "stage list", "producer done", "acc ready", "epilogue done", "exit"]
print("CTA0 timeline (cycles):", {nm: int(t[i] - t0) for i, nm in enumerate(names)})
for i in range(27 * max(1, cin // kci)):
    r = t[16 + 4 * i: 20 + 4 * i]
    if r[0] == 0:
        break
    print(f"  stage {i:3d}: slot {int(r[0]-t0):7d}  issued {int(r[1]-t0):7d}  landed {int(r[2]-t0):7d}  mma-saw-full {int(r[3]-t0):7d}")
