"""Throughput of TMA tile::gather4 row gathers per SM (run on the B200 box): python tools/tma_rate.py"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from imfnet_b200 import _lib

L = _lib.lib()
n = 50000
for ld in (256,):
    X = torch.randn(n, ld, device="cuda").half()
    for pattern in ("random", "half_absent", "half_hot64", "half_hot8192", "all_hot64", "seq_rows"):
        if pattern == "local":
            idx = (torch.arange(128 * 4096, device="cuda") % n).to(torch.int32)
        else:
            idx = torch.randint(0, n, (128 * 4096,), dtype=torch.int32, device="cuda")
        if pattern == "half_absent":
            idx[torch.rand(idx.numel(), device="cuda") < 0.5] = -1
        if pattern == "absent_ge_n":
            idx[torch.rand(idx.numel(), device="cuda") < 0.5] = n
        m = torch.rand(idx.numel(), device="cuda") < 0.5
        ar = torch.arange(idx.numel(), device="cuda", dtype=torch.int32)
        if pattern == "half_hot64":
            idx[m] = (ar % 64)[m]
        if pattern == "half_hot8192":
            idx[m] = (ar % 8192)[m]
        if pattern == "all_hot64":
            idx[:] = ar % 64
        if pattern == "seq_rows":
            idx[:] = (ar // 27) % n
        for nctas in (148,):
            for nwarps, depth in ((8, 1), (4, 3)):
                out = torch.zeros(8, dtype=torch.int64, device="cuda")
                err = torch.zeros(1, dtype=torch.int32, device="cuda")
                iters = 64
                _lib.check(L.imf_debug_gather4_rate(X.data_ptr(), ld, n, idx.data_ptr(), idx.numel(), nwarps, iters, depth, nctas, out.data_ptr(),
                                                    err.data_ptr(), _lib.cur_stream()))
                torch.cuda.synchronize()
                cyc = int(out.max())
                ops = nwarps * iters * 32
                print(f"ld={ld} {pattern:12s} ctas={nctas:3d} warps={nwarps} depth={depth}: {cyc} cycles, {cyc / ops:.1f} cycles/gather4, "
                      f"{ops * 512 / cyc:.1f} B/cycle/SM", flush=True)
