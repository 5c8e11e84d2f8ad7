"""Time the fused ResNet stem (csrc/stem_fused.cu) alone with CUDA events: ten 640x480 frames per launch, as the batched plan runs it.

    python tools/stem_bench.py            # presplit + conv per launch, median of 20
    python tools/stem_bench.py --profile  # one launch between cudaProfilerStart/Stop (for ncu)"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib

profile = "--profile" in sys.argv
H, W, B = 480, 640, 10
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
img = torch.rand(B, 3, H, W, generator=g).cuda()
w7 = torch.zeros(8, 8, 4, 64)
w7[:7, 1:, :3, :] = torch.randn(7, 7, 3, 64, generator=g) / 12
w7 = w7.reshape(4, 64, 64).contiguous().cuda()
packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(4, 64, 64, 64)), dtype=torch.uint8, device="cuda")
_lib.check(L.imf_sparse_conv_h2_pack(w7.data_ptr(), 4, 64, 64, 64, 1024.0, packed.data_ptr(), s))
sc, sh = torch.full((64,), 1.0 / 1024, device="cuda"), torch.zeros(64, device="cuda")
H1, W1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
ws_bytes = int(L.imf_image_stem_workspace_bytes(H, W, B))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
Y = torch.zeros(B * H1 * W1, 64, device="cuda")
err = torch.zeros(1, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    _lib.check(L.imf_image_stem_h2_fwd(img.data_ptr(), H, W, B, packed.data_ptr(), sc.data_ptr(), sh.data_ptr(), ws.data_ptr(), ws_bytes,
                                       Y.data_ptr(), 128, err.data_ptr(), s))


for _ in range(3):
    run()
torch.cuda.synchronize()
if profile:
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
for cold in (False, True):
    ts = []
    for _ in range(20):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    px = B * H1 * W1
    t = float(np.median(ts))
    print(f"stem {B} x {H}x{W}: presplit + conv {t:.1f} us ({'L2 flushed' if cold else 'warm'}); output {px * 256 / 1e6:.0f} MB "
          f"-> {px * 256 / t / 1e3:.0f} GB/s of output; {2.0 * px * 147 * 64 / t / 1e6:.1f} TFLOP/s algorithmic", flush=True)
assert int(err.item()) == 0
