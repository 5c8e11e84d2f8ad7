"""Time one layer1 convolution (csrc/image_conv_p8.cu: 3x3 / 1, 64 -> 64, ten 160x120 feature maps per launch) with CUDA events.

    python tools/image_conv_bench.py            # median of 20, with and without residual
    python tools/image_conv_bench.py --profile  # one launch between cudaProfilerStart/Stop (for ncu)"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib

profile = "--profile" in sys.argv
H, W, B = 120, 160, 10
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
w9 = (torch.randn(9, 64, 64, generator=g) / 24).cuda()
packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(9, 64, 64, 64)), dtype=torch.uint8, device="cuda")
_lib.check(L.imf_sparse_conv_h2_pack(w9.data_ptr(), 9, 64, 64, 64, 1024.0, packed.data_ptr(), s))
sc, sh = torch.full((64,), 1.0 / 1024, device="cuda"), torch.zeros(64, device="cuda")
nbytes = int(L.imf_image_p8_bytes(H, W, B))
x = torch.zeros(B, 16, H + 2, W + 2, 8, dtype=torch.float16, device="cuda")
x[:, :, 1:-1, 1:-1] = (torch.randn(B, 16, H, W, 8, generator=g) * 0.5).half().cuda()
x[:, 8:] *= 1e-3
assert x.numel() * 2 == nbytes
r = x.clone()
y = torch.zeros_like(x)
err = torch.zeros(1, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(res):
    _lib.check(L.imf_image_conv3x3_p8_fwd(x.data_ptr(), H, W, B, packed.data_ptr(), sc.data_ptr(), sh.data_ptr(), r.data_ptr() if res else None, 1,
                                          y.data_ptr(), 0, 0, err.data_ptr(), s))


for _ in range(3):
    run(True)
torch.cuda.synchronize()
if profile:
    torch.cuda.profiler.start()
    run(True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
for res in (False, True):
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(res)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = float(np.median(ts))
    fl = 2.0 * B * H * W * 9 * 64 * 64
    print(f"conv3x3 p8 {B} x {H}x{W} x 64 -> 64{' + residual' if res else ''}: {t:.1f} us, {fl / t / 1e6:.1f} TFLOP/s algorithmic "
          f"({3 * fl / t / 1e6:.0f} executed on the split products)", flush=True)
assert int(err.item()) == 0
