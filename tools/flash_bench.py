"""Time the fusion module and its attention kernel with CUDA events (run on the GPU box).

    python tools/flash_bench.py            # whole module (LN, to_q, attention, to_out, FFN) + the attention launch pair alone
    python tools/flash_bench.py --profile  # one attention launch of every shape between cudaProfilerStart/Stop (for ncu)
Shapes: C2 single fragment (1085 x 4800), a batch of 10 such items in one launch, BASELINE's stress shape (8192 x 4800)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib, load_model, synthetic

profile = "--profile" in sys.argv
model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0))
af = model.eval().cuda().attention_fusion
L = _lib.lib()
s = torch.cuda.current_stream().cuda_stream
w = af.packed()
wp = None if "--tf32" in sys.argv else af.packed_h2()          # --tf32: the 3xTF32 GEMM tier instead of the h2 GEMMs


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


for sizes, Lt in (([1085], 4800), ([1085] * 10, 4800), ([8192], 4800)):
    B, n = len(sizes), sum(sizes)
    P = torch.randn(n, 256, device="cuda")
    tok = torch.randn(B * Lt, 128, device="cuda")
    seg = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device="cuda")
    cnt = torch.tensor(sizes, dtype=torch.int32, device="cuda")
    m_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    kv = torch.empty(int(L.imf_attention_kv_batched_bytes(Lt, B)), dtype=torch.uint8, device="cuda")
    kws = torch.empty(int(L.imf_attention_kv_batched_workspace_bytes(Lt, 128, 128, B)), dtype=torch.uint8, device="cuda")
    ws = torch.empty(int(L.imf_attention_batched_workspace_bytes(n, Lt, 256, 128, B)), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, 256, device="cuda")

    def kvproj():
        _lib.check(L.imf_attention_kv_batched(w, wp, tok.data_ptr(), Lt, B, kv.data_ptr(), kws.data_ptr(), kws.numel(), err.data_ptr(), s))

    def module():
        _lib.check(L.imf_attention_fusion_fwd_batched(w, wp, P.data_ptr(), 256, n, m_dev.data_ptr(), seg.data_ptr(), cnt.data_ptr(), B, kv.data_ptr(), Lt,
                                                      out.data_ptr(), 256, ws.data_ptr(), ws.numel(), err.data_ptr(), s))

    kvproj()
    if profile:
        module()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        module()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    t_kv, t_mod = timeit(kvproj), timeit(module)
    assert int(err.item()) == 0, int(err.item())
    fl_attn = 4.0 * n * Lt * 128
    fl_rest = 2.0 * n * (256 * 128 + 128 * 256 + 256 * 2048 + 1024 * 256)
    print(f"items {B} x {sizes[0]} queries x {Lt} tokens: K/V projection {t_kv:.1f} us; module {t_mod:.1f} us "
          f"(QK^T+PV {fl_attn / 1e9:.2f} GFLOP, projections+FFN {fl_rest / 1e9:.2f} GFLOP algorithmic -> "
          f"{(fl_attn + fl_rest) / t_mod / 1e6:.1f} TFLOP/s)", flush=True)
