"""tcgen05 GEMM self-test on the GPU box (separate process: a protocol bug traps the kernel and poisons the context).

    python tools/tc_selftest.py            # prints one line per shape: max abs err / max|ref| against fp64
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

from imfnet_b200 import _lib


def run(M, N, K, geglu=False, bias=True, res=True, split=False, lda_pad=0, seed=0):
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K + lda_pad, device="cuda", generator=g)
    rowsB = 2 * N if geglu else N
    B = torch.randn(rowsB, K + lda_pad, device="cuda", generator=g) / (K ** 0.5)
    b = torch.randn(rowsB, device="cuda", generator=g) if bias else None
    R = torch.randn(M, N, device="cuda", generator=g) if res else None
    C = torch.full((M, N), float("nan"), device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.imf_tc_gemm_workspace_bytes(M, N, K)) if split else 0
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
    rc = L.imf_tc_gemm(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), C.data_ptr(), N, M, N, K, 0.5,
                       _lib.ptr(b), _lib.ptr(R), N if res else 0, int(geglu), ws.data_ptr() if split else None, ws_bytes,
                       err.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, L.imf_last_error()
    torch.cuda.synchronize()
    Ad, Bd = A[:, :K].double(), B[:, :K].double()
    ref = 0.5 * (Ad @ Bd.t())
    if geglu:
        v, gt = ref[:, :N], ref[:, N:]
        if bias:
            v, gt = v + b[:N].double(), gt + b[N:].double()
        ref = v * torch.nn.functional.gelu(gt)
    elif bias:
        ref = ref + b.double()
    if res:
        ref = ref + R.double()
    e = float((C.double() - ref).abs().max() / ref.abs().max())
    fp32 = 0.5 * (A[:, :K] @ B[:, :K].t())
    print(f"M={M:5d} N={N:5d} K={K:5d} geglu={int(geglu)} split={int(split)} pad={lda_pad}: rel err {e:.3e}  err_flag={int(err.item())}", flush=True)
    return e


if __name__ == "__main__":
    torch.backends.cuda.matmul.allow_tf32 = False
    worst = 0.0
    for args in [dict(M=128, N=64, K=32, bias=False, res=False), dict(M=128, N=128, K=128), dict(M=1085, N=128, K=256),
                 dict(M=300, N=256, K=128), dict(M=1085, N=4800, K=128, bias=False, res=False),
                 dict(M=1085, N=128, K=4800, split=True), dict(M=333, N=128, K=302, split=True, lda_pad=2),
                 dict(M=1085, N=1024, K=256, geglu=True, res=False), dict(M=1085, N=256, K=1024, split=True),
                 dict(M=1, N=7, K=5, lda_pad=3), dict(M=8192, N=4800, K=128, bias=False, res=False)]:
        worst = max(worst, run(**args))
    print("WORST", worst)
    sys.exit(0 if worst < 2e-6 else 1)
