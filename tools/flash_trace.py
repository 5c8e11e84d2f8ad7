"""clock64 timeline of CTA 0 of the attention kernel at the stress shape (needs a library built with
IMFNET_B200_NVCC_FLAGS=-DIMF_FF_TRACE; run on the GPU box):  python tools/flash_trace.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib, load_model, synthetic

model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0))
af = model.eval().cuda().attention_fusion
L = _lib.lib()
raw = L
s = torch.cuda.current_stream().cuda_stream
w, wp = af.packed(), af.packed_h2()
sizes, Lt = [8192], 4800
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
_lib.check(L.imf_attention_kv_batched(w, wp, tok.data_ptr(), Lt, B, kv.data_ptr(), kws.data_ptr(), kws.numel(), err.data_ptr(), s))
for _ in range(3):
    _lib.check(L.imf_attention_fusion_fwd_batched(w, wp, P.data_ptr(), 256, n, m_dev.data_ptr(), seg.data_ptr(), cnt.data_ptr(), B, kv.data_ptr(), Lt,
                                                  out.data_ptr(), 256, ws.data_ptr(), ws.numel(), err.data_ptr(), s))
torch.cuda.synchronize()
t = np.zeros(8 * 256, dtype=np.int64)
if not hasattr(raw, "imf_debug_flash_trace"):
    sys.exit('this library has no trace hooks: rebuild with IMFNET_B200_NVCC_FLAGS="-DIMF_FF_TRACE" python -m imfnet_b200.build --force')
fn = raw.imf_debug_flash_trace
fn.argtypes = [C.c_void_p, C.c_int]
assert fn(t.ctypes.data, t.size) == 0
t = t.reshape(256, 8)
t0 = t[t > 0].min()
names = ["S issue begin", "S issue end", "PV issue begin", "PV issue end", "softmax: S ready", "softmax: S read", "softmax w2: P arrived", "softmax w17: P arrived"]
print("index (running S / block count of CTA 0): " + " | ".join(names))
for i in range(80):
    print(f"{i:3d} " + " ".join(f"{(int(v - t0) if v > 0 else -1):8d}" for v in t[i]))
