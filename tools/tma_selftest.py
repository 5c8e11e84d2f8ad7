"""GPU self-test of the TMA tile::gather4 + tiled store path (run on the B200 box).

    python tools/tma_selftest.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

from imfnet_b200 import _lib

L = _lib.lib()
torch.manual_seed(0)
n, ld = 1000, 256
X = torch.randn(n, ld, device="cuda").half()
idx = torch.randint(0, n, (128,), dtype=torch.int32, device="cuda")
idx[5] = -1
idx[77] = n          # first row outside the tensor
idx[100] = n + 12345
ok_all = True
for box_rows in (1,):      # the gather map needs a one-row box (a 4-row box raises an illegal-instruction fault)
    for col in (0, 64, 192):
        raw = torch.zeros(128 * 64, dtype=torch.float16, device="cuda")
        O = torch.full((300, ld), 7.0, dtype=torch.float16, device="cuda")
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        try:
            _lib.check(L.imf_debug_gather4(X.data_ptr(), ld, n, idx.data_ptr(), col, box_rows, raw.data_ptr(), O.data_ptr(), 300, 200,
                                           err.data_ptr(), _lib.cur_stream()))
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            print(f"box_rows={box_rows} col={col}: FAILED {e}")
            ok_all = False
            continue
        raw = raw.cpu().numpy().reshape(128, 8, 8)          # [row][16-byte chunk position][8 halves]
        Xh, ih = X.cpu().numpy(), idx.cpu().numpy()
        exp = np.zeros((128, 64), np.float16)
        for r in range(128):
            if 0 <= ih[r] < n:
                exp[r] = Xh[ih[r], col:col + 64]
        unsw = np.zeros_like(exp)
        for r in range(128):
            for c in range(8):
                unsw[r, c * 8:(c + 1) * 8] = raw[r, c ^ (r & 7)]
        ok_g = np.array_equal(unsw, exp)
        Oh = O.cpu().numpy()
        ok_s = np.array_equal(Oh[200:300, col:col + 64], exp[:100]) and np.all(Oh[:200] == 7.0) and \
            np.all(np.delete(Oh[200:300], np.s_[col:col + 64], axis=1) == 7.0)
        print(f"box_rows={box_rows} col={col}: gather {'OK' if ok_g else 'MISMATCH'}  store(clipped) {'OK' if ok_s else 'MISMATCH'}  err={int(err.item())}")
        if not ok_g:
            bad = np.argwhere(unsw != exp)
            print("   first mismatches:", bad[:5].tolist(), "zero rows ok:", [bool(np.all(unsw[r] == 0)) for r in (5, 77, 100)])
        ok_all &= ok_g and ok_s
print("TMA SELFTEST", "PASS" if ok_all else "FAIL")
