"""CPU: host logic of the batched captured plan (imfnet_b200/batched.py) -- replication of the image encoder's closed-form
neighbour tables over B images, checked against a brute-force table of the stacked images."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from imfnet_b200.batched import BatchedImagePlan


def conv_table(Hin, Win, K, stride, pad):
    """numpy restatement of csrc/image_ops.cu::k_image_conv_table: [K*K, ld] offset-major, -1 = outside, ld padded to 128."""
    Hout, Wout = (Hin + 2 * pad - K) // stride + 1, (Win + 2 * pad - K) // stride + 1
    n = Hout * Wout
    ld = (n + 127) // 128 * 128
    t = np.full((K * K, ld), -1, dtype=np.int32)
    for k in range(K * K):
        for o in range(n):
            oy, ox = divmod(o, Wout)
            iy, ix = oy * stride - pad + k // K, ox * stride - pad + k % K
            if 0 <= iy < Hin and 0 <= ix < Win:
                t[k, o] = iy * Win + ix
    return t, n


@pytest.mark.parametrize("geom", [(13, 9, 3, 1, 1), (13, 9, 3, 2, 1), (13, 9, 1, 2, 0), (12, 11, 1, 1, 0)])
@pytest.mark.parametrize("B", [1, 3])
def test_replicated_table_equals_per_image_tables(geom, B):
    Hin, Win, K, stride, pad = geom
    t, p_out = conv_table(Hin, Win, K, stride, pad)
    p_in = Hin * Win
    fake = SimpleNamespace(B=B)
    out, ld, mask = BatchedImagePlan._replicate(fake, (torch.from_numpy(t), t.shape[1], None), p_in, p_out)
    out, mask = out.numpy(), mask.numpy()
    n = B * p_out
    assert ld == (n + 127) // 128 * 128 and out.shape == (K * K, ld) and len(mask) == ld // 128 + 1
    for b in range(B):
        ref = np.where(t[:, :p_out] >= 0, t[:, :p_out] + b * p_in, -1)
        assert np.array_equal(out[:, b * p_out:(b + 1) * p_out], ref)
        sel = out[:, b * p_out:(b + 1) * p_out]
        assert sel.max() < (b + 1) * p_in and (sel[sel >= 0].min() >= b * p_in)
    assert (out[:, n:] == -1).all()
    for tile in range(ld // 128):
        bits = 0
        for k in range(K * K):
            if (out[k, tile * 128:(tile + 1) * 128] >= 0).any():
                bits |= 1 << k
        assert int(mask[tile]) == bits
    assert int(mask[-1]) == 0
