"""GPU: the batched captured plan (imfnet_b200/batched.py) -- the execution mode bench.py measures -- against the CPU oracle,
against the golden outputs of the unmodified reference, and against forward() fragment by fragment (expected bit-identical: same
kernels, same per-row summation order)."""
import os

import numpy as np
import pytest
import torch

from imfnet_b200 import synthetic

pytestmark = [pytest.mark.gpu]
TOL = 1e-4


def rel_rows(a, b):
    return float((torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1)).max())


def fragments(sizes, W, H, voxel=0.05):
    out = []
    for i, n in enumerate(sizes):
        c, _ = synthetic.make_fragment(n, voxel, seed=20 + i)
        out.append((torch.from_numpy(c), torch.ones((len(c), 1)), synthetic.make_image(W, H, seed=20 + i)))
    return out


@pytest.mark.parametrize("on_device", [True, False])
def test_batched_plan_matches_single_forwards(cuda_model, on_device):
    import imfnet_b200.me as ME
    frags = fragments([3000, 5000, 4100, 2500, 3000], 160, 120)          # ragged sizes; 4 in batches of 2 + a tail of 1
    singles = [cuda_model(ME.SparseTensor(f.cuda(), coordinates=c.cuda()), im.cuda()).F.cpu() for c, f, im in frags]
    if on_device:
        inp = [(c.cuda(), f.cuda(), im.cuda()) for c, f, im in frags]
    else:
        inp = [(c.pin_memory(), f.pin_memory(), im.pin_memory()) for c, f, im in frags]
    outs = cuda_model.forward_batches(inp, batch=2, streams=2)
    assert len(outs) == len(frags)
    worst = 0.0
    for o, s in zip(outs, singles):
        assert o.shape == s.shape and o.is_cuda == on_device
        worst = max(worst, rel_rows(o.cpu(), s))
    print(f"batched vs single: worst row-wise rel diff {worst:.3e}")
    assert worst < TOL
    # second call re-uses the captured plans
    outs2 = cuda_model.forward_batches(inp, batch=2, streams=2)
    for o, o2 in zip(outs, outs2):
        assert torch.equal(o.cpu(), o2.cpu())


def test_batched_plan_rejects_wrong_batch_index_capacity(cuda_model):
    """An item with more stride-8 rows than the plan's per-item capacity must fall back, not truncate."""
    frags = fragments([3000, 3000], 160, 120)
    import imfnet_b200.me as ME
    singles = [cuda_model(ME.SparseTensor(f.cuda(), coordinates=c.cuda()), im.cuda()).F.cpu() for c, f, im in frags]
    cls = type(cuda_model)
    assert "_cap8" not in cls.__dict__          # inherited from ResUNet2: shadow it on the subclass, remove the shadow afterwards
    try:
        cls._cap8 = staticmethod(lambda rows, scale: 16)          # absurdly small per-item capacity
        cuda_model._graphs.clear()
        outs = cuda_model.forward_batches([(c.cuda(), f.cuda(), im.cuda()) for c, f, im in frags], batch=2)
    finally:
        del cls._cap8
        cuda_model._graphs.clear()
        cuda_model._cap8_scale.clear()
    for o, s in zip(outs, singles):
        assert rel_rows(o.cpu(), s) < TOL


def test_segment_kernels_unit():
    """csrc/batched.cu against numpy: segments / clipped counts / error bits, and the per-item h2 <-> fp32 moves."""
    import numpy as np
    from imfnet_b200 import _lib
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    B, cap = 4, 300
    sizes = [257, 0, 300, 123]
    coords = np.zeros((sum(sizes) + 50, 4), dtype=np.int32)
    coords[:sum(sizes), 0] = np.repeat(np.arange(B), sizes)
    coords[sum(sizes):, 0] = 77                                            # rows past n must not be looked at
    coords[:, 1:] = rng.integers(-100, 100, (len(coords), 3))
    c = torch.from_numpy(coords).cuda()
    n_dev = torch.tensor([sum(sizes)], dtype=torch.int32, device="cuda")
    seg = torch.full((B + 1,), -1, dtype=torch.int32, device="cuda")
    cnt = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.imf_batch_segments_n(c.data_ptr(), n_dev.data_ptr(), len(coords), B, cap, seg.data_ptr(), cnt.data_ptr(), err.data_ptr(), s))
    assert seg.tolist() == [0, 257, 257, 557, 680] and cnt.tolist() == sizes and int(err.item()) == 0
    _lib.check(L.imf_batch_segments_n(c.data_ptr(), n_dev.data_ptr(), len(coords), B, 200, seg.data_ptr(), cnt.data_ptr(), err.data_ptr(), s))
    assert cnt.tolist() == [200, 0, 200, 123] and int(err.item()) == 0x20000
    err.zero_()
    _lib.check(L.imf_batch_segments_n(c.data_ptr(), n_dev.data_ptr(), len(coords), 3, cap, seg.data_ptr(), cnt.data_ptr(), err.data_ptr(), s))
    assert int(err.item()) == 0x40000                                      # batch index 3 present, plan built for 3 items
    err.zero_()
    _lib.check(L.imf_batch_segments_n(c.data_ptr(), n_dev.data_ptr(), len(coords), B, cap, seg.data_ptr(), cnt.data_ptr(), err.data_ptr(), s))
    # h2 matrix of the level: pack all rows with the verified kernel, pull item 2 out, push a modified copy back
    C, KC, n = 256, 64, sum(sizes)
    X = torch.randn(n, C, device="cuda")
    H = torch.zeros(n, 2 * C, dtype=torch.float16, device="cuda")
    _lib.check(L.imf_h2_pack(X.data_ptr(), C, n, C, KC, H.data_ptr(), 2 * C, None, s))
    item = torch.full((cap, C), float("nan"), device="cuda")
    _lib.check(L.imf_h2_unpack_seg(H.data_ptr(), 2 * C, seg.data_ptr() + 8, cnt.data_ptr() + 8, cap, C, KC, item.data_ptr(), C, s))
    back = torch.empty(n, C, device="cuda")
    _lib.check(L.imf_h2_unpack(H.data_ptr(), 2 * C, n, C, KC, back.data_ptr(), C, s))
    assert torch.equal(item[:300], back[257:557])
    item2 = item * 2 + 1
    _lib.check(L.imf_h2_pack_seg(item2.data_ptr(), C, seg.data_ptr() + 8, cnt.data_ptr() + 8, cap, C, KC, H.data_ptr(), 2 * C, err.data_ptr(), s))
    after = torch.empty(n, C, device="cuda")
    _lib.check(L.imf_h2_unpack(H.data_ptr(), 2 * C, n, C, KC, after.data_ptr(), C, s))
    assert torch.equal(after[:257], back[:257]) and torch.equal(after[557:], back[557:])
    assert float((after[257:557] - item2[:300]).abs().max()) <= 1e-6 * float(item2[:300].abs().max())
    assert int(err.item()) == 0


# ---- the headline path against the ORACLE and the reference's golden outputs (not only against forward()) ----------------------
def test_batched_plan_c2_batch_of_10_vs_oracle(state_dict, cuda_model):
    """bench.py's step: groups of 10 C2 fragments (50 k voxels + 640x480) per captured-graph replay.  Two fragments of the first
    group and the ragged tail are compared with the CPU oracle (a full oracle forward of a 50 k fragment takes seconds), the rest
    with forward(), which tests/test_gpu_forward.py pins to the oracle at this size."""
    import imfnet_b200.me as ME
    from oracle import imfnet_oracle
    frags = []
    for i in range(12):                                      # 10 in one batched replay + a tail of 2 through forward_many
        c, _ = synthetic.make_fragment(50000 if i != 3 else 41000, 0.025, seed=i)
        frags.append((torch.from_numpy(c), torch.ones((len(c), 1)), synthetic.make_image(640, 480, seed=i)))
    inp = [(c.cuda(), f.cuda(), im.cuda()) for c, f, im in frags]
    outs = cuda_model.forward_batches(inp, batch=10, streams=2)
    assert len(outs) == 12 and all(o.shape == (len(fr[0]), 32) for o, fr in zip(outs, frags))
    worst = 0.0
    for i in (0, 3, 11):
        c, f, im = frags[i]
        ref = imfnet_oracle.forward(state_dict, c, f, im)
        worst = max(worst, rel_rows(outs[i].cpu(), ref))
    print(f"batched B=10 C2 vs oracle: worst row-wise rel err {worst:.3e}")
    assert worst < TOL
    for i in (1, 5, 9):
        c, f, im = inp[i]
        single = cuda_model(ME.SparseTensor(f, coordinates=c), im).F
        assert torch.equal(single, outs[i]), f"fragment {i}: the batched plan must reproduce forward() bit for bit"
    # pinned-host form (the e2e leg of bench.py) gives the same bits
    pin = [(c.pin_memory(), f.pin_memory(), im.pin_memory()) for c, f, im in frags[:10]]
    outs_h = cuda_model.forward_batches(pin, batch=10, streams=2)
    for a, b in zip(outs[:10], outs_h):
        assert not b.is_cuda and torch.equal(a.cpu(), b)


def test_batched_plan_matches_reference_goldens(golden_dir, cuda_model):
    """The goldens of the UNMODIFIED reference (tests/golden, oracle/make_golden.py) through the batched plan: c1_real twice in one
    batch (each item against its own image), and the items of the batch-of-two golden as two fragments."""
    g = np.load(os.path.join(golden_dir, "c1_real.npz"))
    coords = torch.from_numpy(g["coords"]).cuda()
    feats = torch.ones((len(coords), 1), device="cuda")
    image = torch.from_numpy(g["image"].astype(np.float32)).cuda()
    ref = torch.from_numpy(g["desc"])
    outs = cuda_model.forward_batches([(coords, feats, image), (coords.clone(), feats, image)], batch=2)
    for o in outs:
        assert rel_rows(o.cpu(), ref) < TOL
    g = np.load(os.path.join(golden_dir, "batch2.npz"))
    c, f, im, d = torch.from_numpy(g["coords"]), torch.from_numpy(g["feats"]), torch.from_numpy(g["image"].astype(np.float32)), torch.from_numpy(g["desc"])
    items = []
    for b in range(2):
        sel = c[:, 0] == b
        cb = c[sel].clone()
        cb[:, 0] = 0
        items.append((cb.cuda(), f[sel].cuda(), im[b:b + 1].cuda(), d[sel]))
    outs = cuda_model.forward_batches([it[:3] for it in items], batch=2)
    for o, it in zip(outs, items):
        assert rel_rows(o.cpu(), it[3]) < TOL


def test_batched_plan_bit_reproducible_over_100_replays(cuda_model):
    """The convolution kernel's MMA warps pass their turn on before issuing (early hand-off); every accumulator still has one
    issuing thread, so replays must give the same bits.  100 replays of a batch of 4 fragments at 20 k voxels (row mode at level 1,
    one CTA per tile below), compared with the first replay."""
    frags = []
    for i in range(4):
        c, _ = synthetic.make_fragment(20000, 0.025, seed=40 + i)
        frags.append((torch.from_numpy(c).cuda(), torch.ones((len(c), 1), device="cuda"), synthetic.make_image(320, 240, seed=i).cuda()))
    first = [o.clone() for o in cuda_model.forward_batches(frags, batch=4, streams=1)]
    for rep in range(100):
        outs = cuda_model.forward_batches(frags, batch=4, streams=1)
        for a, b in zip(first, outs):
            assert torch.equal(a, b), f"replay {rep} differs"
