"""CPU: host plumbing of the captured plans (engine.GraphPlan, batched.BatchGraphPlan) without a GPU.

The launch sequences are pure host code around C-ABI calls, so they can be executed here against a recording stand-in of the
library: every call is checked against the ctypes signature table (argument count and convertibility -- what a real call would
raise on), the pure-host entry points (sizes, capacities) go to the real library, and torch.cuda's stream / graph objects are
replaced by inert ones.  Nothing is computed: this catches wrong attribute names, argument lists, slice arithmetic and
buffer shapes -- not numerics (those are the `-m gpu` tests)."""
import contextlib
import ctypes

import pytest
import torch

from imfnet_b200 import _lib, load_model, synthetic

HOST_ONLY = {"imf_last_error", "imf_version", "imf_launch_count", "imf_hash_capacity", "imf_hash_bytes", "imf_device_sm_count", "imf_conv_first_tc_columns"}


class FakeLib:
    def __init__(self, real):
        self.real, self.calls = real, []

    def __getattr__(self, name):
        if name not in _lib.SIGNATURES:
            raise AttributeError(name)
        if name in HOST_ONLY or name.endswith("_bytes"):
            return getattr(self.real, name)
        res, argtypes = _lib.SIGNATURES[name]

        def call(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments, the ABI takes {len(argtypes)}"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    if t in (ctypes.c_int32, ctypes.c_longlong, ctypes.c_size_t):
                        assert isinstance(a, int) and not isinstance(a, bool), a
                        t(a)
                        if t is ctypes.c_int32:
                            assert -2 ** 31 <= a < 2 ** 31, a
                    elif t in (ctypes.c_float, ctypes.c_double):
                        t(float(a))
                    else:
                        t.from_param(a)
                except Exception as e:      # noqa: BLE001
                    raise AssertionError(f"{name}: argument {i} = {a!r} does not convert to {t}") from e
            self.calls.append((name, args))
            return 0

        return call


class FakeStream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def synchronize(self):
        pass


class FakeEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, s=None):
        pass

    def synchronize(self):
        pass


class FakeGraph:
    replays = 0

    def replay(self):
        FakeGraph.replays += 1


@pytest.fixture
def fake_cuda(monkeypatch):
    fake = FakeLib(_lib.lib())
    cur = FakeStream()
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "cur_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: cur)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "graph", lambda g, *a, **k: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    FakeGraph.replays = 0
    return fake


def make_model():
    m = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    m.load_state_dict(synthetic.make_state_dict(0), strict=True)
    return m.eval()


def fragments(sizes, W=64, H=48):
    out = []
    for i, n in enumerate(sizes):
        c, _ = synthetic.make_fragment(n, 0.05, seed=40 + i)
        out.append((torch.from_numpy(c), torch.ones((len(c), 1)), synthetic.make_image(W, H, seed=40 + i)))
    return out


def names(fake):
    return [c[0] for c in fake.calls]


def test_single_fragment_graph_plan_sequence(fake_cuda):
    """The verified single-fragment path through the harness: one capture (3 enqueues), one replay, the expected launch list."""
    m = make_model()
    c, f, im = fragments([700])[0]
    out = torch.empty((len(c), 32))
    res = m.forward_many_host([(c, f, im)], streams=1, out=[out])
    assert res[0].shape == (len(c), 32) and FakeGraph.replays == 1
    seq = names(fake_cuda)
    per_enqueue = seq.count("imf_hash_build")
    assert per_enqueue == 3                                                  # two warm-ups + the capture
    assert seq.count("imf_sparse_conv_g4_fwd_perm") == 3 * 20               # 6 strided / transposed + 14 block convolutions
    assert seq.count("imf_sparse_conv_g4_fwd") == 3 * 9                     # image encoder: layer2 (the stem and layer1 have their own kernels)
    assert seq.count("imf_image_conv3x3_p8_fwd") == 3 * 6 and seq.count("imf_image_maxpool_p8") == 3          # layer1 on the plane layout
    assert seq.count("imf_image_stem_h2_fwd") == 3 and seq.count("imf_image_im2col_h2_batch") == 0
    assert seq.count("imf_conv_first_tc_h2_fwd_keep") == 3 and seq.count("imf_conv_first_h2_fwd") == 0      # conv1 on the tensor-core path
    # ... whose dense grid the neighbour tables read before it is released: conv1, tables, release, in that order, per enqueue
    order = [n for n in seq if n in ("imf_conv_first_tc_h2_fwd_keep", "imf_kernel_map_t_batch", "imf_conv_first_tc_release")]
    assert order == ["imf_conv_first_tc_h2_fwd_keep", "imf_kernel_map_t_batch", "imf_conv_first_tc_release"] * 3
    # the fusion module of the single-fragment plan is the batched chain with B = 1
    assert seq.count("imf_attention_fusion_fwd_batched") == 3 == seq.count("imf_attention_kv_batched") == seq.count("imf_batch_segments_n")
    # tail: conv1_tr -> ReLU -> final -> L2 norm as one fused tensor-core kernel
    assert seq.count("imf_tail_fused_h2_fwd") == 3 and seq.count("imf_h2_unpack_l2norm") == 0 == seq.count("imf_pointwise_tail_h2_fwd")


def test_batched_plan_sequence_and_slices(fake_cuda):
    from imfnet_b200.batched import BatchGraphPlan
    m = make_model()
    frags = fragments([700, 900, 650, 800, 500])                             # two batches of 2 + a tail of 1
    outs = [torch.full((len(c), 32), float("nan")) for c, _f, _im in frags]
    res = m.forward_batches(frags, batch=2, streams=2, out=outs)
    assert [tuple(r.shape) for r in res] == [(len(c), 32) for c, _f, _im in frags]
    assert all(r.data_ptr() == o.data_ptr() for r, o in zip(res, outs))      # results land in the caller's buffers
    plans = [g for k, pool in m._graphs.items() if k[0] == "batch" for g in pool]
    assert len(plans) == 2 and all(isinstance(g, BatchGraphPlan) for g in plans)
    g = plans[0]
    n0, n1 = len(frags[0][0]), len(frags[1][0])
    # coordinates: rows of item b carry batch index b, the other columns are the fragment's
    assert torch.equal(g.coords[1][:n0, 1:], frags[0][0][:, 1:]) and torch.equal(g.coords[1][n0:n0 + n1, 1:], frags[1][0][:, 1:])
    assert int(g.coords[1][:n0, 0].max()) == 0 and int(g.coords[1][n0:n0 + n1, 0].min()) == 1 == int(g.coords[1][n0:n0 + n1, 0].max())
    assert int(g.n1[0]) == n0 + n1 and g.image.shape == (2, 3, 48, 64)
    assert torch.equal(g.image[1], frags[1][2].reshape(3, 48, 64))
    seq = names(fake_cuda)
    # per enqueue of a batched plan: one image-encoder pass, one segment kernel, ONE fusion chain over all items
    # (the trailing single fragment goes through a single-fragment plan: 3 more enqueues of the same chain with B = 1)
    enq = len([a for n, a in fake_cuda.calls if n == "imf_batch_segments_n" and a[3] == 2])
    assert enq == 2 * 3                                                       # two plans x (two warm-ups + capture)
    assert len([a for n, a in fake_cuda.calls if n == "imf_attention_fusion_fwd_batched" and a[8] == 2]) == enq
    assert len([a for n, a in fake_cuda.calls if n == "imf_attention_kv_batched" and a[4] == 2]) == enq
    assert len([a for n, a in fake_cuda.calls if n == "imf_image_stem_h2_fwd" and a[3] == 2]) == enq          # both images, one fused stem launch
    assert len([a for n, a in fake_cuda.calls if n == "imf_image_maxpool_p8" and a[9] == 2]) == enq
    # the image encoder's batched launches cover B * P rows
    ip = g.image_plan
    layer1 = [a for n, a in fake_cuda.calls if n == "imf_image_conv3x3_p8_fwd" and a[3] == 2]
    assert len(layer1) == 6 * enq, "layer1: six convolutions over both images per enqueue"
    assert len([a for n, a in fake_cuda.calls if n == "imf_sparse_conv_g4_fwd" and a[8] == 2 * ip.P2]) == 9 * enq          # layer2 over both images' rows
    assert ip.s0.shape[0] == 2 * ip.P0 and ip.tokens.shape == (2 * ip.P2, 128)
    # fusion chain: the level's rows through one call with the plan's segment / count arrays and the stride-8 row count on the device
    att = [a for n, a in fake_cuda.calls if n == "imf_attention_fusion_fwd_batched" and a[2] == g.P8.data_ptr()][-1]
    assert att[4] == g.cap8 == 2 * g.item_cap8 and att[5] == g._n(8) and att[6] == g.seg.data_ptr() and att[7] == g.cnt.data_ptr()
    assert att[9] == g.kv.data_ptr() and att[10] == ip.P2 and att[11] == g.fused32.data_ptr()          # image tokens per item
    kvc = [a for n, a in fake_cuda.calls if n == "imf_attention_kv_batched" and a[5] == g.kv.data_ptr()][-1]
    assert kvc[2] == ip.tokens.data_ptr() and kvc[3] == ip.P2 and kvc[4] == 2


def test_streaming_calls_leave_groups_in_flight_until_drained(fake_cuda):
    """forward_batches(..., carry=state): a call does not wait for its last groups (the next call's copies overlap them); a plan is
    only reused after its previous group retired; drain_batches fills in what is left, in the lists the calls returned."""
    m = make_model()
    frags = fragments([700, 900, 650, 800])                                  # two groups of 2
    carry = {}
    o1, o2, o3 = ([torch.full((len(c), 32), float("nan")) for c, _f, _im in frags] for _ in range(3))
    r1 = m.forward_batches(frags, batch=2, streams=2, out=o1, carry=carry)
    assert r1 == [None] * 4 and len(carry["inflight"]) == 2                  # both groups still in flight
    r2 = m.forward_batches(frags, batch=2, streams=2, out=o2, carry=carry)   # same plans: the first call's groups are retired one by one
    assert all(r is not None and tuple(r.shape) == (len(c), 32) for r, (c, _f, _im) in zip(r1, frags))
    assert r2 == [None] * 4 and len(carry["inflight"]) == 2
    m.drain_batches(carry)
    assert not carry["inflight"] and all(tuple(r.shape) == (len(c), 32) for r, (c, _f, _im) in zip(r2, frags))
    # a call without carry behaves as before (everything retired on return)
    r3 = m.forward_batches(frags, batch=2, streams=2, out=o3)
    assert all(r is not None for r in r3)


def test_batched_plan_errors(fake_cuda):
    from imfnet_b200.batched import BatchGraphPlan
    from imfnet_b200.engine import PlanCapacityError
    m = make_model()
    m._ensure_plan()
    g = BatchGraphPlan(m._plan, 4096, 48, 64, 2, 512)
    frags = fragments([3000, 3000])
    with pytest.raises(PlanCapacityError):
        g.launch_batch(frags)
    with pytest.raises(ValueError):
        g.launch_batch(frags[:1])
    small = fragments([300, 400])
    g.launch_batch(small)
    g.meta_host[16] = BatchGraphPlan.ERR_ITEM_CAPACITY
    with pytest.raises(PlanCapacityError):
        g.finish_batch()
    g.launch_batch(small)
    g.meta_host[16] = BatchGraphPlan.ERR_BATCH_INDEX
    with pytest.raises(ValueError):
        g.finish_batch()
    g.launch_batch(small)
    outs = g.finish_batch()
    assert [len(o) for o in outs] == [len(small[0][0]), len(small[1][0])]
