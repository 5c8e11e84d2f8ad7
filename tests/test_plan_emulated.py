"""CPU: the captured plans' host orchestration executed end to end on the C-ABI emulator (tests/abi_emulator.py) and compared
with the oracle -- buffer wiring, concatenation windows, BatchNorm folding / weight scaling, image-encoder launch order, per-item
slots of the batched plan.  The CUDA kernels are NOT exercised here (they are checked against the oracle in the `-m gpu` tests)."""
import contextlib

import pytest
import torch

from imfnet_b200 import _lib, load_model, synthetic
from oracle import imfnet_oracle

from abi_emulator import Emulator
from test_plan_plumbing import FakeEvent, FakeStream

TOL = 1e-4


class ReplayByEnqueue:
    """Stands in for the captured CUDA graph: a replay re-runs the plan's launch sequence (on the emulator)."""

    def __init__(self, plan):
        self.plan = plan

    def replay(self):
        self.plan._enqueue()


@pytest.fixture
def emu(monkeypatch):
    from imfnet_b200.engine import GraphPlan
    e = Emulator(_lib.lib())
    cur = FakeStream()
    monkeypatch.setattr(_lib, "lib", lambda: e)
    monkeypatch.setattr(_lib, "require_cuda", lambda t, what: None)
    monkeypatch.setattr(_lib, "cur_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: cur)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)

    def capture(self):
        self.graph = ReplayByEnqueue(self)
        self.launches_per_replay = 0

    monkeypatch.setattr(GraphPlan, "capture", capture)
    return e


def rel_rows(a, b):
    return float((torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1)).max())


@pytest.fixture(scope="module")
def model_and_sd():
    sd = synthetic.make_state_dict(0)
    m = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def fragments(sizes, W=64, H=48):
    out = []
    for i, n in enumerate(sizes):
        c, _ = synthetic.make_fragment(n, 0.05, seed=60 + i)
        out.append((torch.from_numpy(c), torch.ones((len(c), 1)), synthetic.make_image(W, H, seed=60 + i)))
    return out


def test_single_fragment_plan_matches_oracle(emu, model_and_sd):
    m, sd = model_and_sd
    m._plan, m.img_encoder._plans = None, {}          # plans of an earlier test refer to that test's emulator
    c, f, im = fragments([900])[0]
    out = torch.zeros((len(c), 32))
    res = m.forward_many_host([(c, f, im)], streams=1, out=[out])[0]
    ref = imfnet_oracle.forward(sd, c, f, im)
    assert rel_rows(res, ref) < TOL
    assert abs(float(torch.linalg.norm(res, dim=1).mean()) - 1.0) < 1e-5


def test_batched_plan_matches_oracle_per_fragment_and_batched(emu, model_and_sd):
    m, sd = model_and_sd
    m._plan, m.img_encoder._plans = None, {}          # plans of an earlier test refer to that test's emulator
    frags = fragments([900, 600, 750, 820, 500])               # two batches of 2 + a tail of 1 (single-fragment path)
    outs = [torch.zeros((len(c), 32)) for c, _f, _im in frags]
    res = m.forward_batches(frags, batch=2, streams=2, out=outs)
    for (c, f, im), r in zip(frags, res):
        assert rel_rows(r, imfnet_oracle.forward(sd, c, f, im)) < TOL
    # the reference's own batched form (batch index in column 0, images stacked: model/resunet.py:237-273) gives the same rows
    c0, c1 = frags[0][0].clone(), frags[1][0].clone()
    c1[:, 0] = 1
    both = imfnet_oracle.forward(sd, torch.cat([c0, c1]), torch.cat([frags[0][1], frags[1][1]]), torch.cat([frags[0][2], frags[1][2]]))
    assert rel_rows(torch.cat([res[0], res[1]]), both) < TOL


def test_batched_plan_item_overflow_falls_back(emu, model_and_sd, monkeypatch):
    m, sd = model_and_sd
    m._plan, m.img_encoder._plans = None, {}          # plans of an earlier test refer to that test's emulator
    frags = fragments([700, 650])
    monkeypatch.setattr(type(m), "_cap8", staticmethod(lambda rows, scale: 8 * scale))      # far too small per item / per plan
    outs = [torch.zeros((len(c), 32)) for c, _f, _im in frags]
    res = m.forward_batches(frags, batch=2, out=outs)
    for (c, f, im), r in zip(frags, res):
        assert rel_rows(r, imfnet_oracle.forward(sd, c, f, im)) < TOL


def test_eager_plan_batch_of_two_matches_oracle(emu, model_and_sd):
    """forward(x, image) with a collated batch (the reference's training-style input, lib/data_loaders.py:68-69)."""
    import imfnet_b200.me as ME
    m, sd = model_and_sd
    m._plan, m.img_encoder._plans = None, {}          # plans of an earlier test refer to that test's emulator
    (c0, f0, i0), (c1, f1, i1) = fragments([800, 700])
    c1 = c1.clone()
    c1[:, 0] = 1
    coords, feats, images = torch.cat([c0, c1]), torch.cat([f0, f1]), torch.cat([i0, i1])
    out = m(ME.SparseTensor(feats, coordinates=coords), images)
    assert torch.equal(out.C, coords)
    assert rel_rows(out.F, imfnet_oracle.forward(sd, coords, feats, images)) < TOL


@pytest.mark.parametrize("name", ["ResUNetBN2E", "ResUNetBN2B"])
def test_other_channel_configurations_match_oracle(emu, name):
    """Chunk widths and concatenation windows depend on CHANNELS / TR_CHANNELS (model/resunet.py:276-326)."""
    sd = synthetic.make_state_dict(2, name)
    m = load_model(name)(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    m.load_state_dict(sd, strict=True)
    m.eval()
    c, f, im = fragments([700])[0]
    out = torch.zeros((len(c), 32))
    res = m.forward_many_host([(c, f, im)], streams=1, out=[out])[0]
    assert rel_rows(res, imfnet_oracle.forward(sd, c, f, im)) < TOL
