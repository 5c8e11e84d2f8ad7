"""The drop-in claim of INTEGRATION.md section 1, exercised with the REFERENCE's own caller code: /root/reference/util/misc.py is
imported unmodified (by file path) after the documented `sys.modules` swap, and its `extract_features` (util/misc.py:21-104) is run
against the imfnet_b200 model.  On the CPU the C-ABI calls go to tests/abi_emulator.py (host logic only); the descriptors are
compared with the oracle.  /root/reference only exists in the build container: elsewhere (the GPU box) these tests skip, and
tests/test_gpu_forward.py::test_reference_call_sequence_on_gpu replays the same call sequence, restated, on the real kernels."""
import functools
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

import imfnet_b200
import imfnet_b200.me
import imfnet_b200.model
from imfnet_b200 import synthetic
from oracle import imfnet_oracle, sparse_ops

from test_plan_emulated import emu, model_and_sd, rel_rows  # noqa: F401  (fixtures)

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "util", "misc.py")), reason="the reference tree is not present")


@pytest.fixture
def reference_misc(monkeypatch):
    """the two sys.modules lines, then the reference's util/misc.py loaded as it is."""
    monkeypatch.setitem(sys.modules, "MinkowskiEngine", imfnet_b200.me)
    monkeypatch.setitem(sys.modules, "model", imfnet_b200.model)
    spec = importlib.util.spec_from_file_location("reference_util_misc", os.path.join(REF, "util", "misc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ME is imfnet_b200.me
    return mod


def test_reference_extract_features_runs_on_the_drop_in(emu, model_and_sd, reference_misc, monkeypatch):  # noqa: F811
    m, sd = model_and_sd
    m._plan, m.img_encoder._plans = None, {}
    # without a GPU the voxel hash of ME.utils.sparse_quantize runs on the emulator too: keep its tensors on the CPU
    orig = imfnet_b200.me.utils.sparse_quantize
    monkeypatch.setattr(imfnet_b200.me.utils, "sparse_quantize", staticmethod(functools.partial(orig, device="cpu")))
    _, pts = synthetic.make_fragment(700, 0.05, seed=77)
    rng = np.random.default_rng(0)
    xyz = np.concatenate([pts + 0.01, pts + 0.02, rng.permutation(pts)[:200] + 0.011])          # duplicates: several points per voxel
    image = synthetic.make_image(64, 48, seed=77).numpy()
    return_coords, F = reference_misc.extract_features(m, xyz, voxel_size=0.05, device=torch.device("cpu"), skip_check=False, image=image)
    q = np.floor(xyz / 0.05)
    idx = sparse_ops.unique_first(q.astype(np.int32))
    assert np.array_equal(return_coords, xyz[idx]), "kept points = first occurrence per voxel, in source order (util/misc.py:83-87)"
    coords = torch.from_numpy(np.concatenate([np.zeros((len(idx), 1)), q[idx]], 1).astype(np.int32))
    ref = imfnet_oracle.forward(sd, coords, torch.ones((len(idx), 1)), torch.from_numpy(image))
    assert F.shape == (len(idx), 32) and rel_rows(F, ref) < 1e-4


def test_reference_model_registry_and_eval_hash_helpers(reference_misc):
    """What scripts/generate_desc.py:165-173 and scripts/evaluation_3dmatch.py:164-168 touch besides forward()."""
    import model as swapped                                     # the sys.modules entry installed by the fixture
    Model = swapped.load_model("ResUNetBN2C")
    net = Model(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    net.load_state_dict(synthetic.make_state_dict(0), strict=True)
    # ME.utils.fnv_hash_vec: the evaluation script intersects keypoints with voxel coordinates through it
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "standin"))
    try:
        standin = importlib.util.spec_from_file_location("standin_me_utils", os.path.join(os.path.dirname(__file__), "..", "oracle", "standin",
                                                                                           "MinkowskiEngine", "utils.py"))
        su = importlib.util.module_from_spec(standin)
        standin.loader.exec_module(su)
    finally:
        sys.path.pop(0)
    rng = np.random.default_rng(1)
    a = np.floor(rng.normal(0, 3, (5000, 3)) / 0.025)
    h = reference_misc.ME.utils.fnv_hash_vec(a)
    assert h.dtype == np.uint64 and np.array_equal(h, su.fnv_hash_vec(a))
    # restated from SURVEY.md Appendix A, one row by hand (python ints, wrap at 2**64)
    row = [int(v) % (1 << 64) for v in a[17]]
    x = 14695981039346656037
    for v in row:
        x = (x * 1099511628211) % (1 << 64)
        x ^= v
    assert int(h[17]) == x
    sub = rng.permutation(len(a))[:1000]
    assert np.array_equal(np.where(np.isin(h, h[sub]))[0], np.sort(np.where(np.isin(h, h[sub]))[0]))
    # the reference's own hash helper in util/misc.py works on our arrays unchanged
    assert reference_misc._hash(a.astype(np.int64), 1000).shape == (5000,)
