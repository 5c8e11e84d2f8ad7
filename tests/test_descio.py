"""Descriptor file format: what scripts/generate_desc.py:118-123 writes, read back the way
scripts/evaluation_3dmatch.py:129-132 does."""
import os

import numpy as np
import pytest
import torch

from imfnet_b200.descio import DescriptorWriter, load_descriptors, save_descriptors


def _frag(seed, n=1000, u=300):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n, 3)), rng.normal(size=(u, 3)), rng.normal(size=(u, 32)).astype(np.float32)


def test_blocking_writer_is_read_by_the_reference_loader_code(tmp_path):
    points, xyz, feat = _frag(0)
    path = os.path.join(tmp_path, "cloud_bin_0.npz")
    save_descriptors(path, points, xyz, torch.from_numpy(feat))
    data_i = np.load(path)                                            # evaluation_3dmatch.py:129-130, verbatim access pattern
    coord_i, points_i, feat_i = data_i['xyz'], data_i['points'], data_i['feature']
    assert np.array_equal(coord_i, xyz) and np.array_equal(points_i, points) and np.array_equal(feat_i, feat)
    assert feat_i.dtype == np.float32 and points_i.dtype == np.float64
    ref = os.path.join(tmp_path, "ref.npz")                           # generate_desc.py:118-123, verbatim call
    np.savez_compressed(ref, points=np.array(points), xyz=xyz, feature=torch.from_numpy(feat).detach().cpu().numpy())
    a, b = np.load(path), np.load(ref)
    assert sorted(a.files) == sorted(b.files) == ["feature", "points", "xyz"]
    assert all(np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype for k in a.files)
    p2, x2, f2 = load_descriptors(path[:-4])                          # benchmark_util.py:67-71 passes the name without suffix
    assert np.array_equal(f2, feat) and np.array_equal(x2, xyz) and np.array_equal(p2, points)


def test_async_writer_cpu_tensors(tmp_path):
    frags = [_frag(s) for s in range(5)]
    with DescriptorWriter(workers=2) as w:
        for i, (p, x, f) in enumerate(frags):
            w.submit(os.path.join(tmp_path, "scene", "seq-01", f"cloud_bin_{i}.npz"), p, x, torch.from_numpy(f))
    for i, (p, x, f) in enumerate(frags):
        p2, x2, f2 = load_descriptors(os.path.join(tmp_path, "scene", "seq-01", f"cloud_bin_{i}"))
        assert np.array_equal(p2, p) and np.array_equal(x2, x) and np.array_equal(f2, f)


@pytest.mark.gpu
def test_async_writer_gpu_tensors(tmp_path):
    frags = [_frag(s, u=5000) for s in range(4)]
    with DescriptorWriter(workers=2, device="cuda:0") as w:
        for i, (p, x, f) in enumerate(frags):
            g = torch.from_numpy(f).cuda()
            w.submit(os.path.join(tmp_path, f"f{i}.npz"), p, x, g)      # g is dropped here: the writer keeps it alive until copied
    for i, (p, x, f) in enumerate(frags):
        assert np.array_equal(load_descriptors(os.path.join(tmp_path, f"f{i}"))[2], f)
