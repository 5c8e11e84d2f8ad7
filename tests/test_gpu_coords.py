"""GPU: coordinate kernels vs the oracle -- integer work, bit-exact (hash, unique-first, stride maps, neighbour tables)."""
import os

import numpy as np
import pytest
import torch

from imfnet_b200 import synthetic
from oracle import sparse_ops

pytestmark = pytest.mark.gpu


def _cm(coords_np):
    from imfnet_b200.sparse import CoordinateManager
    return CoordinateManager(torch.from_numpy(coords_np).cuda())


@pytest.mark.parametrize("target,voxel,seed", [(5000, 0.05, 0), (20000, 0.025, 1)])
def test_pyramid_and_tables_bit_exact(target, voxel, seed):
    coords, _ = synthetic.make_fragment(target, voxel, seed)
    coords[:, 1:] -= np.array([40, 25, 10], dtype=np.int32)          # exercise negative coordinates / floor division
    ocm = sparse_ops.CoordinateManager(coords)
    for t in (1, 2, 4):
        ocm.stride(t, 2)
    cm = _cm(coords)
    cm.build_pyramid([2, 4, 8])
    for t in (1, 2, 4, 8):
        assert cm.level(t).n == len(ocm.get(t))
        assert np.array_equal(cm.level(t).coords.cpu().numpy(), ocm.get(t).C), f"coordinates differ at stride {t}"
    for t in (1, 2, 4, 8):
        assert np.array_equal(cm.table(t, t, 3, False).cpu().numpy(), ocm.table(t, t, 3, False))
    for a, b in ((1, 2), (2, 4), (4, 8)):
        assert np.array_equal(cm.table(a, b, 3, False).cpu().numpy(), ocm.table(a, b, 3, False))
        assert np.array_equal(cm.table(b, a, 3, True).cpu().numpy(), ocm.table(b, a, 3, True))
    assert np.array_equal(cm.table(1, 1, 5, False).cpu().numpy(), ocm.table(1, 1, 5, False))


def test_unique_first_matches_head_map_golden(golden_dir):
    """The reference's own known answer: files/3D_head_map.ply pins the order of sparse_quantize."""
    from imfnet_b200.voxelize import unique_first, voxelize
    import imfnet_b200.me as ME
    g = np.load(os.path.join(golden_dir, "quantize_prefix.npz"))
    xyz = g["xyz"].astype(np.float64)
    coords, idx = voxelize(torch.from_numpy(xyz).cuda(), float(g["voxel"]))
    idx = idx.cpu().numpy()
    assert np.array_equal(xyz[idx].astype(np.float32), g["head_map_vertices"])
    assert np.array_equal(coords.cpu().numpy()[:, 1:], np.floor(xyz / 0.025).astype(np.int32)[idx])
    q, inds = ME.utils.sparse_quantize(np.floor(xyz / 0.025), return_index=True)      # reference call form, util/misc.py:83
    assert isinstance(q, np.ndarray) and q.dtype == np.int32 and np.array_equal(inds, idx)
    oidx = sparse_ops.unique_first(np.floor(xyz / 0.025).astype(np.int32))
    assert np.array_equal(unique_first(torch.from_numpy(np.floor(xyz / 0.025).astype(np.int32)).cuda()).cpu().numpy(), oidx)


def test_float32_cloud_is_quantised_in_float32_like_numpy():
    """util/misc.py:82 computes np.floor(xyz / voxel_size) in the cloud's dtype: for a float32 cloud that is a float32 division, which
    puts points next to a voxel boundary into other voxels than float64 would.  The GPU path must follow the input dtype."""
    from imfnet_b200.voxelize import voxelize
    rng = np.random.default_rng(5)
    k = rng.integers(-400, 400, size=(200000, 3))
    xyz32 = (k * np.float32(0.025) + rng.normal(0, 1e-7, k.shape)).astype(np.float32)      # points hugging the voxel boundaries
    for voxel in (0.025, 0.05, 0.3):
        ref32 = np.floor(xyz32 / voxel)
        assert ref32.dtype == np.float32
        ref64 = np.floor(xyz32.astype(np.float64) / voxel)
        c, idx = voxelize(torch.from_numpy(xyz32).cuda(), voxel)
        idx = idx.cpu().numpy()
        assert np.array_equal(c.cpu().numpy()[:, 1:], ref32.astype(np.int32)[idx])
        assert np.array_equal(idx, sparse_ops.unique_first(ref32.astype(np.int32)))
        c64, idx64 = voxelize(torch.from_numpy(xyz32.astype(np.float64)).cuda(), voxel)
        assert np.array_equal(c64.cpu().numpy()[:, 1:], ref64.astype(np.int32)[idx64.cpu().numpy()])
        if voxel == 0.025:
            assert (ref32 != ref64).any(), "the test cloud must contain points where the two dtypes disagree"


def test_edge_cases_empty_single_duplicates_range():
    from imfnet_b200.sparse import CoordinateManager
    from imfnet_b200.voxelize import unique_first
    assert len(unique_first(torch.zeros((0, 4), dtype=torch.int32, device="cuda"))) == 0
    one = torch.tensor([[0, 5, -7, 3]], dtype=torch.int32, device="cuda")
    cm = CoordinateManager(one)
    cm.build_pyramid([2, 4, 8])
    assert [cm.level(t).n for t in (1, 2, 4, 8)] == [1, 1, 1, 1]
    assert cm.level(8).coords.cpu().tolist() == [[0, 0, -8, 0]]
    t = cm.table(1, 1, 3, False).cpu().numpy()
    assert t[0, 13] == 0 and (t >= 0).sum() == 1
    dup = torch.tensor([[0, 1, 1, 1], [0, 2, 2, 2], [0, 1, 1, 1]], dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError, match="duplicate"):
        CoordinateManager(dup).build_pyramid([2])
    far = torch.tensor([[0, 40000, 0, 0]], dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError, match="out of range"):
        CoordinateManager(far).build_pyramid([2])
    same = torch.tensor([[3, 3, 3]] * 1000 + [[1, 2, 3]], dtype=torch.int32, device="cuda")
    assert unique_first(same).cpu().tolist() == [0, 1000]


def test_batched_coordinates_and_segments():
    import imfnet_b200.me as ME
    ca, _ = synthetic.make_fragment(1500, 0.05, seed=11)
    cb, _ = synthetic.make_fragment(1100, 0.05, seed=12)
    bc, bf = ME.utils.sparse_collate([ca[:, 1:], cb[:, 1:]], [np.ones((1500, 1), np.float32), np.ones((1100, 1), np.float32)])
    assert bc.shape == (2600, 4) and bc.dtype == torch.int32 and bf.shape == (2600, 1)
    cm = _cm(bc.numpy())
    cm.build_pyramid([2, 4, 8])
    ocm = sparse_ops.CoordinateManager(bc.numpy())
    for t in (1, 2, 4):
        ocm.stride(t, 2)
    c8 = ocm.get(8).C
    assert np.array_equal(cm.level(8).coords.cpu().numpy(), c8)
    n0 = int((c8[:, 0] == 0).sum())
    assert cm.batch_segments(8, 2) == [0, n0, len(c8)]


def test_full_size_properties_c2():
    """BASELINE config 1 size (50 k voxels): size-independent properties instead of a slow oracle."""
    coords, _ = synthetic.make_fragment(50000, 0.025, 0)
    cm = _cm(coords)
    cm.build_pyramid([2, 4, 8])
    t = cm.table(1, 1, 3, False)
    assert torch.equal(t[:, 13], torch.arange(50000, dtype=torch.int32, device="cuda"))      # centre offset = identity
    # symmetry: j is the k-neighbour of i  <=>  i is the (26-k)-neighbour of j
    i = torch.arange(50000, device="cuda").repeat_interleave(27)
    k = torch.arange(27, device="cuda").repeat(50000)
    j = t.reshape(-1).long()
    ok = j >= 0
    assert torch.equal(t[j[ok], 26 - k[ok]].long(), i[ok])
    # every fine voxel has its parent; the transposed table holds exactly the strided pairs
    dn, up = cm.table(1, 2, 3, False), cm.table(2, 1, 3, True)
    assert int((dn >= 0).sum()) == int((up >= 0).sum())
    assert bool(((up >= 0).sum(1) >= 1).all())
    sizes = [cm.level(s).n for s in (1, 2, 4, 8)]
    assert sizes[0] == 50000 and sizes[0] > sizes[1] > sizes[2] > sizes[3] > 0
