"""Descriptor matching: GPU nearest-neighbour search vs the CPU oracle, indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import matching_oracle


def _descs(n, seed, c=32):
    rng = np.random.default_rng(seed)
    d = rng.normal(0, 1, (n, c)).astype(np.float32)
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def test_oracle_agrees_with_ckdtree():
    from scipy.spatial import cKDTree
    a, b = _descs(700, 1), _descs(900, 2)
    assert np.array_equal(matching_oracle.knn_search(a, b), cKDTree(b.astype(np.float64)).query(a.astype(np.float64), k=1)[1])


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb,c", [(5000, 5000, 32), (1, 7, 32), (333, 64, 32), (65, 1000, 16), (257, 129, 64)])
def test_nn_search_matches_oracle(na, nb, c):
    from imfnet_b200.matching import find_nn_gpu, nn_search
    a, b = _descs(na, 3, c), _descs(nb, 4, c)
    ref = matching_oracle.knn_search(a, b)
    idx, d2 = nn_search(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), return_distance=True)
    assert np.array_equal(idx.cpu().numpy(), ref)
    exp = ((a.astype(np.float64) - b[ref].astype(np.float64)) ** 2).sum(1)
    assert np.allclose(d2.cpu().numpy(), exp, rtol=1e-5, atol=1e-7)
    inds, dists = find_nn_gpu(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), nn_max_n=250, return_distance=True)
    assert inds.dtype == torch.int64 and dists.shape == (na, 1) and np.array_equal(inds.numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb,c", [(5000, 5000, 32), (1, 7, 32), (333, 64, 32), (65, 1000, 16), (129, 128, 32)])
def test_tensor_core_search_is_bit_identical_to_brute_force(na, nb, c):
    """The tcgen05 distance product only filters candidates; the survivors are re-evaluated like the brute-force kernel: same indices
    AND the same distance bits, also on near-ties (clusters of almost identical descriptors) and exact ties (duplicated rows)."""
    from imfnet_b200.matching import nn_search
    rng = np.random.default_rng(na + nb)
    a, b = _descs(na, 3, c), _descs(nb, 4, c)
    if nb >= 64:          # near-ties and exact ties among the candidates
        b[nb // 2:nb // 2 + 16] = b[:16] + rng.normal(0, 1e-7, (16, c)).astype(np.float32)
        b[nb // 2 + 16:nb // 2 + 24] = b[:8]
        a[: min(na, 24)] = b[: min(na, 24)] + rng.normal(0, 1e-3, (min(na, 24), c)).astype(np.float32)
    A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    i_tc, d_tc = nn_search(A, B, return_distance=True, tensor_cores=True)
    i_bf, d_bf = nn_search(A, B, return_distance=True, tensor_cores=False)
    assert torch.equal(i_tc, i_bf) and torch.equal(d_tc, d_bf)
    # un-normalised descriptors of very different magnitudes (the filter margin scales with the norms)
    A2, B2 = A * torch.logspace(-2, 2, na, device="cuda")[:, None], B * 30.0
    i_tc, d_tc = nn_search(A2, B2, return_distance=True, tensor_cores=True)
    i_bf, d_bf = nn_search(A2, B2, return_distance=True, tensor_cores=False)
    assert torch.equal(i_tc, i_bf) and torch.equal(d_tc, d_bf)


@pytest.mark.gpu
def test_mutual_nn_5000_keypoints_and_edge_cases():
    """BASELINE config 3 shape: 5000 keypoints per fragment; fragment 2 = noisy permuted copy of half of fragment 1 + outliers."""
    from imfnet_b200.matching import mutual_nn, nn_search
    rng = np.random.default_rng(5)
    d1 = _descs(5000, 6)
    perm = rng.permutation(5000)[:2500]
    d2 = np.concatenate([d1[perm] + rng.normal(0, 0.02, (2500, 32)).astype(np.float32), _descs(2500, 7)], 0)
    nn21, m = mutual_nn(d1, d2)
    rnn21, rm = matching_oracle.mutual(d1, d2)
    assert np.array_equal(nn21, rnn21) and np.array_equal(m, rm)
    assert (nn21[:2500] == perm).mean() > 0.99
    # ties: identical rows -> the first index wins; empty sets
    b = np.repeat(_descs(4, 8), 3, axis=0)
    idx = nn_search(torch.from_numpy(b[[0, 3, 6, 9]]).cuda(), torch.from_numpy(b).cuda())
    assert idx.cpu().tolist() == [0, 3, 6, 9]
    assert nn_search(torch.zeros((0, 32)).cuda(), torch.from_numpy(b).cuda()).numel() == 0
    assert nn_search(torch.from_numpy(b).cuda(), torch.zeros((0, 32)).cuda()).cpu().tolist() == [-1] * 12


@pytest.mark.gpu
def test_describe_and_match_pairs_on_a_fragment_and_its_permuted_copy():
    """A fragment paired with a row-permuted copy of itself: every sampled keypoint that exists on both sides must match its twin
    mutually (identical descriptors up to fp32 summation order)."""
    import imfnet_b200.me as ME
    from imfnet_b200 import load_model, synthetic
    from imfnet_b200.pipeline import describe_and_match_pairs
    model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    model.load_state_dict(synthetic.make_state_dict(0))
    model = model.eval().cuda()
    coords, _ = synthetic.make_fragment(4000, 0.05, seed=8)
    perm = np.random.default_rng(1).permutation(len(coords))
    image = synthetic.make_image(160, 120, seed=8).cuda()
    a = (ME.SparseTensor(torch.ones((4000, 1)), coordinates=torch.from_numpy(coords), device="cuda"), image)
    b = (ME.SparseTensor(torch.ones((4000, 1)), coordinates=torch.from_numpy(coords[perm].copy()), device="cuda"), image)
    (r,) = describe_and_match_pairs(model, [(a, b)], num_keypoints=4000, seed=0)     # all rows are keypoints
    ki, kj = r["kpts_i"].cpu().numpy(), r["kpts_j"].cpu().numpy()
    nn21 = r["nn21"].cpu().numpy()
    # keypoint t of b is voxel coords[perm[kj[t]]]; its match in a must be the same voxel: ki[nn21[t]] == perm[kj[t]]
    assert (ki[nn21] == perm[kj]).mean() > 0.999
    assert len(r["mutual"]) > 0.999 * 4000
