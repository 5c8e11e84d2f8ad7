"""CPU: the oracle (oracle/) against the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py, run in the build container where /root/reference exists)."""
import json
import os
import sys

import numpy as np
import torch

from oracle import imfnet_oracle, sparse_ops


def rel_rows(a, b):
    return float((torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1)).max())


def test_manifest_records_reference_agreement(golden_dir):
    man = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))
    assert man["state_dict"] == {"entries": 361, "elements": 31461153}
    assert man["c1_real"]["oracle_vs_reference_rel"] < 1e-6
    assert man["batch2"]["oracle_vs_reference_rel"] < 1e-6
    assert man["quantize"]["full_cloud_max_abs_diff"] == 0.0


def test_quantize_order_pinned_by_head_map(golden_dir):
    """files/3D_head_map.ply == xyz[first-occurrence indices] (prefix of cloud_bin_0 @ 2.5 cm)."""
    g = np.load(os.path.join(golden_dir, "quantize_prefix.npz"))
    xyz = g["xyz"].astype(np.float64)
    idx = sparse_ops.unique_first(np.floor(xyz / float(g["voxel"])).astype(np.int32))
    assert np.array_equal(xyz[idx].astype(np.float32), g["head_map_vertices"])
    assert np.all(np.diff(idx) > 0)


def test_standin_utils_match_sparse_ops(golden_dir):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle", "standin"))
    import MinkowskiEngine as ME
    g = np.load(os.path.join(golden_dir, "quantize_prefix.npz"))
    c = np.floor(g["xyz"].astype(np.float64) / 0.025)
    q, idx = ME.utils.sparse_quantize(c, return_index=True)
    assert q.dtype == np.int32 and np.array_equal(q, c[idx].astype(np.int32))
    bc = ME.utils.batched_coordinates([q[:10], q[10:15]])
    assert bc.dtype == torch.int32 and bc[:, 0].tolist() == [0] * 10 + [1] * 5
    h = ME.utils.fnv_hash_vec(np.array([[1, 2, 3], [1, 2, 3], [3, 2, 1]]))
    assert h[0] == h[1] and h[0] != h[2]


def test_oracle_forward_matches_reference_c1_real(golden_dir, state_dict):
    g = np.load(os.path.join(golden_dir, "c1_real.npz"))
    coords = torch.from_numpy(g["coords"])
    feats = torch.ones((len(coords), 1))
    image = torch.from_numpy(g["image"].astype(np.float32))
    d, acts = imfnet_oracle.forward(state_dict, coords, feats, image, return_intermediates=True)
    ref = torch.from_numpy(g["desc"])
    assert rel_rows(d, ref) < 2e-6
    assert np.array_equal(acts["levels"][8], g["s8_coords"])
    assert torch.allclose(acts["fused"], torch.from_numpy(g["fused"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(torch.linalg.norm(d, dim=1), torch.ones(len(d)), atol=1e-5)


def test_oracle_forward_matches_reference_batch2(golden_dir, state_dict):
    g = np.load(os.path.join(golden_dir, "batch2.npz"))
    d = imfnet_oracle.forward(state_dict, torch.from_numpy(g["coords"]), torch.from_numpy(g["feats"]),
                              torch.from_numpy(g["image"].astype(np.float32)))
    assert rel_rows(d, torch.from_numpy(g["desc"])) < 2e-6


def test_oracle_attention_matches_reference(golden_dir, state_dict):
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    out = imfnet_oracle.attention_fusion(state_dict, torch.from_numpy(g["data"]), torch.from_numpy(g["queries"]))
    assert torch.allclose(out, torch.from_numpy(g["out"]), rtol=1e-5, atol=1e-5)


def test_neighbour_table_brute_force():
    """kernel map against an O(N^2) dictionary search, forward / strided / transposed."""
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(-6, 6, (300, 3)), axis=0)
    C = np.concatenate([np.zeros((len(c), 1), np.int64), c], axis=1).astype(np.int32)
    cm = sparse_ops.CoordinateManager(C)
    cm.stride(1, 2)
    fine, coarse = cm.get(1).C, cm.get(2).C
    look_f = {tuple(r): i for i, r in enumerate(fine.tolist())}
    look_c = {tuple(r): i for i, r in enumerate(coarse.tolist())}
    offs = sparse_ops.kernel_offsets(3)
    assert offs[0].tolist() == [-1, -1, -1] and offs[1].tolist() == [0, -1, -1] and offs[26].tolist() == [1, 1, 1]
    t_same = cm.table(1, 1, 3, False)
    t_down = cm.table(1, 2, 3, False)
    t_up = cm.table(2, 1, 3, True)
    for o, r in enumerate(fine.tolist()):
        for k, off in enumerate(offs.tolist()):
            assert t_same[o, k] == look_f.get((r[0], r[1] + off[0], r[2] + off[1], r[3] + off[2]), -1)
            assert t_up[o, k] == look_c.get((r[0], r[1] - off[0], r[2] - off[1], r[3] - off[2]), -1)
    for o, r in enumerate(coarse.tolist()):
        for k, off in enumerate(offs.tolist()):
            assert t_down[o, k] == look_f.get((r[0], r[1] + off[0], r[2] + off[1], r[3] + off[2]), -1)
    # the transposed table is the strided forward table with in/out swapped
    pairs_down = {(int(t_down[o, k]), o, k) for o in range(len(coarse)) for k in range(27) if t_down[o, k] >= 0}
    pairs_up = {(f, int(t_up[f, k]), k) for f in range(len(fine)) for k in range(27) if t_up[f, k] >= 0}
    assert pairs_down == pairs_up


def test_stride_floor_for_negative_coordinates():
    C = np.array([[0, -1, -2, -3], [0, -4, 3, 0], [0, -1, -1, -4]], dtype=np.int32)
    out = sparse_ops.stride_coords(C, 2)
    assert out.tolist() == [[0, -2, -2, -4], [0, -4, 2, 0]]
