"""Caller-side pipeline with the reference's `extract_features` signature (/root/reference/util/misc.py:21-104),
plus fragment sharding over the GPUs of one box."""
from __future__ import annotations

import numpy as np
import torch

from . import me as ME
from .voxelize import voxelize


def extract_features(model, xyz, rgb=None, normal=None, voxel_size=0.05, device=None, skip_check=False, is_eval=True,
                     image=None):
    """xyz [N,3] -> (xyz of the kept points [U,3], descriptors [U,C] on the device).

    Same steps as util/misc.py:44-104; the voxelisation (floor, first-occurrence unique, batching) runs on the GPU
    and gives bit-identical indices to np.floor + ME.utils.sparse_quantize(return_index=True)."""
    if is_eval:
        model.eval()
    xyz = np.asarray(xyz)
    if not skip_check:
        assert xyz.shape[1] == 3
        n = xyz.shape[0]
        if rgb is not None:
            assert n == len(rgb) and rgb.shape[1] == 3
            if np.any(rgb > 1):
                raise ValueError('Invalid color. Color must range from [0, 1]')
        if normal is not None:
            assert n == len(normal) and normal.shape[1] == 3
            if np.any(normal > 1):
                raise ValueError('Invalid normal. Normal must range from [-1, 1]')
    if device is None:
        device = torch.device("cuda:0")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("imfnet_b200 runs on CUDA devices only (no CPU fallback)")

    feats = []
    if rgb is not None:
        feats.append(np.asarray(rgb) - 0.5)
    if normal is not None:
        feats.append(np.asarray(normal) / 2)
    if rgb is None and normal is None:
        feats.append(np.ones((len(xyz), 1)))
    feats = np.hstack(feats)

    pts = torch.as_tensor(xyz).to(device, non_blocking=True)          # quantised in the cloud's own dtype, as np.floor(xyz / voxel_size) is
    coords, inds = voxelize(pts, voxel_size, batch_index=0)          # util/misc.py:82-86 on the GPU
    inds_host = inds.cpu().numpy()
    return_coords = xyz[inds_host]
    f = torch.as_tensor(feats[inds_host], dtype=torch.float32).to(device, non_blocking=True)
    stensor = ME.SparseTensor(f, coordinates=coords, device=device)
    image = torch.as_tensor(image, dtype=torch.float32, device=device)
    return return_coords, model(stensor, image).F


def shard_indices(num_items: int, rank: int, world_size: int, weights=None):
    """Fragment indices handled by `rank`.  Fragments are independent (scripts/generate_desc.py:65-123), so the only
    partitioning decision is balance: round-robin, or greedy longest-first when per-fragment weights are given."""
    if weights is None:
        return list(range(rank, num_items, world_size))
    order = sorted(range(num_items), key=lambda i: (-weights[i], i))
    loads, mine = [0.0] * world_size, []
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        loads[r] += weights[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def gather_records(rec: torch.Tensor, world_size: int) -> torch.Tensor:
    """The only collective of this path: all-gather of one small per-rank record (e.g. [rank, voxels, ms])
    -> [world_size, k] on every rank.  Works on whatever backend the default process group uses (NCCL on GPUs, gloo in tests)."""
    if world_size == 1:
        return rec.reshape(1, -1).clone()
    import torch.distributed as dist
    out = [torch.zeros_like(rec) for _ in range(world_size)]
    dist.all_gather(out, rec)
    return torch.stack(out, dim=0)


def aggregate_throughput(records: torch.Tensor):
    """records [world, 3] = (rank, voxels processed, device milliseconds) -> (whole-job voxels/s, slowest rank's ms):
    the job's time is the MAX over ranks, its work the SUM."""
    ms = float(records[:, 2].max())
    return float(records[:, 1].sum()) / (ms * 1e-3), ms


def mutual_from_nn(nn12: torch.Tensor, nn21: torch.Tensor) -> torch.Tensor:
    """Rows j of fragment 2 whose nearest neighbour i = nn21[j] in fragment 1 has j as ITS nearest neighbour (nn12[i] == j).
    A row without a match (index -1: empty other side, or a NaN descriptor that no distance comparison selects) is never mutual --
    it must not be used as an index (nn12[-1] would silently wrap to the last row)."""
    ok = nn21 >= 0
    back = nn12[nn21.clamp_min(0).long()] if len(nn12) else torch.full_like(nn21, -1)
    j = torch.arange(len(nn21), device=nn21.device, dtype=nn21.dtype)
    return torch.nonzero(ok & (back == j)).flatten()


@torch.no_grad()
def describe_and_match_pairs(model, pairs, num_keypoints: int = 5000, seed: int = 0, streams: int = 4):
    """BASELINE config 3 ("fragment pairs ... descriptors + 5000-keypoint L2 feature matching"): for every pair
    ((x_i, image_i), (x_j, image_j)) extract descriptors of both fragments (independent fragments, `forward_many`), draw
    `num_keypoints` random rows of each (scripts/evaluation_3dmatch.py:154-174 samples 5000 keypoints per fragment) and match
    them by mutual nearest neighbour in descriptor space (evaluation_3dmatch.py:207-217).
    Returns one dict per pair: keypoint rows of both fragments, nn21 (for each keypoint of j its match in i) and the mutual
    subset (indices into fragment j's keypoints)."""
    from .matching import nn_search
    flat = [f for pair in pairs for f in pair]
    outs = model.forward_many(flat, streams=streams)
    g = torch.Generator().manual_seed(seed)
    results = []
    for p in range(len(pairs)):
        Fi, Fj = outs[2 * p].F, outs[2 * p + 1].F
        ki = torch.randperm(len(Fi), generator=g)[:num_keypoints].to(Fi.device)
        kj = torch.randperm(len(Fj), generator=g)[:num_keypoints].to(Fj.device)
        di, dj = Fi[ki], Fj[kj]
        nn21 = nn_search(dj, di)
        nn12 = nn_search(di, dj)
        mutual = mutual_from_nn(nn12, nn21)
        results.append({"kpts_i": ki, "kpts_j": kj, "nn21": nn21, "mutual": mutual})
    return results
