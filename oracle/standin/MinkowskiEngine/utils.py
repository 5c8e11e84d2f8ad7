"""CPU ORACLE stand-in for MinkowskiEngine.utils (ME 0.5.4 semantics).

Call sites in the reference: util/misc.py:83,86 (sparse_quantize, batched_coordinates);
lib/data_loaders.py:68-69 (sparse_collate); scripts/evaluation_3dmatch.py:154-174 (fnv_hash_vec)."""
import numpy as np
import torch

from oracle import sparse_ops as _ops


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    is_torch = isinstance(coordinates, torch.Tensor)
    c = coordinates.detach().cpu().numpy() if is_torch else np.asarray(coordinates)
    if quantization_size is not None:
        c = c / quantization_size
    d = np.floor(c).astype(np.int32)
    idx = _ops.unique_first(d)
    out = d[idx]
    if is_torch:
        out, idx_o = torch.from_numpy(out), torch.from_numpy(idx)
    else:
        idx_o = idx
    if return_maps_only:
        return idx_o
    res = [out]
    if features is not None:
        res.append(features[idx])
    if return_index:
        res.append(idx_o)
    return res[0] if len(res) == 1 else tuple(res)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    rows = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c) if not isinstance(c, torch.Tensor) else c)
        if c.is_floating_point():
            c = torch.floor(c)
        c = c.to(dtype)
        rows.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype), c], dim=1))
    out = torch.cat(rows, dim=0) if rows else torch.zeros((0, 4), dtype=dtype)
    return out if device is None else out.to(device)


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    bc = batched_coordinates(coords, dtype=dtype, device=device)
    fs = torch.cat([torch.as_tensor(np.asarray(f) if not isinstance(f, torch.Tensor) else f) for f in feats], dim=0)
    if labels is not None:
        ls = torch.cat([torch.as_tensor(np.asarray(l) if not isinstance(l, torch.Tensor) else l) for l in labels], dim=0)
        return bc, fs, ls
    return bc, fs


def fnv_hash_vec(arr):
    """FNV64-1A over the columns of an integer array."""
    arr = np.asarray(arr).copy().astype(np.uint64, copy=False)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h
