"""GPU voxelisation front end: what util/misc.py:82-86 does on the host with numpy + ME.

  coords = floor(xyz / voxel)            util/misc.py:82  -> imf_quantize_points / _f32 (in the cloud's dtype, as numpy: bit-exact)
  sparse_quantize(..., return_index)     util/misc.py:83  -> imf_stride_map(stride=1, first_idx)
Order of the returned rows = first occurrence, ascending source index (pinned by files/3D_head_map.ply).
"""
from __future__ import annotations

import torch

from . import _lib


def unique_first(coords: torch.Tensor) -> torch.Tensor:
    """coords int32 [N,3] or [N,4] on CUDA -> int32 [U] source rows of the first occurrence of every distinct row."""
    _lib.require_cuda(coords, "coordinates")
    L = _lib.lib()
    c = coords.to(torch.int32)
    if c.shape[1] == 3:
        c = torch.cat([torch.zeros((len(c), 1), dtype=torch.int32, device=c.device), c], dim=1)
    c = c.contiguous()
    n = len(c)
    dev = c.device
    if n == 0:
        return torch.zeros(0, dtype=torch.int32, device=dev)
    cap = int(L.imf_hash_capacity(n))
    table = torch.empty(int(L.imf_hash_bytes(cap)), dtype=torch.uint8, device=dev)
    out = torch.empty((n, 4), dtype=torch.int32, device=dev)
    first = torch.empty(n, dtype=torch.int32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = int(L.imf_stride_map_workspace_bytes(n))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.imf_stride_map(_lib.ptr(c), None, n, 1, _lib.ptr(table), cap, _lib.ptr(out), _lib.ptr(meta[1:]),
                                    _lib.ptr(first), _lib.ptr(ws), ws_bytes, _lib.ptr(meta), _lib.cur_stream()))
        status, nu = meta.cpu().tolist()
    if status:
        from .sparse import _raise_status
        _raise_status(status)
    return first[:nu]


def voxelize(xyz: torch.Tensor, voxel_size: float, batch_index: int = 0):
    """xyz float64 or float32 [N,3] on CUDA -> (coords int32 [U,4] (b,x,y,z), idx int32 [U]) exactly as
    np.floor(xyz/voxel) -> sparse_quantize(return_index=True) -> batched_coordinates would give.  The division runs in the cloud's
    own dtype, as numpy's does (a float32 cloud is NOT promoted: near voxel boundaries float64 would pick other voxels)."""
    _lib.require_cuda(xyz, "points")
    L = _lib.lib()
    if xyz.dtype not in (torch.float32, torch.float64):
        xyz = xyz.to(torch.float64)          # integer / half clouds: numpy's true division gives float64 (float16 is not a cloud dtype)
    x = xyz.contiguous()
    n = len(x)
    c = torch.empty((n, 4), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        if x.dtype == torch.float32:
            _lib.check(L.imf_quantize_points_f32(_lib.ptr(x), n, float(voxel_size), int(batch_index), _lib.ptr(c), _lib.cur_stream()))
        else:
            _lib.check(L.imf_quantize_points(_lib.ptr(x), n, float(voxel_size), int(batch_index), _lib.ptr(c), _lib.cur_stream()))
    idx = unique_first(c)
    return c[idx.long()], idx
