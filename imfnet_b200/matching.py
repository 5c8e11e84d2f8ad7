"""Descriptor matching on the GPU (BASELINE config 3: 5000-keypoint L2 feature matching).

Host mirror of the two reference entry points that do nearest-neighbour search in descriptor space:
  find_nn_gpu(F0, F1, nn_max_n=-1, return_distance=False, dist_type='SquareL2')      /root/reference/lib/eval.py:18-48
  mutual_nn(frag1_descs, frag2_descs)    = the two uio.knn_search calls + mutual check  scripts/evaluation_3dmatch.py:207-217
Both run imf_nn_search_tc (csrc/matching_tc.cu: the 5000 x 5000 x 32 distance matrix as a tcgen05 product that filters the candidates,
which are then re-evaluated exactly) for 16- / 32-channel descriptors and imf_nn_search (csrc/matching.cu: brute force in fp32) otherwise;
both give the same indices and distances bit for bit, first index wins ties.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def nn_search(A: torch.Tensor, B: torch.Tensor, return_distance: bool = False, tensor_cores: bool = True):
    """idx[i] = argmin_j ||A[i] - B[j]||^2 (int32 on A's device); optionally the squared distances.
    tensor_cores=False forces the brute-force kernel (the two are bit-identical; tests compare them)."""
    _lib.require_cuda(A, "descriptors")
    if B.device != A.device:
        raise RuntimeError("both descriptor sets must live on the same CUDA device")
    A, B = A.float(), B.float()
    if A.stride(1) != 1:
        A = A.contiguous()
    if B.stride(1) != 1:
        B = B.contiguous()
    if A.shape[1] != B.shape[1] or A.shape[1] not in (16, 32, 64):
        raise ValueError("descriptors must have 16, 32 or 64 channels (IMFNet: 32) and the same width on both sides")
    L = _lib.lib()
    na, nb = A.shape[0], B.shape[0]
    idx = torch.empty(na, dtype=torch.int32, device=A.device)
    d2 = torch.empty(na, dtype=torch.float32, device=A.device) if return_distance else None
    tc = tensor_cores and A.shape[1] in (16, 32)
    ws_bytes = int(L.imf_nn_search_tc_workspace_bytes(na, nb) if tc else L.imf_nn_search_workspace_bytes(na))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=A.device)
    with torch.cuda.device(A.device):
        fn = L.imf_nn_search_tc if tc else L.imf_nn_search
        _lib.check(fn(_lib.ptr(A), A.stride(0) if na else A.shape[1], na, _lib.ptr(B), B.stride(0) if nb else B.shape[1], nb,
                                   A.shape[1], _lib.ptr(idx), _lib.ptr(d2), _lib.ptr(ws), ws_bytes, _lib.cur_stream()))
    return (idx, d2) if return_distance else idx


def find_nn_gpu(F0, F1, nn_max_n=-1, return_distance=False, dist_type='SquareL2'):
    """lib/eval.py:18-48: nearest neighbour of every row of F0 in F1; returns CPU tensors like the reference
    (inds int64 [N]; dists [N,1], squared for 'SquareL2', Euclidean for 'L2').  nn_max_n (chunking against the
    reference's N x M distance matrix) is accepted and irrelevant: nothing N x M is materialised here."""
    if dist_type not in ('SquareL2', 'L2'):
        raise NotImplementedError(f'Not implemented: dist_type {dist_type}')
    idx, d2 = nn_search(F0, F1, return_distance=True)
    inds = idx.long().cpu()
    if not return_distance:
        return inds
    d = d2 if dist_type == 'SquareL2' else torch.sqrt(d2 + 1e-7)          # lib/metrics.py:26-29
    return inds, d.unsqueeze(1).cpu()


def mutual_nn(frag1_descs, frag2_descs, device="cuda:0"):
    """scripts/evaluation_3dmatch.py:207-217 -> (frag21_nnindices [N2] int32, frag2_match_indices): for every descriptor of
    fragment 2 its nearest neighbour in fragment 1, and the rows of fragment 2 whose match is mutual."""
    d1 = torch.as_tensor(np.asarray(frag1_descs), dtype=torch.float32).to(device)
    d2 = torch.as_tensor(np.asarray(frag2_descs), dtype=torch.float32).to(device)
    nn21 = nn_search(d2, d1)
    nn12 = nn_search(d1, d2)
    from .pipeline import mutual_from_nn
    mutual = mutual_from_nn(nn12, nn21)          # (rows without a match, index -1, are never mutual)
    return nn21.cpu().numpy().astype(np.int32), mutual.cpu().numpy()
