"""Sparse-tensor container and device coordinate manager (host mirror of the MinkowskiEngine surface that
IMFNet uses; semantics = oracle/sparse_ops.py = SURVEY.md Appendix A).

Reference call sites served:
  ME.SparseTensor(feats, coordinates=coords, device=device)                       util/misc.py:95
  ME.SparseTensor(F, coordinate_map_key=..., coordinate_manager=...)              model/resunet.py:229-233
  .F / ._F (assignable) / .C / len() / +=                                         model/resunet.py:189; residual_block.py:50
  ME.cat, MEF.relu                                                                model/resunet.py:171-225
All coordinate work (hash build, stride maps, neighbour tables) runs in the CUDA library through the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

STATUS_MSG = {1: "coordinate out of range (batch in [0,65534], |x|,|y|,|z| < 32768)",
              2: "duplicate coordinates (rows of a SparseTensor must be unique; use utils.sparse_quantize)",
              4: "coordinate hash table overflow"}


def _raise_status(status: int):
    if status:
        msgs = [m for b, m in STATUS_MSG.items() if status & b]
        raise ValueError("invalid sparse coordinates: " + "; ".join(msgs))


class CoordinateMapKey:
    """Identifies a coordinate set of a manager by its tensor stride (ME CoordinateMapKey)."""

    def __init__(self, tensor_stride: int):
        self.tensor_stride = int(tensor_stride)

    def get_tensor_stride(self):
        return [self.tensor_stride] * 3

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and other.tensor_stride == self.tensor_stride

    def __hash__(self):
        return hash(("imf", self.tensor_stride))

    def __repr__(self):
        return f"CoordinateMapKey(tensor_stride={self.tensor_stride})"


@dataclass
class Level:
    coords: torch.Tensor      # int32 [n,4] on device (rows 0..n-1 valid)
    n: int
    table: torch.Tensor       # uint8 hash-table storage
    capacity: int


class CoordinateManager:
    """Coordinate sets (one per tensor stride) + cached neighbour tables, all resident on the device."""

    def __init__(self, coords: torch.Tensor, check: bool = True):
        _lib.require_cuda(coords, "coordinates")
        if coords.dtype != torch.int32 or coords.dim() != 2 or coords.shape[1] != 4:
            raise ValueError("coordinates must be int32 [N,4] rows of (batch, x, y, z)")
        coords = coords.contiguous()
        self.device = coords.device
        self._l1 = (coords, coords.shape[0])     # the stride-1 table is built on first use (the captured-graph plan never needs it)
        self.levels = {}
        self.meta = None
        self._next_meta = 1
        self._tables = {}
        self._checked = not check
        self._segments = {}

    def _ensure_level1(self):
        if 1 in self.levels:
            return
        coords, n = self._l1
        L = _lib.lib()
        cap = int(L.imf_hash_capacity(n))
        table = torch.empty(int(L.imf_hash_bytes(cap)), dtype=torch.uint8, device=self.device)
        # meta = [status, n(stride 2), n(stride 4), ... ] device scalars
        self.meta = torch.zeros(16, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(L.imf_hash_build(_lib.ptr(coords), None, n, _lib.ptr(table), cap, _lib.ptr(self.meta), _lib.cur_stream()))
        self.levels[1] = Level(coords, n, table, cap)

    # -- status ---------------------------------------------------------------------------------
    def _check_status(self, meta_host=None):
        if self._checked:
            return
        status = int(self.meta[0].item()) if meta_host is None else int(meta_host[0])
        _raise_status(status)
        self._checked = True

    # -- coordinate sets ------------------------------------------------------------------------
    def level(self, t: int) -> Level:
        self._ensure_level1()
        return self.levels[t]

    def num_rows(self, t: int) -> int:
        return self._l1[1] if t == 1 else self.level(t).n

    def coords_of(self, t: int) -> torch.Tensor:
        return self._l1[0] if t == 1 else self.level(t).coords

    def build_pyramid(self, strides):
        """Create the coarser sets for every tensor stride in `strides` (ascending, each 2x the previous or any
        integer multiple) with ONE host read-back of the sizes."""
        L = _lib.lib()
        self._ensure_level1()
        todo = sorted(t for t in strides if t not in self.levels)
        if not todo and self._checked:
            return
        pending = []
        with torch.cuda.device(self.device):
            if todo:
                prev_t = max(t for t in self.levels if t < todo[0] and todo[0] % t == 0)
                prev = self.levels[prev_t]
                prev_n_dev, prev_max = None, prev.n       # exact size known for an existing level
            for t in todo:
                if t % prev_t != 0:
                    raise ValueError(f"tensor stride {t} is not a multiple of {prev_t}")
                cap = int(L.imf_hash_capacity(prev_max))
                table = torch.empty(int(L.imf_hash_bytes(cap)), dtype=torch.uint8, device=self.device)
                coords = torch.empty((max(prev_max, 1), 4), dtype=torch.int32, device=self.device)
                ws_bytes = int(L.imf_stride_map_workspace_bytes(prev_max))
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
                slot = self._next_meta
                self._next_meta += 1
                n_dev = self.meta[slot: slot + 1]
                _lib.check(L.imf_stride_map(_lib.ptr(prev.coords), _lib.ptr(prev_n_dev), prev_max, t, _lib.ptr(table), cap,
                                            _lib.ptr(coords), _lib.ptr(n_dev), None, _lib.ptr(ws), ws_bytes,
                                            _lib.ptr(self.meta), _lib.cur_stream()))
                lvl = Level(coords, -1, table, cap)
                pending.append((t, lvl, slot, ws))
                prev, prev_t, prev_n_dev = lvl, t, n_dev   # chained through the device-side count; prev_max stays an upper bound
            meta_host = self.meta.cpu()          # the one synchronisation of the coordinate phase
        self._check_status(meta_host)
        for t, lvl, slot, _ws in pending:
            lvl.n = int(meta_host[slot])
            lvl.coords = lvl.coords[: lvl.n]
            self.levels[t] = lvl

    def stride(self, t: int, s: int) -> int:
        t2 = t * s
        if t2 not in self.levels:
            self.build_pyramid([t2])
        return t2

    # -- neighbour tables -----------------------------------------------------------------------
    def table(self, t_in: int, t_out: int, K: int, transposed: bool) -> torch.Tensor:
        """int32 [n_out, K^3]; see include/imfnet_b200.h::imf_kernel_map."""
        key = (t_in, t_out, K, bool(transposed))
        if key not in self._tables:
            L = _lib.lib()
            self._ensure_level1()
            src, dst = self.levels[t_in], self.levels[t_out]
            scale = -t_out if transposed else t_in
            # allocated at the stride-1 row count (an upper bound for every level) so that fragments of one size class reuse
            # the same allocator blocks whatever their level sizes turn out to be
            nbr = torch.empty((max(self.levels[1].n, dst.n), K ** 3), dtype=torch.int32, device=self.device)[: dst.n]
            with torch.cuda.device(self.device):
                _lib.check(L.imf_kernel_map(_lib.ptr(dst.coords), None, dst.n, _lib.ptr(src.table), src.capacity, K, scale,
                                            _lib.ptr(nbr), _lib.cur_stream()))
            self._tables[key] = nbr
        return self._tables[key]

    def table_t(self, t_in: int, t_out: int, K: int, transposed: bool):
        """Offset-major neighbour table for the TMA-gather convolution: (nbr_t int32 [K^3, ld_n], ld_n, tile_mask uint32);
        see include/imfnet_b200.h::imf_kernel_map_t."""
        key = ("t", t_in, t_out, K, bool(transposed))
        if key not in self._tables:
            L = _lib.lib()
            self._ensure_level1()
            src, dst = self.levels[t_in], self.levels[t_out]
            scale = -t_out if transposed else t_in
            ld_n = (dst.n + 127) // 128 * 128
            nbr_t = torch.empty((K ** 3, max(ld_n, 128)), dtype=torch.int32, device=self.device)
            tile_mask = torch.empty(ld_n // 128 + 1, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(L.imf_kernel_map_t(_lib.ptr(dst.coords), None, dst.n, _lib.ptr(src.table), src.capacity, K, scale,
                                              _lib.ptr(nbr_t), nbr_t.stride(0), _lib.ptr(tile_mask), _lib.cur_stream()))
            self._tables[key] = (nbr_t, nbr_t.stride(0), tile_mask)
        return self._tables[key]

    def batch_segments(self, t: int, num_batches: int):
        """Host list seg[0..B] of row offsets per batch item at tensor stride t (rows are batch-sorted)."""
        key = (t, num_batches)
        if key not in self._segments:
            L = _lib.lib()
            lvl = self.level(t)
            seg = torch.empty(num_batches + 1, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(L.imf_batch_segments(_lib.ptr(lvl.coords), None, lvl.n, num_batches, _lib.ptr(seg), _lib.cur_stream()))
            self._segments[key] = seg.cpu().tolist()
        return self._segments[key]


class SparseTensor:
    """Features [N,C] fp32 + coordinates int32 [N,4] on one CUDA device."""

    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None, tensor_stride=1,
                 device=None, **_ignored):
        if not isinstance(features, torch.Tensor):
            features = torch.as_tensor(np.asarray(features), dtype=torch.float32)
        if device is not None:
            features = features.to(device)
        if coordinates is not None:
            if coordinate_manager is not None or coordinate_map_key is not None:
                raise ValueError("pass either coordinates or (coordinate_map_key, coordinate_manager)")
            if not isinstance(coordinates, torch.Tensor):
                coordinates = torch.as_tensor(np.asarray(coordinates))
            coordinates = coordinates.to(device=features.device, dtype=torch.int32)
            if len(coordinates) != len(features):
                raise ValueError("features and coordinates must have the same number of rows")
            coordinate_manager = CoordinateManager(coordinates)
            coordinate_map_key = CoordinateMapKey(1)
        elif coordinate_manager is None or coordinate_map_key is None:
            raise ValueError("coordinates, or coordinate_map_key and coordinate_manager, are required")
        elif coordinate_manager.num_rows(coordinate_map_key.tensor_stride) != len(features):
            raise ValueError("feature rows do not match the coordinate map")
        self._F = features
        self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key

    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.coords_of(self.coordinate_map_key.tensor_stride)

    coordinates = C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def device(self):
        return self._F.device

    @property
    def D(self):
        return 3

    def __len__(self):
        return len(self._F)

    def _same_map(self, other):
        if other.coordinate_manager is not self.coordinate_manager or other.coordinate_map_key != self.coordinate_map_key:
            raise ValueError("sparse tensors live on different coordinate maps")

    def __iadd__(self, other):
        self._same_map(other)
        self._F = self._F + other._F
        return self

    def __add__(self, other):
        self._same_map(other)
        return SparseTensor(self._F + other._F, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def __repr__(self):
        return f"SparseTensor(rows={len(self)}, channels={self._F.shape[1]}, stride={self.coordinate_map_key.tensor_stride}, device={self.device})"


def cat(*tensors):
    """Channel concatenation of tensors on the same coordinate map (ME.cat, model/resunet.py:197,208,219)."""
    t0 = tensors[0]
    for t in tensors[1:]:
        t0._same_map(t)
    return SparseTensor(torch.cat([t.F for t in tensors], dim=1), coordinate_map_key=t0.coordinate_map_key,
                        coordinate_manager=t0.coordinate_manager)
