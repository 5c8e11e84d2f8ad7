"""`import imfnet_b200.me as ME` -- the slice of the MinkowskiEngine Python API that IMFNet's descriptor path
touches, backed by the sm_100a kernels.  Layer modules keep ME's parameter names
(`kernel`, `bias`, `bn.*`) so reference checkpoints load unchanged.

These per-layer modules are the general (unfused) route, one kernel launch per layer; the configured model
(`imfnet_b200.model.resunet.ResUNet2.forward`) runs the fused plan in imfnet_b200/engine.py instead.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .sparse import CoordinateManager, CoordinateMapKey, SparseTensor, cat  # noqa: F401


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    """ME.MinkowskiConvolution / MinkowskiConvolutionTranspose (dimension 3, cubic kernel, dilation 1).
    kernel: [K^3, Cin, Cout], or [Cin, Cout] when K == 1 and stride == 1; bias: [1, Cout]."""
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, kernel_generator=None,
                 expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension != 3 or dilation != 1 or kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("imfnet_b200 covers dimension=3, dilation=1, cubic kernels (the IMFNet path)")
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride, self.dilation = int(kernel_size), int(stride), int(dilation)
        self.kernel_volume = self.kernel_size ** 3
        self.use_mm = self.kernel_volume == 1 and self.stride == 1
        shape = (self.in_channels, self.out_channels) if self.use_mm else (self.kernel_volume, self.in_channels, self.out_channels)
        self.kernel = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty((1, self.out_channels), dtype=torch.float32)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
        bound = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def extra_repr(self):
        return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"transpose={self.is_transpose}")

    def forward(self, x: SparseTensor) -> SparseTensor:
        if torch.is_grad_enabled() and (self.kernel.requires_grad and x.F.requires_grad):
            raise NotImplementedError("imfnet_b200 implements the inference forward only (no autograd)")
        L = _lib.lib()
        X = x.F.contiguous()
        _lib.require_cuda(X, "features")
        cm, t = x.coordinate_manager, x.coordinate_map_key.tensor_stride
        bias = None if self.bias is None else self.bias.detach().reshape(-1).contiguous()
        with torch.cuda.device(X.device):
            if self.use_mm:
                Y = torch.empty((len(X), self.out_channels), dtype=torch.float32, device=X.device)
                _lib.check(L.imf_linear_fwd(_lib.ptr(X), X.shape[1], _lib.ptr(self.kernel.detach()), _lib.ptr(bias), len(X),
                                            self.in_channels, self.out_channels, _lib.ptr(Y), self.out_channels,
                                            _lib.cur_stream()))
                return SparseTensor(Y, coordinate_map_key=x.coordinate_map_key, coordinate_manager=cm)
            if self.is_transpose:
                if t % self.stride != 0 or (t // self.stride) not in cm.levels:
                    raise ValueError("transposed convolution needs the finer coordinate map to exist already")
                t_out = t // self.stride
            else:
                t_out = cm.stride(t, self.stride) if self.stride > 1 else t
            n_out = cm.level(t_out).n
            Y = torch.empty((n_out, self.out_channels), dtype=torch.float32, device=X.device)
            W = self.kernel.detach()
            if self.in_channels % 32 != 0:
                if self.is_transpose or self.stride != 1 or self.in_channels not in (1, 3, 6):
                    raise NotImplementedError("Cin not a multiple of 32: only stride-1 convolutions with 1, 3 or 6 input "
                                              "channels (ones / rgb / rgb+normal, util/misc.py:66-77) are implemented")
                lvl = cm.level(t)
                _lib.check(L.imf_conv_first_fwd(_lib.ptr(X), X.shape[1], self.in_channels, _lib.ptr(W), _lib.ptr(lvl.coords),
                                                None, lvl.n, _lib.ptr(lvl.table), lvl.capacity, self.kernel_size, t,
                                                self.out_channels, None, None, 0, _lib.ptr(Y), self.out_channels,
                                                _lib.cur_stream()))
            else:
                nbr = cm.table(t, t_out, self.kernel_size, self.is_transpose)
                _lib.check(L.imf_sparse_conv_fwd(_lib.ptr(X), X.shape[1], _lib.ptr(W), _lib.ptr(nbr), None, n_out,
                                                 self.kernel_volume, self.in_channels, self.out_channels, None, None, None, 0,
                                                 0, _lib.ptr(Y), self.out_channels, _lib.cur_stream()))
            if bias is not None:
                Y += bias
        return SparseTensor(Y, coordinate_map_key=CoordinateMapKey(t_out), coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(_ConvBase):
    is_transpose = True


class MinkowskiBatchNorm(nn.Module):
    """ME.MinkowskiBatchNorm: nn.BatchNorm1d on the feature matrix (state_dict keys `<name>.bn.*`)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return SparseTensor(self.bn(x.F), coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)

    def folded(self):
        """(scale, shift) of the eval-mode affine map y = x*scale + shift."""
        bn = self.bn
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return scale.contiguous(), (bn.bias.detach() - bn.running_mean * scale).contiguous()


class MinkowskiFunctional:
    @staticmethod
    def relu(x: SparseTensor) -> SparseTensor:
        return SparseTensor(torch.relu(x.F), coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)


class utils:
    """ME.utils subset (util/misc.py:83,86; lib/data_loaders.py:68-69; scripts/evaluation_3dmatch.py:164-168)."""

    @staticmethod
    def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                        return_inverse=False, return_maps_only=False, quantization_size=None, device="cuda"):
        """floor -> int32 -> first-occurrence unique rows, ascending source index (runs on the GPU)."""
        if labels is not None or return_inverse:
            raise NotImplementedError("labels / return_inverse are not used by the IMFNet path")
        from .voxelize import unique_first
        is_torch = isinstance(coordinates, torch.Tensor)
        c = coordinates if is_torch else torch.as_tensor(np.asarray(coordinates))
        if quantization_size is not None:
            c = c / quantization_size
        d = torch.floor(c).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
        idx = unique_first(d.to(device)).to(d.device)
        out = d[idx.long()]
        if not is_torch:
            out, idx = out.numpy(), idx.numpy()
        if return_maps_only:
            return idx
        res = [out]
        if features is not None:
            res.append(features[idx.long() if is_torch else idx])
        if return_index:
            res.append(idx)
        return res[0] if len(res) == 1 else tuple(res)

    @staticmethod
    def fnv_hash_vec(arr):
        """FNV64-1A over the columns of an integer-valued [N, D] array -> uint64 [N] (ME.utils.fnv_hash_vec as called by
        scripts/evaluation_3dmatch.py:164-168 on np.floor(points / voxel_size) to intersect keypoints with voxel coordinates).
        Host numpy arithmetic like the original: it hashes a few thousand keypoints per fragment pair, not a hot path."""
        a = np.asarray(arr)
        if a.ndim != 2:
            raise ValueError("fnv_hash_vec expects a 2-D array [N, D]")
        a = a.copy().astype(np.uint64, copy=False)
        h = np.uint64(14695981039346656037) * np.ones(a.shape[0], dtype=np.uint64)
        with np.errstate(over="ignore"):
            for j in range(a.shape[1]):
                h *= np.uint64(1099511628211)
                h = np.bitwise_xor(h, a[:, j])
        return h

    @staticmethod
    def batched_coordinates(coords, dtype=torch.int32, device=None):
        rows = []
        for b, c in enumerate(coords):
            c = c if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c))
            if c.is_floating_point():
                c = torch.floor(c)
            c = c.to(dtype)
            rows.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype, device=c.device), c], dim=1))
        out = torch.cat(rows, dim=0) if rows else torch.zeros((0, 4), dtype=dtype)
        return out if device is None else out.to(device)

    @staticmethod
    def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
        as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
        bc = utils.batched_coordinates(coords, dtype=dtype, device=device)
        fs = torch.cat([as_t(f) for f in feats], dim=0)
        if labels is not None:
            return bc, fs, torch.cat([as_t(l) for l in labels], dim=0)
        return bc, fs
