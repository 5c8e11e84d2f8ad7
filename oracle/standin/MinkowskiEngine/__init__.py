"""CPU ORACLE (test infrastructure) -- pure-torch stand-in for the 12 MinkowskiEngine 0.5.4 symbols
that /root/reference/model/*.py and util/misc.py use, so the reference's own model files execute
UNMODIFIED on CPU.  Put `oracle/standin` on sys.path to `import MinkowskiEngine`.

Used surface (call sites in the reference):
  MinkowskiNetwork                 model/resunet.py:14,33
  MinkowskiConvolution             model/resunet.py:42,54,66,78,136,149; model/residual_block.py:23,26
  MinkowskiConvolutionTranspose    model/resunet.py:101,114,125
  MinkowskiBatchNorm / InstanceNorm  model/common.py:6,8
  SparseTensor (.F ._F .C key/manager, +=, two ctor forms)   util/misc.py:95; model/resunet.py:189,229-233
  cat                              model/resunet.py:197,208,219
  MinkowskiFunctional.relu         model/resunet.py:171..224; model/residual_block.py:42,51
  utils.sparse_quantize / batched_coordinates / sparse_collate / fnv_hash_vec
                                   util/misc.py:83,86; lib/data_loaders.py:68-69; scripts/evaluation_3dmatch.py:154-174

Semantics: SURVEY.md Appendix A (restated from ME 0.5.4; ME's source is not in /root/reference).
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.abspath(os.path.join(_here, "..", "..", ".."))
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from oracle import sparse_ops as _ops  # noqa: E402

from . import utils  # noqa: E402,F401
from . import MinkowskiFunctional  # noqa: E402,F401

__version__ = "0.5.4-oracle-standin"


class CoordinateMapKey:
    def __init__(self, tensor_stride: int):
        self.tensor_stride = int(tensor_stride)

    def get_tensor_stride(self):
        return [self.tensor_stride] * 3

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and other.tensor_stride == self.tensor_stride

    def __hash__(self):
        return hash(self.tensor_stride)

    def __repr__(self):
        return f"CoordinateMapKey(tensor_stride={self.tensor_stride})"


class SparseTensor:
    def __init__(self, features, coordinates=None, coordinate_map_key=None, coordinate_manager=None,
                 tensor_stride=1, device=None, **_unused):
        if device is not None:
            features = features.to(device)
        self._F = features
        if coordinates is not None:
            assert coordinate_manager is None and coordinate_map_key is None
            C = coordinates.detach().cpu().numpy() if isinstance(coordinates, torch.Tensor) else np.asarray(coordinates)
            assert C.ndim == 2 and C.shape[1] == 4 and len(C) == len(features)
            self.coordinate_manager = _ops.CoordinateManager(C.astype(np.int32))
            self.coordinate_map_key = CoordinateMapKey(1)
        else:
            assert coordinate_manager is not None and coordinate_map_key is not None
            self.coordinate_manager = coordinate_manager
            self.coordinate_map_key = coordinate_map_key
            assert len(coordinate_manager.get(coordinate_map_key.tensor_stride)) == len(features)

    @property
    def F(self):
        return self._F

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.get(self.coordinate_map_key.tensor_stride).C)

    coordinates = C
    features = F

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

    def __iadd__(self, other):
        assert other.coordinate_map_key == self.coordinate_map_key
        assert other.coordinate_manager is self.coordinate_manager
        self._F = self._F + other._F
        return self

    def __add__(self, other):
        assert other.coordinate_map_key == self.coordinate_map_key
        return SparseTensor(self._F + other._F, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)


def cat(*tensors):
    t0 = tensors[0]
    for t in tensors[1:]:
        assert t.coordinate_map_key == t0.coordinate_map_key and t.coordinate_manager is t0.coordinate_manager
    return SparseTensor(torch.cat([t.F for t in tensors], dim=1), coordinate_map_key=t0.coordinate_map_key,
                        coordinate_manager=t0.coordinate_manager)


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class _ConvBase(nn.Module):
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3 and dilation == 1 and kernel_generator is None
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = int(kernel_size), int(stride), int(dilation)
        self.kernel_volume = self.kernel_size ** 3
        self.use_mm = self.kernel_volume == 1 and self.stride == 1
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        t = x.coordinate_map_key.tensor_stride
        if self.use_mm:
            Y = x.F @ self.kernel
            if self.bias is not None:
                Y = Y + self.bias
            return SparseTensor(Y, coordinate_map_key=x.coordinate_map_key, coordinate_manager=cm)
        if self.is_transpose:
            assert t % self.stride == 0
            t_out = t // self.stride
            nbr = cm.table(t, t_out, self.kernel_size, True)
        else:
            t_out = cm.stride(t, self.stride) if self.stride > 1 else t
            nbr = cm.table(t, t_out, self.kernel_size, False)
        Y = _ops.conv_forward(x.F, self.kernel, nbr, self.bias)
        return SparseTensor(Y, coordinate_map_key=CoordinateMapKey(t_out), coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(_ConvBase):
    is_transpose = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return SparseTensor(self.bn(x.F), coordinate_map_key=x.coordinate_map_key,
                            coordinate_manager=x.coordinate_manager)


class MinkowskiInstanceNorm(nn.Module):
    """Per batch item, per channel normalisation over points (only the *IN* model variants)."""

    def __init__(self, num_features, dimension=-1):
        super().__init__()
        self.eps = 1e-6
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x: SparseTensor) -> SparseTensor:
        b = x.C[:, 0].long()
        out = torch.empty_like(x.F)
        for i in range(int(b.max()) + 1 if len(b) else 0):
            m = b == i
            f = x.F[m]
            mu = f.mean(0, keepdim=True)
            var = ((f - mu) ** 2).mean(0, keepdim=True)
            out[m] = (f - mu) / torch.sqrt(var + self.eps) * self.weight + self.bias
        return SparseTensor(out, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)
