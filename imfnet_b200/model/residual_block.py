"""Residual block container: conv3 -> norm -> ReLU -> conv3 -> norm -> (+x) -> ReLU
(/root/reference/model/residual_block.py:9-53; same sub-module names, so the same state_dict keys).

`forward` is the general per-layer route; the configured model executes these layers through the fused plan in
imfnet_b200/engine.py, where norm/ReLU/residual live in the convolution epilogues."""
import torch.nn as nn

from .. import me as ME
from .common import get_norm


class BasicBlockBase(nn.Module):
    expansion = 1
    NORM_TYPE = 'BN'

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, D=3):
        super().__init__()
        self.conv1 = ME.MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=stride, dimension=D)
        self.norm1 = get_norm(self.NORM_TYPE, planes, bn_momentum=bn_momentum, D=D)
        self.conv2 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=dilation, bias=False, dimension=D)
        self.norm2 = get_norm(self.NORM_TYPE, planes, bn_momentum=bn_momentum, D=D)
        self.downsample = downsample

    def forward(self, x):
        out = ME.MinkowskiFunctional.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        out += x if self.downsample is None else self.downsample(x)
        return ME.MinkowskiFunctional.relu(out)


class BasicBlockBN(BasicBlockBase):
    NORM_TYPE = 'BN'


class BasicBlockIN(BasicBlockBase):
    NORM_TYPE = 'IN'


def get_block(norm_type, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, D=3):
    if norm_type == 'BN':
        return BasicBlockBN(inplanes, planes, stride, dilation, downsample, bn_momentum, D)
    if norm_type == 'IN':
        return BasicBlockIN(inplanes, planes, stride, dilation, downsample, bn_momentum, D)
    raise ValueError(f'Type {norm_type}, not defined')
