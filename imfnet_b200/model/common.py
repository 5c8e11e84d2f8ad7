"""Norm factory with the reference's signature (/root/reference/model/common.py:4-10)."""
from .. import me as ME


def get_norm(norm_type, num_feats, bn_momentum=0.05, D=-1):
    if norm_type == 'BN':
        return ME.MinkowskiBatchNorm(num_feats, momentum=bn_momentum)
    if norm_type == 'IN':
        raise NotImplementedError("instance-norm blocks (ResUNetIN2*) are outside the configured descriptor path (ResUNetBN2C)")
    raise ValueError(f'Type {norm_type}, not defined')
