"""CPU ORACLE stand-in for MinkowskiEngine.MinkowskiFunctional (only `relu` is used:
/root/reference/model/resunet.py:171,176,181,186,194,205,216,225; model/residual_block.py:42,51)."""
import torch


def relu(x):
    from . import SparseTensor
    return SparseTensor(torch.relu(x.F), coordinate_map_key=x.coordinate_map_key,
                        coordinate_manager=x.coordinate_manager)
