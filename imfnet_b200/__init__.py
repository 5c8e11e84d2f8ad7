"""imfnet_b200 -- B200-native (sm_100a) implementation of IMFNet's descriptor-extraction hot path.

    from imfnet_b200 import load_model, extract_features
    import imfnet_b200.me as ME            # SparseTensor, utils.sparse_quantize, ... (MinkowskiEngine-like surface)

Everything numeric runs in the in-tree CUDA library (imfnet_b200/csrc, C ABI in include/imfnet_b200.h);
there is no CPU fallback.  Importing the package does not need a GPU; running a forward does.
"""
from .model import load_model  # noqa: F401
from .pipeline import extract_features  # noqa: F401

__version__ = "0.1.0"
