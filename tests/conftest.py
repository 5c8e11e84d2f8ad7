import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def state_dict():
    from imfnet_b200 import synthetic
    return synthetic.make_state_dict(0)


@pytest.fixture(scope="session")
def cuda_model(state_dict):
    import torch
    from imfnet_b200 import load_model
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    Model = load_model("ResUNetBN2C")
    m = Model(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    m.load_state_dict(state_dict, strict=True)
    return m.eval().to("cuda:0")
