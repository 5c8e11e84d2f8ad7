"""CPU: host-side mirror -- registry, state_dict contract, synthetic inputs, sharding."""
import numpy as np
import pytest
import torch

from imfnet_b200 import load_model, synthetic
from imfnet_b200.pipeline import shard_indices
from oracle import sparse_ops


def test_load_model_registry():
    assert load_model("ResUNetBN2C").__name__ == "ResUNetBN2C"
    assert load_model("NoSuchNet") is None
    for name in ("ResUNetBN2", "ResUNetBN2B", "ResUNetBN2C", "ResUNetBN2D", "ResUNetBN2E"):
        assert load_model(name) is not None


@pytest.mark.parametrize("name", ["ResUNetBN2C", "ResUNetBN2E", "ResUNetBN2B"])
def test_state_dict_contract(name):
    m = load_model(name)(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    spec = {n: tuple(s) for n, s, _ in synthetic.state_dict_spec(name, 1, 32, 5)}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == spec
    m.load_state_dict(synthetic.make_state_dict(1, name), strict=True)
    if name == "ResUNetBN2C":
        assert len(spec) == 361 and sum(int(np.prod(s)) for s in spec.values()) == 31461153
        assert spec["conv1.kernel"] == (125, 1, 32) and spec["final.kernel"] == (64, 32) and spec["final.bias"] == (1, 32)
        assert spec["conv1_tr.kernel"] == (96, 64) and spec["conv3_tr.kernel"] == (27, 256, 64)


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        load_model("ResUNetIN2C")(1, 32, conv1_kernel_size=5, D=3)
    m = load_model("ResUNetBN2C")(1, 32, conv1_kernel_size=5, D=3)
    m.train()
    with pytest.raises(NotImplementedError):
        m(None, None)
    m.eval()
    with pytest.raises(RuntimeError):        # CPU parameters: no CPU fallback
        m(None, torch.zeros(1, 3, 8, 8))


def test_weights_are_seed_deterministic():
    a, b = synthetic.make_state_dict(0), synthetic.make_state_dict(0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["conv2.kernel"], synthetic.make_state_dict(1)["conv2.kernel"])
    assert float(a["norm1.bn.running_var"].min()) >= 0.5


def test_synthetic_fragment_statistics():
    coords, pts = synthetic.make_fragment(5000, 0.05, seed=0)
    assert coords.shape == (5000, 4) and coords.dtype == np.int32 and pts.shape == (5000, 3)
    assert len(np.unique(coords, axis=0)) == 5000
    assert np.array_equal(np.floor(pts / 0.05).astype(np.int32), coords[:, 1:])
    cm = sparse_ops.CoordinateManager(coords)
    t = cm.table(1, 1, 3, False)
    fill = (t >= 0).sum() / len(t)
    assert 9.0 < fill < 19.0          # real fragments: ~14 of 27
    c2, _ = synthetic.make_fragment(5000, 0.05, seed=0)
    assert np.array_equal(coords, c2)


def test_shard_indices_cover_everything_once():
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in shard_indices(37, r, world))
        assert got == list(range(37))
    w = [5, 1, 1, 1, 4, 4]
    parts = [shard_indices(len(w), r, 2, w) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == list(range(6))
    loads = [sum(w[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 1


def test_plan_cache_is_lru_and_bounded_by_bytes():
    """The model's cache of captured plans (engine.PlanCache) evicts least-recently-used entries beyond its byte budget."""
    from imfnet_b200.engine import PlanCache

    class P:
        def __init__(self, n):
            self.n = n

        def nbytes(self):
            return self.n

    c = PlanCache(budget_bytes=100)
    c[("a",)] = P(40)
    pool = c.setdefault(("pool", 1), [])
    pool.append(P(30))
    pool.append(P(20))
    c.trim(keep=("pool", 1))
    assert len(c) == 2 and c.nbytes() == 90
    assert c.get(("a",)).n == 40                      # touch: ("a",) is now the most recent entry
    c[("b",)] = P(50)
    c.trim(keep=("b",))
    assert ("pool", 1) not in c and ("a",) in c and ("b",) in c and c.evictions == 1
    c[("huge",)] = P(500)
    c.trim(keep=("huge",))                           # the entry in use is never dropped, everything else goes
    assert list(c.keys()) == [("huge",)]
    c.clear()
    assert len(c) == 0


def test_mutual_nn_ignores_rows_without_a_match():
    from imfnet_b200.pipeline import mutual_from_nn
    nn21 = torch.tensor([2, -1, 0, 1], dtype=torch.int32)          # row 1 of fragment 2 has no match (e.g. NaN descriptor)
    nn12 = torch.tensor([2, 3, 0], dtype=torch.int32)
    assert mutual_from_nn(nn12, nn21).tolist() == [0, 2, 3]
    assert mutual_from_nn(torch.zeros(0, dtype=torch.int32), torch.full((3,), -1, dtype=torch.int32)).tolist() == []
