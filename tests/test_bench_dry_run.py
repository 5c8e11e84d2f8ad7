"""CPU: bench.py's measurement logic (the batched execution mode, the per-layer roofline leg, the C4 pair workload, the JSON
contract) run end to end on the C-ABI emulator with inert torch.cuda objects.  The numbers are meaningless (every "CUDA event"
reports 1 ms); the keys and the bookkeeping are what is checked."""
import contextlib
import importlib.util
import io
import json
import os
import types

import pytest
import torch

from imfnet_b200 import synthetic

from test_plan_emulated import emu  # noqa: F401  (fixture)

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture
def bench(emu, monkeypatch):  # noqa: F811
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    from test_plan_plumbing import FakeEvent

    class Ev(FakeEvent):
        def elapsed_time(self, other):
            return 1.0

    real_device, real_empty = torch.device, torch.empty
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{n: v for n, v in k.items() if n != "pin_memory"}))
    monkeypatch.setitem(synthetic.CONFIGS, "T", (400, 0.05, 64, 48))
    monkeypatch.setitem(b.WORKLOADS, "C2", ("T", 2))
    monkeypatch.setitem(b.WORKLOADS, "C4", ("T", 2))
    monkeypatch.setattr(b, "N_FRAGMENTS", 3)
    monkeypatch.setattr(b, "KEYPOINTS", 50)
    monkeypatch.setattr(b, "count_conv1_pairs", lambda group, k, device: 40 * sum(len(c) for c, _f, _im in group))
    import imfnet_b200.matching as matching

    def nn_cpu(A, B, return_distance=False):          # (the matching kernel is not part of the emulator)
        d = torch.cdist(A.double(), B.double())
        return d.argmin(dim=1).to(torch.int32)

    monkeypatch.setattr(matching, "nn_search", nn_cpu)
    return b


def run(b, **over):
    args = types.SimpleNamespace(config="C2", batch=0, steps=1, warmup=1, profile=False, gpus=1, plans=2)
    for k, v in over.items():
        setattr(args, k, v)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        b.run_ours(args, 0, 1, 0)
    lines = buf.getvalue().strip().splitlines()
    assert len(lines) == 1, "bench must print exactly one line"
    return json.loads(lines[0])


def test_bench_line_contract_and_roofline_leg(bench):
    d = run(bench)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["config"]["fragments_per_step"] == 4 and d["config"]["fragments_per_graph_replay"] == 2 and d["config"]["plans_in_flight"] == 2
    assert d["value"] == pytest.approx(400 * 4 / 1e-3) and d["ms_per_step"] == pytest.approx(1.0)          # ms_per_step x steps reproduces value
    assert d["e2e"]["h2d_bytes_per_step"] == 4 * (400 * 16 + 400 * 4 + 3 * 48 * 64 * 4) and d["e2e"]["d2h_bytes_per_step"] == 4 * 400 * 32 * 4
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "sparse_part", "layers"):
        assert key in r, key
    names = [x["layer"] for x in r["layers"]]
    assert names[0] == "conv1" and names[-1] == "conv1_tr+final" and len(names) == 22 and "block2_tr.conv1" in names
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    sp = r["sparse_part"]
    assert sp["alg_bytes"] == pytest.approx(sum(x["alg_MB"] for x in r["layers"]) * 1e6, rel=1e-6) and sp["ms"] == pytest.approx(22.0)
    # SURVEY 8(d) accounting of one layer, by hand
    x = next(x for x in r["layers"] if x["layer"] == "block2_tr.conv2")
    assert x["alg_MB"] * 1e6 == pytest.approx(4 * x["pairs"] * 64 + 8 * x["pairs"] + 4 * 27 * 64 * 64 + 2 * 4 * x["rows"] * 64)


def test_bench_pairs_workload(bench):
    d = run(bench, config="C4")
    c = d["config"]
    assert c["pairs_per_step"] == 2 and c["fragments_per_step"] == 4 and c["pairs_per_s"] == pytest.approx(2 / 1e-3)
    assert "mutual-NN" in c["workload"]
    assert d["e2e"]["d2h_bytes_per_step"] == 4 * 400 * 32 * 4 + 2 * 50 * 4


def test_reference_arm_line(bench, monkeypatch, capsys):
    args = types.SimpleNamespace(config="C2", steps=1, warmup=0, gpus=1)
    bench.run_reference(args, 0, 1)
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["metric"] == bench.METRIC
    bench.run_reference(args, 1, 2)                      # other ranks exit without work or output
    assert capsys.readouterr().out == ""
