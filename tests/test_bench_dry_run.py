"""CPU: bench.py's measurement logic (execution modes, fallback when the batched plan fails, JSON contract) run end to end on the
C-ABI emulator with inert torch.cuda objects.  The numbers are meaningless (every "CUDA event" reports 1 ms); the keys, the
mode selection and the bookkeeping are what is checked."""
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
    monkeypatch.setattr(b, "N_FRAGMENTS", 3)
    monkeypatch.setattr(b, "dominant_kernel_roofline", lambda model, frag, flush: {"bound": "hbm", "frac": 0.0})
    return b


def run(b, **over):
    args = types.SimpleNamespace(config="T", streams=2, batched=2, steps=1, warmup=1, profile=False, gpus=1, batched_note="forced")
    for k, v in over.items():
        setattr(args, k, v)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        b.run_ours(args, 0, 1, 0)
    lines = buf.getvalue().strip().splitlines()
    assert len(lines) == 1, "bench must print exactly one line"
    return json.loads(lines[0])


def test_bench_times_both_modes_and_reports_the_faster(bench):
    d = run(bench)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in d, key
    modes = d["config"]["execution_modes_timed"]
    assert set(modes) == {"single-fragment plans", "batched plan"}
    assert modes["batched plan"]["fragments_per_step"] == 4 and modes["single-fragment plans"]["fragments_per_step"] == 2
    assert d["config"]["fragments_per_step"] == 4 and "per batch of 2 fragments" in d["config"]["execution"]      # 4 fragments per fake ms wins
    assert d["e2e"]["h2d_bytes_per_step"] == 4 * (400 * 16 + 400 * 4 + 3 * 48 * 64 * 4) and d["e2e"]["d2h_bytes_per_step"] == 4 * 400 * 32 * 4
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0


def test_bench_falls_back_when_the_batched_plan_fails(bench, monkeypatch):
    import imfnet_b200.batched as batched

    def boom(self, *a, **k):
        raise RuntimeError("injected failure")

    monkeypatch.setattr(batched.BatchGraphPlan, "launch_batch", boom)
    d = run(bench)
    assert set(d["config"]["execution_modes_timed"]) == {"single-fragment plans"}
    assert "injected failure" in d["config"]["mode_selection"] and d["config"]["fragments_per_step"] == 2
    assert "per fragment" in d["config"]["execution"]


def test_bench_without_batched_mode(bench):
    d = run(bench, batched=0, batched_note="off")
    assert set(d["config"]["execution_modes_timed"]) == {"single-fragment plans"} and d["config"]["mode_selection"] == "off"


def test_mode_selection_logic(bench, monkeypatch, tmp_path_factory):
    """select_modes(): a variant library is only chosen when bit-identical and faster (the fastest wins); the batched plan only when
    the chosen library's probe of it passed."""
    import os
    ok = {"probe": "done", "B": 10, "hashes": ["a", "b"], "seq_ms_per_step": 10.0, "batched": "ok", "max_rowwise_rel_diff_vs_forward_many": 0.0}

    def with_probes(default, x, z, y=None, w=None):
        calls = []

        def fake(args, v="", sizes=""):
            calls.append(v)
            d = {"": default, "x": x, "z": z, "y": y, "w": w}[v]
            return (dict(d), "ok") if d is not None else (None, "probe failed (rc 1): boom")

        monkeypatch.setattr(bench, "run_probe", fake)
        monkeypatch.delenv("IMFNET_B200_VARIANT", raising=False)
        monkeypatch.setattr(bench, "CACHE_DIR", str(tmp_path_factory.mktemp("modes")))      # no decision cache between cases
        args = types.SimpleNamespace(config="T", streams=10, variant_probe=True)
        res = bench.select_modes(args)
        chosen = os.environ.get("IMFNET_B200_VARIANT", "")
        monkeypatch.delenv("IMFNET_B200_VARIANT", raising=False)
        return res, chosen, calls

    (B, note), var, calls = with_probes(ok, dict(ok, seq_ms_per_step=8.0), dict(ok, seq_ms_per_step=9.0))
    assert B == 10 and var == "x" and "variant x in use" in note and calls == ["", "x", "z", "y", "w"]
    (B, note), var, _ = with_probes(ok, dict(ok, seq_ms_per_step=8.0), dict(ok, seq_ms_per_step=7.0))
    assert var == "z"
    (B, note), var, _ = with_probes(ok, dict(ok, seq_ms_per_step=8.0), dict(ok, seq_ms_per_step=7.0), dict(ok, seq_ms_per_step=6.0))
    assert var == "y"
    (B, note), var, _ = with_probes(ok, dict(ok, seq_ms_per_step=9.9), dict(ok, seq_ms_per_step=10.5))
    assert B == 10 and var == "" and "default library in use" in note
    (B, note), var, _ = with_probes(ok, dict(ok, hashes=["a", "c"], seq_ms_per_step=5.0), None)
    assert B == 10 and var == "" and "differ" in note and "probe failed" in note
    (B, note), var, _ = with_probes(dict(ok, batched="mismatch"), dict(ok, seq_ms_per_step=8.0), None)
    assert B == 10 and var == "x"                                  # the chosen library's own batched probe passed
    (B, note), var, _ = with_probes(dict(ok, batched="failed: x"), dict(ok, seq_ms_per_step=20.0, batched="failed: x"), None)
    assert B == 0 and var == "" and "not used" in note
    (B, note), var, calls = with_probes(None, ok, ok)
    assert B == 0 and var == "" and calls == [""]


def test_mode_selection_is_cached_per_box(bench, monkeypatch, tmp_path):
    import os
    ok = {"probe": "done", "B": 10, "hashes": ["a"], "seq_ms_per_step": 10.0, "batched": "ok", "max_rowwise_rel_diff_vs_forward_many": 0.0}
    calls = []

    def fake(args, v="", sizes=""):
        calls.append(v)
        return dict(ok, seq_ms_per_step=10.0 if v == "" else 8.0 if v == "x" else 9.0), "ok"

    monkeypatch.setattr(bench, "run_probe", fake)
    monkeypatch.setattr(bench, "CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.delenv("IMFNET_B200_VARIANT", raising=False)
    args = types.SimpleNamespace(config="T", streams=10, variant_probe=True)
    B1, note1 = bench.select_modes(args)
    assert os.environ.get("IMFNET_B200_VARIANT") == "x" and B1 == 10 and len(calls) == 5
    assert set(args.probe_table) == {"default", "x", "z", "y", "w"} and args.probe_table["x"]["bit_identical"] is True
    assert args.probe_table["x"]["seq_ms_per_step"] == 8.0
    table1 = args.probe_table
    args = types.SimpleNamespace(config="T", streams=10, variant_probe=True)
    monkeypatch.delenv("IMFNET_B200_VARIANT", raising=False)
    B2, note2 = bench.select_modes(args)                      # second run on the same box: no probes, same decision
    assert os.environ.get("IMFNET_B200_VARIANT") == "x" and B2 == 10 and len(calls) == 5 and "cached" in note2 and note2.startswith(note1)
    assert args.probe_table == table1
    monkeypatch.delenv("IMFNET_B200_VARIANT", raising=False)
