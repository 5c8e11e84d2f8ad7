"""Headline benchmark: descriptor-extraction throughput (voxels/s) on 50 k-voxel / 640x480 synthetic 3DMatch fragments
(BASELINE.json configs[1], "C2"), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = full ResUNetBN2C.forward(x, image) on --streams (default 10) independent fragments, each through its own captured
CUDA-graph plan and stream (fragments are independent units; the single-fragment latency is reported in
config.single_fragment_latency_ms), coordinate maps rebuilt for every fragment (what the reference does for every new
SparseTensor), inputs already resident in HBM.  A second execution mode of the same public API -- the batched captured plan,
`model.forward_batches`, two groups of --streams fragments per step -- is timed as well when a parity probe of it passes in a
subprocess on this GPU (it must reproduce forward_many's descriptors); the faster mode is the headline and every timed mode is
listed in config.execution_modes_timed (--batched 0 switches this off).  The same probes decide whether the kernel variant library
(imfnet_b200/build.py VARIANTS: same sources, experiment switches on the convolution, attention and GEMM kernels) is loaded instead of the
default one:
only when its descriptors are bit-identical and its step is shorter (config.mode_selection, config.library).  Steps rotate over 8 distinct fragments per rank and the L2 is flushed
between steps (a 256 MiB write), outside the per-step CUDA events.  Multi-GPU: fragments are independent, each rank runs
its own (weak scaling); the only collective is the all-gather of per-rank timings.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CACHE_DIR = os.path.join(ROOT, ".bench_cache")      # mode decisions of the automatic mode (select_modes), keyed by machine boot id
METRIC = "descriptor-extraction throughput: voxels/sec/GPU on 50k-voxel fragments"
N_FRAGMENTS = 8
NCU_DRAM_BYTES_PER_LAUNCH = 18674944      # k_sparse_conv_g4<64,64>, 64->64 @ 50 000 voxels (profiles/r01/call55_g4_ncu_summary.txt)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(cfg: str, rank: int):
    from imfnet_b200 import synthetic
    target, voxel, W, H = synthetic.CONFIGS[cfg]
    frags = []
    for i in range(N_FRAGMENTS):
        seed = rank * N_FRAGMENTS + i
        coords, _ = synthetic.make_fragment(target, voxel, seed)
        frags.append((torch.from_numpy(coords), torch.ones((len(coords), 1)), synthetic.make_image(W, H, seed)))
    return frags, (target, voxel, W, H)


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path = the oracle port (the reference is pure Python over MinkowskiEngine,
    which is not installable here; oracle/ restates it and is pinned to the unmodified model files by tests/golden)."""
    if rank != 0:
        return
    from imfnet_b200 import synthetic
    from oracle import imfnet_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    frags, (target, voxel, W, H) = make_inputs(args.config, 0)
    sd = synthetic.make_state_dict(0)
    times = []
    for i in range(args.warmup + args.steps):
        c, f, im = frags[i % N_FRAGMENTS]
        t0 = time.perf_counter()
        imfnet_oracle.forward(sd, c, f, im)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = target * len(times) / total
    sample = f"{len(times)} cold forwards (coordinate maps rebuilt) of one {target}-voxel fragment each, fp32, torch CPU"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {target}-voxel synthetic 3DMatch fragment + {W}x{H} image, ResUNetBN2C 32-D descriptors",
                   "fragments_per_step": 1, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------------------
def dominant_kernel_roofline(model, frag, flush, layer="block2_tr.conv1", check=True):
    """block2_tr-shaped convolution (64->64, 3^3, stride-1 level, BatchNorm + ReLU folded): the largest single launch of
    the forward (179.6 MB algorithmic bytes at C2), run through the same entry point and packed weights the
    forward uses (imf_sparse_conv_g4_fwd).  Timed live with CUDA events on the launching stream, L2 flushed between launches."""
    from imfnet_b200 import _lib
    from imfnet_b200.sparse import CoordinateManager
    L = _lib.lib()
    coords = frag[0].cuda()
    cm = CoordinateManager(coords)
    nbr_t, ld_n, tile_mask = cm.table_t(1, 1, 3, False)
    n = len(coords)
    conv, packed, scale, shift, kci = model._plan.conv[layer]
    cin, cout = conv.in_channels, conv.out_channels
    kco = 64 if cout % 64 == 0 else 32
    s = torch.cuda.current_stream().cuda_stream
    X = torch.randn(n, cin, device="cuda")
    Xh = torch.empty(n, 2 * cin, dtype=torch.float16, device="cuda")
    _lib.check(L.imf_h2_pack(X.data_ptr(), cin, n, cin, kci, Xh.data_ptr(), 2 * cin, None, s))
    Yh = torch.empty(n, 2 * cout, dtype=torch.float16, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(cout))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    pairs = int((nbr_t[:, :n] >= 0).sum())
    # gathered inputs + in/out indices + weights once + output (activations are 4 bytes/channel: fp16 hi + lo)
    alg_bytes = 4 * pairs * cin + 8 * pairs + 4 * 27 * cin * cout + 4 * n * cout
    times = []
    for i in range(23):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.imf_sparse_conv_g4_fwd(Xh.data_ptr(), 2 * cin, kci, packed.data_ptr(), nbr_t.data_ptr(), ld_n, tile_mask.data_ptr(),
                                            None, n, 27, cin, cout, scale.data_ptr(), shift.data_ptr(), None, 0, 0, 1, Yh.data_ptr(),
                                            2 * cout, n, kco, ws.data_ptr(), ws_bytes, err.data_ptr(), s))
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    assert not check or int(err.item()) == 0
    ms = float(np.mean(times))
    peak, how = load_peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": f"k_sparse_conv_g4<{cout},{kci}> 3x3x3 {cin}->{cout} @ {n} voxels ({pairs} pairs)", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "alg_bytes_per_launch": alg_bytes,
            "ms_per_launch": ms, "flops_per_launch": 2 * pairs * cin * cout, "peak_source": how,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch (profiles/r01/call55_g4_ncu_summary.txt)"}


def probe_batched(args, local_rank):
    """Subprocess body of the automatic mode selection (runs with the library variant named by IMFNET_B200_VARIANT, default
    none).  Prints one JSON line with
      * sha256 of the descriptors forward_many gives for the distinct fragments (the parent compares variants bit for bit),
      * a short timing of the default execution mode (the parent only switches to a variant that is faster),
      * whether the batched captured plan reproduces forward_many fragment by fragment (device and pinned-host inputs)."""
    import hashlib
    import imfnet_b200.me as ME
    from imfnet_b200 import load_model, synthetic
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    frags, (target, voxel, W, H) = make_inputs(args.config, 0)
    model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    model.load_state_dict(synthetic.make_state_dict(0), strict=True)
    model = model.eval().to(dev)
    B = max(1, args.streams)
    sel = [frags[j % N_FRAGMENTS] for j in range(2 * B)]
    dev_frags = [(c.to(dev), f.to(dev), im.to(dev)) for c, f, im in sel]
    pin_frags = [(c.pin_memory(), f.pin_memory(), im.pin_memory()) for c, f, im in sel]
    out = {"probe": "done", "variant": os.environ.get("IMFNET_B200_VARIANT", ""), "B": B}
    with torch.no_grad():
        items = [(ME.SparseTensor(f, coordinates=c), im) for c, f, im in dev_frags]
        ref = [o.F for o in model.forward_many(items, streams=B)]
        if not all(bool(torch.isfinite(r).all()) for r in ref):
            raise SystemExit("probe: non-finite descriptors")
        out["hashes"] = [hashlib.sha256(r.cpu().numpy().tobytes()).hexdigest() for r in ref[:N_FRAGMENTS]]
        for rep in range(3):          # run-to-run reproducibility of this library (every repeat must give the same bits)
            again = [o.F for o in model.forward_many(items, streams=B)]
            if [hashlib.sha256(r.cpu().numpy().tobytes()).hexdigest() for r in again[:N_FRAGMENTS]] != out["hashes"]:
                raise SystemExit("probe: descriptors differ from run to run")
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.forward_many(items[:B], streams=B)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["seq_ms_per_step"] = float(np.median(ts[3:]))
        try:          # the dominant kernel alone with this library (information for the next round; the decision does not use it)
            fb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            out["conv64_us_per_launch"] = 1e3 * dominant_kernel_roofline(model, frags[0], lambda: fb.fill_(1))["ms_per_launch"]
            out["conv32_us_per_launch"] = 1e3 * dominant_kernel_roofline(model, frags[0], lambda: fb.fill_(1), "block1.conv1")["ms_per_launch"]
            # the kernel's own profiling flags (results meaningless, timing only): 3 = no gathers + no weight copies, 6 = no gathers +
            # no MMAs, 7 = pipeline skeleton only -- what tools/conv_g4_bench.py --flags prints
            from imfnet_b200 import _lib
            out["conv64_us_by_debug_flags"] = {}
            try:
                for fl in (3, 6, 7):
                    _lib.lib().imf_debug_conv_g4_trace(None, 0, 0, fl)
                    out["conv64_us_by_debug_flags"][str(fl)] = 1e3 * dominant_kernel_roofline(model, frags[0], lambda: fb.fill_(1),
                                                                                             check=False)["ms_per_launch"]
            except Exception as ex:      # noqa: BLE001
                out["conv64_us_by_debug_flags"]["error"] = f"{ex!r}"[:120]
            finally:
                _lib.lib().imf_debug_conv_g4_trace(None, 0, 0, 0)
            del fb
        except Exception as ex:      # noqa: BLE001
            out["conv64_us_per_launch"] = f"failed: {ex!r}"[:120]
        # the batched captured plan at batch sizes B and B / 2 (a batch of 10 x 50 k voxels is 128 MB of 64-channel activations, about
        # the L2's size: the smaller batch may be the faster one); each must reproduce forward_many before its time counts
        out["batched_by_size"] = {}
        best = None
        sizes = [int(x) for x in args.probe_sizes.split(",") if x] if args.probe_sizes else [B] + ([B // 2] if B >= 4 else [])
        for b in sizes:
            rec = {}
            try:
                worst = 0.0
                sub_d, sub_p, sub_r = dev_frags[:2 * b], pin_frags[:2 * b], ref[:2 * b]          # (b <= B: 2 B fragments were prepared)
                for rep in range(2):          # the second round re-uses the captured plans
                    outs = model.forward_batches(sub_d, b, streams=2)
                    outs_h = model.forward_batches(sub_p, b, streams=2)
                    for o, oh, r in zip(outs, outs_h, sub_r):
                        for x in (o, oh.to(dev)):
                            if x.shape != r.shape or not bool(torch.isfinite(x).all()):
                                raise RuntimeError("bad output")
                            worst = max(worst, float((torch.linalg.norm(x - r, dim=1) / torch.linalg.norm(r, dim=1)).max()))
                torch.cuda.synchronize()
                # expected 0 (same kernels, same per-row summation order); the parity bar against the oracle is 1e-4
                rec["verdict"] = "ok" if worst <= 1e-5 else "mismatch"
                rec["max_rowwise_rel_diff_vs_forward_many"] = worst
                ts = []
                for i in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    model.forward_batches(sub_d, b, streams=2)
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                rec["ms_per_fragment"] = float(np.median(ts[2:])) / (2 * b)
                if rec["verdict"] == "ok" and (best is None or rec["ms_per_fragment"] < out["batched_by_size"][str(best)]["ms_per_fragment"]):
                    best = b
            except Exception as ex:      # noqa: BLE001
                rec["verdict"] = f"failed: {ex!r}"[:200]
            out["batched_by_size"][str(b)] = rec
            try:
                model._graphs.clear()          # free this size's plans before the next one
            except Exception:      # noqa: BLE001
                break
        if best is not None:
            out["batched"], out["B"] = "ok", best
            out["max_rowwise_rel_diff_vs_forward_many"] = out["batched_by_size"][str(best)]["max_rowwise_rel_diff_vs_forward_many"]
            out["batched_ms_per_fragment"] = out["batched_by_size"][str(best)]["ms_per_fragment"]
        else:
            out["batched"] = "; ".join(f"B={k}: {v['verdict']}" for k, v in out["batched_by_size"].items())
    print(json.dumps(out))


def run_probe(args, variant="", sizes=""):
    """One probe subprocess -> (dict or None, note).  sizes: batch sizes of the batched plan to probe (default: --streams and half of it)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--probe-batched", "--config", args.config, "--streams", str(args.streams)]
    if sizes:
        cmd += ["--probe-sizes", sizes]
    env = dict(os.environ)
    env.pop("IMFNET_B200_VARIANT", None)
    if variant:
        env["IMFNET_B200_VARIANT"] = variant
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    except subprocess.TimeoutExpired:
        return None, "probe timed out"
    line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), "")
    try:
        d = json.loads(line)
    except ValueError:
        d = {}
    if r.returncode == 0 and d.get("probe") == "done":
        return d, "ok"
    tail = (r.stderr.strip().splitlines() or ["no stderr"])[-1][:200]
    return None, f"probe failed (rc {r.returncode}): {tail}"


def select_modes(args):
    """Automatic mode (--batched -1): decide, from subprocess probes on the GPU this rank is about to measure,
      * whether to load one of the kernel variant libraries (imfnet_b200/build.py AUTO_VARIANTS): the fastest one whose descriptors
        are bit-identical to the default library's and whose step is at least 2 % shorter;
      * whether to time the batched captured plan as a second execution mode: only if it reproduces forward_many.
    Any failure of a probe only costs time: the default library and execution mode are what is measured then.
    Returns (batched B or 0, note)."""
    from imfnet_b200.build import AUTO_VARIANTS, lib_path
    # the decision (not any measurement) is remembered per box, GPU index and build, so that the back-to-back runs of a scaling sweep
    # (N = 1, 2, 4, 8 on the same box) probe once
    import hashlib
    libs = [lib_path(v) for v in [""] + AUTO_VARIANTS] + [os.path.abspath(__file__)]
    stamp = [(os.path.basename(f), os.path.getsize(f), int(os.path.getmtime(f))) for f in libs if os.path.exists(f)]
    try:
        boot = open("/proc/sys/kernel/random/boot_id").read().strip()      # a decision never travels to another machine
    except OSError:
        boot = str(os.getpid())
    key = hashlib.sha1(json.dumps([stamp, boot, args.config, args.streams, os.environ.get("LOCAL_RANK", "0"), args.variant_probe]).encode()).hexdigest()[:16]
    cache = os.path.join(CACHE_DIR, f"modes_{key}.json")
    if os.environ.get("IMFNET_B200_VARIANT", "") == "":
        try:
            if time.time() - os.path.getmtime(cache) < 3600:
                c = json.load(open(cache))
                if c["variant"]:
                    os.environ["IMFNET_B200_VARIANT"] = c["variant"]
                args.probe_table = c.get("table")
                return int(c["batched"]), c["note"] + " [decision cached by an earlier bench.py run on this box]"
        except (OSError, ValueError, KeyError):
            pass

    def remember(batched, variant, note):
        try:
            os.makedirs(CACHE_DIR, exist_ok=True)
            with open(cache, "w") as f:
                json.dump({"batched": batched, "variant": variant, "note": note, "table": getattr(args, "probe_table", None)}, f)
        except OSError:
            pass
        return batched, note

    def row(d):
        return {k: d.get(k) for k in ("seq_ms_per_step", "conv64_us_per_launch", "conv32_us_per_launch", "conv64_us_by_debug_flags", "batched", "batched_by_size", "B")}

    d0, n0 = run_probe(args)
    if d0 is None:
        return 0, f"default library: {n0} (batched plan and kernel variants not used)"
    chosen, best_name, notes = d0, "", []
    args.probe_table = {"default": row(d0)}
    if os.environ.get("IMFNET_B200_VARIANT", "") == "" and args.variant_probe:
        sizes = str(d0["B"]) if d0.get("batched") == "ok" else ""      # the variants only probe the batch size the default library preferred
        for name in AUTO_VARIANTS:
            dx, nx = run_probe(args, name, sizes)
            args.probe_table[name] = {"probe": nx} if dx is None else dict(row(dx), bit_identical=dx["hashes"] == d0["hashes"])
            if dx is None:
                notes.append(f"variant {name}: {nx}")
            elif dx["hashes"] != d0["hashes"]:
                notes.append(f"variant {name}: descriptors differ from the default library")
            else:
                notes.append(f"variant {name}: bit-identical, {dx['seq_ms_per_step']:.2f} ms per step")
                if dx["seq_ms_per_step"] < 0.98 * d0["seq_ms_per_step"] and dx["seq_ms_per_step"] < chosen["seq_ms_per_step"]:
                    chosen, best_name = dx, name
        if best_name:
            os.environ["IMFNET_B200_VARIANT"] = best_name
    note = f"default library {d0['seq_ms_per_step']:.2f} ms per step; " + "; ".join(notes) + (f"; variant {best_name} in use; " if best_name else "; default library in use; ")
    if chosen.get("batched") == "ok":
        return remember(int(chosen["B"]), best_name, note + "batched plan: probe ok on this GPU (max row-wise rel diff vs forward_many "
                        f"{chosen['max_rowwise_rel_diff_vs_forward_many']:.1e})")
    return remember(0, best_name, note + f"batched plan: probe {chosen.get('batched', 'no result')} (not used)")


def run_ours(args, rank, world, local_rank):
    import imfnet_b200.me as ME
    from imfnet_b200 import _lib, load_model, synthetic
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = _lib.lib()       # raises if the CUDA extension is missing
    frags, (target, voxel, W, H) = make_inputs(args.config, rank)
    sd = synthetic.make_state_dict(0)
    model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(dev)
    dev_frags = [(c.to(dev), f.to(dev), im.to(dev)) for c, f, im in frags]
    pin_frags = [(c.pin_memory(), f.pin_memory(), im.pin_memory()) for c, f, im in frags]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        flush_buf.fill_(1)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    K = max(1, args.streams)      # independent fragments in flight per step (one captured plan + CUDA stream each)
    if world > 1:                 # every rank must time the same modes (their barriers are collectives): batched only if all probes passed
        agree = torch.tensor([args.batched], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(agree, op=torch.distributed.ReduceOp.MIN)
        if int(agree.item()) != args.batched:
            args.batched_note += "; another rank's probe failed"
        args.batched = int(agree.item())
    Bt = max(0, args.batched)     # > 0: the batched captured plan is timed too (groups of Bt fragments per graph replay, two groups per step)

    def make_steps(batch):
        """(fragments per step, resident step, end-to-end step) of one execution mode: batch = 0 -> K single-fragment plans in
        flight (forward_many / forward_many_host); batch = B -> two groups of B fragments, one BatchGraphPlan replay each."""
        kf = 2 * batch if batch else K
        host_out = torch.empty((target, 32), dtype=torch.float32).pin_memory()
        host_outs = [torch.empty((target, 32), dtype=torch.float32).pin_memory() for _ in range(kf)]

        def pick(frs, i):
            return [frs[(i * kf + j) % N_FRAGMENTS] for j in range(kf)]

        def resident(i):
            if batch:
                return model.forward_batches(pick(dev_frags, i), batch, streams=2)
            if K == 1:
                c, f, im = dev_frags[i % N_FRAGMENTS]
                return model(ME.SparseTensor(f, coordinates=c), im).F
            return [o.F for o in model.forward_many([(ME.SparseTensor(f, coordinates=c), im) for c, f, im in pick(dev_frags, i)], streams=K)]

        def e2e(i):
            if batch:
                return model.forward_batches(pick(pin_frags, i), batch, streams=2, out=host_outs)
            if K == 1:
                c, f, im = pin_frags[i % N_FRAGMENTS]
                x = ME.SparseTensor(f.to(dev, non_blocking=True), coordinates=c.to(dev, non_blocking=True))
                out = model(x, im.to(dev, non_blocking=True)).F
                host_out.copy_(out, non_blocking=True)
                return out
            # the public end-to-end call: pinned host fragments in, pinned host descriptors out (copies ride the plans' streams)
            return model.forward_many_host(pick(pin_frags, i), streams=K, out=host_outs)

        return kf, resident, e2e

    def timed(step_fn):
        """Returns (ms, launches, clocks, wall), or the exception when the mode failed; the barriers are executed either way (N > 1)."""
        ok = True
        try:
            for i in range(args.warmup):
                step_fn(i)
        except Exception as ex:      # noqa: BLE001  (only the optional batched mode may fail; the caller re-raises for the default one)
            ok = ex
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        from imfnet_b200.engine import GraphPlan
        l0 = L.imf_launch_count() + GraphPlan.replayed_launches
        ms = 0.0
        wall0 = time.perf_counter()
        for i in range(args.steps):
            if ok is not True:
                break
            try:
                flush()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step_fn(args.warmup + i)
                e1.record()
                e1.synchronize()
                ms += e0.elapsed_time(e1)
            except Exception as ex:      # noqa: BLE001
                ok = ex
        barrier()
        wall = time.perf_counter() - wall0
        launches = L.imf_launch_count() + GraphPlan.replayed_launches - l0
        clocks = sampler.stop()
        if ok is not True:
            return ok
        return ms, launches, clocks, wall

    K_seq, step_resident, step_e2e = make_steps(0)
    if args.profile:      # under ncu: just the resident steps, nothing else
        with torch.no_grad():
            for i in range(args.warmup + args.steps):
                step_resident(i)
            torch.cuda.synchronize()
        return
    modes = {}
    with torch.no_grad():
        r, r2 = timed(step_resident), timed(step_e2e)
        for x in (r, r2):
            if isinstance(x, Exception):
                raise x
        modes["single-fragment plans"] = dict(k=K_seq, ms=r[0], launches=r[1], clocks=r[2], wall=r[3], ms_e2e=r2[0])
        roof = dominant_kernel_roofline(model, frags[0], flush) if rank == 0 else None
        # latency of ONE fragment on an otherwise idle GPU (the reference's own usage pattern: scripts/generate_desc.py:65-123)
        lat = []
        for i in range(13):
            c, f, im = dev_frags[i % N_FRAGMENTS]
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(ME.SparseTensor(f, coordinates=c), im)
            e1.record()
            e1.synchronize()
            if i >= 3:
                lat.append(e0.elapsed_time(e1))
        latency_ms = float(np.median(lat))
        # the same with the model's low-latency setting (small levels split over more CTAs; plans are rebuilt for it)
        model.low_latency = True
        lat = []
        for i in range(13):
            c, f, im = dev_frags[i % N_FRAGMENTS]
            flush()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model(ME.SparseTensor(f, coordinates=c), im)
            e1.record()
            e1.synchronize()
            if i >= 3:
                lat.append(e0.elapsed_time(e1))
        latency_ll_ms = float(np.median(lat))
        model.low_latency = False
        # the batched captured plan, timed LAST so that nothing measured above depends on it; any failure leaves the default mode
        if Bt:
            try:
                K_b, b_res, b_e2e = make_steps(Bt)
                r, r2 = timed(b_res), timed(b_e2e)
                if isinstance(r, Exception) or isinstance(r2, Exception):
                    args.batched_note += f"; batched timing failed: {r if isinstance(r, Exception) else r2!r}"[:300]
                else:
                    modes["batched plan"] = dict(k=K_b, ms=r[0], launches=r[1], clocks=r[2], wall=r[3], ms_e2e=r2[0])
            except Exception as ex:      # noqa: BLE001
                args.batched_note += f"; batched timing failed: {ex!r}"[:300]
        # headline = the faster execution mode of the public API (by end-to-end throughput); every timed mode is reported in config
        best = max(modes, key=lambda n: modes[n]["k"] / modes[n]["ms_e2e"])
        mb = modes[best]
        K, ms_rank, launches, clocks, wall, ms_rank_e2e = mb["k"], mb["ms"], mb["launches"], mb["clocks"], mb["wall"], mb["ms_e2e"]

    # per-rank timing records: the one collective of this path
    from imfnet_b200.pipeline import aggregate_throughput, gather_records
    rec_dev = dev if world > 1 else "cpu"
    rec = torch.tensor([rank, target * args.steps * K, ms_rank], dtype=torch.float64, device=rec_dev)
    rec_e2e = torch.tensor([rank, target * args.steps * K, ms_rank_e2e], dtype=torch.float64, device=rec_dev)
    allrec, allrec_e2e = gather_records(rec, world), gather_records(rec_e2e, world)
    if rank != 0:
        return
    value, ms_total = aggregate_throughput(allrec.cpu())
    e2e_v, ms_e2e = aggregate_throughput(allrec_e2e.cpu())

    cpu = None
    if world == 1:
        from oracle import imfnet_oracle
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        c, f, im = frags[0]
        ts = []
        t_budget = time.perf_counter()
        while len(ts) < 5 and time.perf_counter() - t_budget < 25:
            t0 = time.perf_counter()
            imfnet_oracle.forward(sd, c, f, im)
            ts.append(time.perf_counter() - t0)
        cpu = {"value": target / float(np.median(ts)), "unit": "voxels/s", "cores": cores, "kind": "port",
               "sample": f"median of {len(ts)} cold oracle forwards of one {target}-voxel fragment (same weights), fp32 torch CPU"}

    h2d = K * (target * 16 + target * 4 + 3 * H * W * 4)
    d2h = K * target * 32 * 4
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {target}-voxel synthetic 3DMatch fragment + {W}x{H} image, ResUNetBN2C 32-D descriptors",
                   "fragments_per_step": K, "streams": 2 if best == "batched plan" else K, "single_fragment_latency_ms": latency_ms,
                   "single_fragment_latency_ms_low_latency_setting": latency_ll_ms,
                   "small_level_splits": "off (model.low_latency=False: throughput setting, used for value and e2e)",
                   "distinct_fragments_per_rank": N_FRAGMENTS,
                   "l2": "flushed between steps (256 MiB write, outside the per-step CUDA events)",
                   "timing": "sum of per-step CUDA-event times on the launching stream, max over ranks",
                   "coordinate_maps": "rebuilt every step (cold), as the reference does per SparseTensor",
                   "execution": (f"one captured CUDA graph replay per batch of {Bt} fragments (device-side sizes), 2 batches per step, 2 plans in flight"
                                 if best == "batched plan" else
                                 f"one captured CUDA graph replay per fragment (device-side sizes), {K} independent fragments in flight on {K} streams"
                                 if model.use_cuda_graph else "eager launches"),
                   "execution_modes_timed": {n: {"fragments_per_step": m["k"], "voxels_per_s": target * args.steps * m["k"] / (m["ms"] * 1e-3),
                                                 "voxels_per_s_e2e": target * args.steps * m["k"] / (m["ms_e2e"] * 1e-3)}
                                             for n, m in modes.items()},
                   "mode_selection": args.batched_note,
                   "mode_probes": getattr(args, "probe_table", None),
                   "library": os.path.basename(_lib.lib_path()),
                   "parallelism": f"fragments sharded over {world} GPU(s), no data-path collective"},
        "e2e": {"value": e2e_v, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        "wall_s_timed_region": wall,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--streams", type=int, default=10, help="independent fragments in flight per step (captured plan + stream each)")
    ap.add_argument("--batched", type=int, default=-1,
                    help="B > 0: also time the batched captured plan (imfnet_b200/batched.py: groups of B fragments per graph replay) and "
                         "report the faster mode; 0: off; -1 (default): B = --streams if a parity probe of that path passes in a "
                         "subprocess on this GPU, else off")
    ap.add_argument("--probe-batched", action="store_true", help="(internal) probe subprocess of the automatic mode; prints one JSON line")
    ap.add_argument("--probe-sizes", default="", help="(internal) batch sizes the probe tries for the batched plan, comma separated")
    ap.add_argument("--no-variant-probe", dest="variant_probe", action="store_false",
                    help="automatic mode: do not try the kernel variant library (imfnet_b200/build.py VARIANTS)")
    ap.add_argument("--profile", action="store_true", help="run only warm-up + steps (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "ours" and not args.profile:
        args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.probe_batched:
        probe_batched(args, local_rank)
        return
    args.batched_note = "off"
    if args.batched > 0:
        args.batched_note = f"forced (--batched {args.batched})"
    elif args.batched < 0 and not args.profile:
        args.batched, args.batched_note = select_modes(args)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL may print its version banner on stdout while the communicator is created; stdout must carry exactly one JSON line,
        # so the descriptor is pointed at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.distributed.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
