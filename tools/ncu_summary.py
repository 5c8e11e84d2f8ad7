"""Summarise an ncu report (.ncu-rep, `ncu --set full`) per launch: duration, DRAM bytes and throughput, tensor-pipe and SM activity,
registers, shared memory -- the numbers DESIGN.md / bench.py quote.  Runs where ncu is installed (no GPU needed):

    python tools/ncu_summary.py gpurun_out/<tag>/g4.ncu-rep [--traffic-key 'k_sparse_conv_g4<64,64>@batched' --pick longest:k_sparse_conv_g4<64, 64>]

--traffic-key K --pick longest:<kernel substring> | rank:<n>:<kernel substring>: also records dram bytes of the (n-th) longest launch whose name contains the substring
in profiles/ncu_traffic.json under key K (bench.py's roofline.traffic reads that file)."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "MB_rd", 1e-6),
    ("dram__bytes_write.sum", "MB_wr", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2hit%", 1.0),
    ("lts__t_bytes.sum", "L2_MB", 1e-6),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1.0),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__grid_size", "grid", 1.0),
]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "byte": 1.0, "Kbyte": 1e3,
        "Mbyte": 1e6, "Gbyte": 1e9}


def load(path):
    if path.endswith(".csv"):          # already exported with `ncu -i ... --page raw --csv`
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    recs = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {}
        for n, u, v in zip(names, units, r):
            if n in ("Kernel Name", "ID"):
                d[n] = v
                continue
            try:
                d[n] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                pass
        recs.append(d)
    return recs


def main():
    path = sys.argv[1]
    recs = load(path)
    print(f"{os.path.basename(path)}: {len(recs)} launches (ncu --set full: cold caches, serialised; times are NOT bench values)")
    print("  " + " ".join(f"{lab:>8s}" for _m, lab, _s in METRICS) + "  kernel")
    for d in recs:
        vals = []
        for m, lab, sc in METRICS:
            v = d.get(m)
            vals.append(f"{v * sc:8.1f}" if v is not None else "       -")
        print("  " + " ".join(vals) + "  " + d["Kernel Name"].split("(")[0][-60:])
    if "--traffic-key" in sys.argv:
        key = sys.argv[sys.argv.index("--traffic-key") + 1]
        pick = sys.argv[sys.argv.index("--pick") + 1]          # longest:<substring> | rank:<n>:<substring> (n-th longest, 1-based)
        rank, sub = (1, pick.split(":", 1)[1]) if pick.startswith("longest:") else (int(pick.split(":", 2)[1]), pick.split(":", 2)[2])
        cand = sorted((d for d in recs if sub in d["Kernel Name"]), key=lambda d: -d.get("gpu__time_duration.sum", 0))
        best = cand[rank - 1]
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur[key] = {"dram_bytes_per_launch": int(best["dram__bytes_read.sum"] + best["dram__bytes_write.sum"]),
                    "ncu_duration_us": best["gpu__time_duration.sum"] * 1e-3,
                    "tensor_pipe_pct": best.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "source": f"dram__bytes_read.sum + dram__bytes_write.sum of the {'longest' if rank == 1 else f'number-{rank} (by duration)'} {sub} launch in one ncu --set full capture of "
                              f"`bench.py --profile` ({os.path.basename(path)}; summary under profiles/)"}
        json.dump(cur, open(p, "w"), indent=1)
        print(f"recorded {key}: {cur[key]}")


if __name__ == "__main__":
    main()
