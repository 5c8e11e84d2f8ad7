"""Print the essentials of a bench.py JSON line (value, e2e, roofline, per-layer table)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2f M voxels/s (%.3f ms/step, %d fragments/step)  e2e %.2f M  launches %d  clocks %s" % (
    d["value"] / 1e6, d["ms_per_step"], d["config"]["fragments_per_step"], d["e2e"]["value"] / 1e6, d["gpu_launches"], d.get("clocks")))
print("single-fragment latency %.3f ms; cpu baseline %s" % (d["config"]["single_fragment_latency_ms"], d.get("cpu_baseline")))
r = d.get("roofline")
if r:
    print("dominant: %s\n  %.1f us  %.0f GB/s  frac %.3f  traffic %s" % (r["kernel"], r["ms_per_launch"] * 1e3, r["achieved"], r["frac"], r["traffic"]))
    sp = r["sparse_part"]
    print("sparse part: %.1f MB/fragment  %.3f ms per %d fragments  %.0f GB/s  frac %.3f" % (
        sp["alg_bytes_per_fragment"] / 1e6, sp["ms"], sp["fragments"], sp["achieved"], sp["frac"]))
    for x in r["layers"]:
        print("  %-18s rows %7d pairs %8d %3d->%3d  %8.1f us  %7.1f MB  %6.0f GB/s  %.3f" % (
            x["layer"], x["rows"], x["pairs"], x["cin"], x["cout"], x["us"], x["alg_MB"], x["GBps"], x["frac"]))
