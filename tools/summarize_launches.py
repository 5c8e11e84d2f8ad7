"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total device time).

    python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches]
"""
import collections
import csv
import re
import sys


def main(path, skip=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        rows.append((row["Kernel Name"], v))
    rows = rows[skip:]
    agg, tot = collections.OrderedDict(), 0.0
    for name, v in rows:
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("<unnamed>::", "")[:80]
        a = agg.setdefault(short, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    print(f"{len(rows)} launches, {tot:.1f} us total device time (cold-cache, serialised: compare shares)")
    for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{v:10.1f} us {c:5d}x {100 * v / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
