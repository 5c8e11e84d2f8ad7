"""Counts of the Blackwell-specific SASS mnemonics per kernel of the shipped library (cuobjdump -sass of csrc/*.o): tcgen05 MMAs (UTC*MMA),
TMEM loads / stores (LDTM / STTM), tiled TMA loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP), LDGSTS, and the 4-byte vs 16-byte global
stores.  Runs on the build box (no GPU):  python tools/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PAT = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("UBLKCP", r"\bUBLKCP"), ("LDGSTS", r"\bLDGSTS"), ("SYNCS", r"\bSYNCS"), ("STG.128", r"\bSTG\.E\.128"), ("STG.32", r"\bSTG\.E "), ("STS.128", r"\bSTS\.128")]
print("%-16s %-46s" % ("object", "kernel") + "".join("%9s" % n for n, _ in PAT) + "   regs")
for obj in sorted(glob.glob(os.path.join(ROOT, "imfnet_b200", "csrc", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    cur, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                counts[cur][name] += 1
    for fn, c in counts.items():
        dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        dem = re.sub(r"\((int|bool|unsigned int)\)", "", dem)
        dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("<unnamed>::", "")
        print("%-16s %-46s" % (os.path.basename(obj), dem[:46]) + "".join("%9d" % c[n] for n, _ in PAT) + "   %4s" % regs.get(fn, "?"))
    print("%-16s arch: %s" % (os.path.basename(obj), ",".join(sorted(arch))))
