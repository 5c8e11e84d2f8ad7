"""In-tree nvcc build of csrc/*.cu -> imfnet_b200/csrc/libimfnet_b200.so (sm_100a only).

    python -m imfnet_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  Objects are rebuilt only when
their source (or a header) is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libimfnet_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# experiment switches of the kernels (e.g. IMFNET_B200_NVCC_FLAGS="-DIMF_G4_SKIP_CLEAN_ZERO" + --force); empty in the shipped build
FLAGS += os.environ.get("IMFNET_B200_NVCC_FLAGS", "").split()


# Kernel variants: the same sources with experiment switches on some files, linked into libimfnet_b200_<name>.so next to the default
# library (all other objects are shared).  Selected at load time by IMFNET_B200_VARIANT=<name> (imfnet_b200/_lib.py); bench.py only
# does so after a subprocess probe showed bit-identical descriptors and a shorter step on the GPU at hand.
VARIANTS = {"x": {"sparse_conv_g4.cu": ["-DIMF_G4_LEAN_PRODUCER", "-DIMF_G4_VEC_RESIDUAL", "-DIMF_G4_NO_TRACE"],
                  "flash_fusion.cu": ["-DIMF_FLASH_UNIFORM_ISSUE"],
                  "tc_gemm.cu": ["-DIMF_TCGEMM_UNIFORM_ISSUE"]}}
# z = x + zero fills of clean ring rows skipped (fewer LDGSTS wavefronts, more producer instructions: which wins is a measurement)
VARIANTS["z"] = dict(VARIANTS["x"], **{"sparse_conv_g4.cu": VARIANTS["x"]["sparse_conv_g4.cu"] + ["-DIMF_G4_SKIP_CLEAN_ZERO"]})
# y = x + the MMA warps of the convolution kernel pass their turn on before issuing (shorter hand-off chain; same single-owner accumulators)
VARIANTS["y"] = dict(VARIANTS["x"], **{"sparse_conv_g4.cu": VARIANTS["x"]["sparse_conv_g4.cu"] + ["-DIMF_G4_EARLY_TURN"]})
VARIANTS["w"] = dict(VARIANTS["x"], **{"sparse_conv_g4.cu": VARIANTS["z"]["sparse_conv_g4.cu"] + ["-DIMF_G4_EARLY_TURN"]})      # z + y
AUTO_VARIANTS = ["x", "z", "y", "w"]      # what bench.py's automatic mode may load (bit-identical descriptors and a shorter step required)


def lib_path(variant: str = "") -> str:
    return LIB if not variant else os.path.join(CSRC, f"libimfnet_b200_{variant}.so")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "imfnet_b200.h"))
    objs, jobs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o, []))
    variant_objs = {}
    for name, switches in VARIANTS.items():
        vobjs = []
        for src in sources():
            o = os.path.join(CSRC, src[:-3] + ".o")
            if src in switches:
                s = os.path.join(CSRC, src)
                o = os.path.join(CSRC, f"{src[:-3]}_{name}.o")
                if force or _stale(o, [s] + headers):
                    jobs.append((s, o, switches[src]))
            vobjs.append(o)
        variant_objs[name] = vobjs

    def compile_one(job):
        s, o, extra = job
        cmd = [NVCC] + ARCH + FLAGS + extra + ["-I", CSRC, "-I", os.path.join(HERE, "..", "include"), "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {os.path.basename(s)} ---\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    for target, tobjs in [(LIB, objs)] + [(lib_path(n), o) for n, o in variant_objs.items()]:
        if jobs or force or _stale(target, tobjs):
            cmd = [NVCC] + ARCH + ["-shared", "-o", target] + tobjs
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
