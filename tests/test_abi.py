"""CPU: the C-ABI library loads without a GPU and exports exactly what include/imfnet_b200.h declares."""
import ctypes
import os
import re
import subprocess

from imfnet_b200 import _lib, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    src = open(os.path.join(ROOT, "include", "imfnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(imf_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.lib()
    assert lib.imf_version() >= 100
    assert lib.imf_hash_capacity(50000) == 131072
    assert lib.imf_hash_bytes(1024) == 16384


def test_every_declared_symbol_is_exported_and_bound():
    declared = header_functions()
    assert len(declared) >= 19
    lib = ctypes.CDLL(build.build())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table and header disagree"


def test_no_undeclared_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", build.build()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (imf_[a-z0-9_]+)$", out, flags=re.M)))
    assert exported == header_functions()


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.lib()
    rc = lib.imf_kernel_map(None, None, 10, None, 1000, 3, 1, None, None)    # capacity not a power of two
    assert rc == -1 and b"bad argument" in lib.imf_last_error()
    rc = lib.imf_sparse_conv_fwd(None, 30, None, None, None, 10, 27, 30, 32, None, None, None, 0, 0, None, 32, None)
    assert rc == -1


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
