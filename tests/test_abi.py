"""The C-ABI shared library builds, loads and exports exactly what include/ldot.h declares (no compute calls)."""
import ctypes
import os
import re

from lightningdot_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ldot.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(ldot_[a-z0-9_]+)\s*\(", text))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ldot.h but not exported"


def test_binding_covers_header():
    assert set(_lib.SIGNATURES) == header_symbols()


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.ldot_abi_version() == _lib.ABI_VERSION == 6
    # argument validation happens before any CUDA call, so it is testable without a GPU
    rc = lib.ldot_topk_merge(None, None, 2, 4, 10, 0, 0, None, None, None)
    assert rc == -1
    assert b"null pointer" in lib.ldot_last_error()
    assert lib.ldot_flatip_search_workspace_bytes(10, 1000, 770, 10, 0) == 0   # d not a multiple of 8
    assert b"multiple of 8" in lib.ldot_last_error()
    assert lib.ldot_flatip_search_workspace_bytes(10, 1000, 768, 5000, 0) == 0  # k too large
    need = lib.ldot_flatip_search_workspace_bytes(10000, 1000000, 768, 100, 0)
    assert 0 < need < 4 << 30


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under lightningdot_b200/ (or the dvl/ uniter_model/ horovod/
    drop-in packages) may import it."""
    bad = []
    for pkg in ("lightningdot_b200", "dvl", "uniter_model", "horovod"):
        base = os.path.join(ROOT, pkg)
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
