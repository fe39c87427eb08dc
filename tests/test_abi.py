"""CPU test: libpgb200.so loads and exports every symbol include/pgb200.h declares; no compute call is made."""
import ctypes
import os

import pytest

from pilotguru_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.declared_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/pgb200.h but not exported: {missing}"


def test_binding_table_covers_header():
    declared = set(_lib.declared_symbols())
    bound = set(_lib._SIGS)
    assert bound <= declared, bound - declared
    assert declared <= bound, declared - bound


def test_no_gpu_fails_loudly():
    """Without a usable sm_100 device the product must raise, never fall back to CPU code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pilotguru_b200 import PgbError
    from pilotguru_b200.orb import ORBextractor
    with pytest.raises(PgbError):
        ORBextractor(1000, 1.2, 8, 20, 7)


def test_product_does_not_reference_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "pilotguru_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                for needle in ("oracle/", "oracle_lib", "liboracle", "pgo_", "pgo.h"):
                    assert needle not in src, f"{f} references the test oracle ({needle})"


def test_binding_arity_matches_header():
    """Every ctypes signature lists exactly as many arguments as include/pgb200.h declares (a wrong count would not fail
    at load time, it would corrupt the call)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "pgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    seen = {}
    for m in re.finditer(r"\b(pgb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), " ".join(m.group(2).split())
        seen[name] = 0 if params in ("", "void") else params.count(",") + 1
    assert set(seen) == set(_lib._SIGS)
    wrong = {n: (len(_lib._SIGS[n][1]), seen[n]) for n in seen if len(_lib._SIGS[n][1]) != seen[n]}
    assert not wrong, f"(bound, declared) argument counts differ: {wrong}"
