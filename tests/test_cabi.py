"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, exports every symbol the headers declare,
and refuses to run without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import errno
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from longtail_b200 import build
    path = build.build()
    return C.CDLL(path)


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        names.update(re.findall(r"LT_B200_EXPORT\s+[^;(]*?\b((?:lt_b200|Longtail_B200|Longtail_CreateB200)\w*)\s*\(", text))
    return sorted(names)


def test_headers_declare_something():
    assert len(declared_symbols()) >= 15


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.lt_b200_context_create(0, C.byref(h)) == errno.ENODEV
    import longtail_b200
    with pytest.raises(longtail_b200.LongtailB200Error):
        longtail_b200.Context(0)


def test_product_does_not_reference_oracle():
    """nothing under longtail_b200/ or include/ may import, link or load oracle/ (the oracle is test infrastructure)"""
    bad = []
    for base in ("longtail_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_lib" in dirpath or "__pycache__" in dirpath:
                continue
            for fn in files:
                if fn.endswith((".so", ".o", ".log", ".pyc")):
                    continue
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                if re.search(r"lt_oracle|liblt_oracle|libref_shim|oracle/|oracle_lib", text):
                    bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
