"""CPU checks of the drop-in boundary: libjpezy_b200.so loads, exports every symbol include/jpezy_b200.h
declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import jpezy_b200 as J
from jpezy_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            text = open(os.path.join(ROOT, "include", fn)).read()
            names |= set(re.findall(r"JPEZYB200_API[^;(]*?\b(jpezyb200_\w+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_expected_entry_points():
    d = declared_symbols()
    assert len(d) >= 18
    for must in ("jpezyb200_encode", "jpezyb200_decode", "jpezyb200_ctx_create", "jpezyb200_strerror"):
        assert must in d
    assert set(d) == set(capi.EXPORTS), "python binding and header disagree: %s" % (set(d) ^ set(capi.EXPORTS))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.library_path())
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert J.abi_version() == 1


def test_strerror_and_frame_helpers():
    L = capi.load_library()
    assert L.jpezyb200_strerror(0) == b"ok"
    assert b"no CPU path" in L.jpezyb200_strerror(capi.ENODEVICE)
    f = J.default_frame(1920, 1080)
    assert (f.width, f.height, f.ncomp) == (1920, 1080, 3)
    assert list(f.hs) == [2, 1, 1] and list(f.vs) == [2, 1, 1] and list(f.tq) == [0, 1, 1]
    # src/decoder/jpezy_decoder.hpp:94-101: 1080 rows -> 68 MCU rows -> 1088 padded rows
    assert J.plane_bytes(f) == 1920 * 1088
    assert J.plane_bytes(J.default_frame(17, 33)) == 32 * 48
    assert list(f.qt[0])[:4] == [16, 11, 10, 16] and list(f.qt[1])[:4] == [17, 18, 24, 47]
    assert sum(f.ht[1][0].bits) == 162 and sum(f.ht[0][1].bits) == 12


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(J.JpezyError) as e:
        J.Context(0)
    assert e.value.code == capi.ENODEVICE


def test_product_package_does_not_touch_the_oracle():
    # the oracle is test infrastructure: nothing under jpezy_b200/ or include/ may import, link or call it
    bad = []
    for base in ("jpezy_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc", "Makefile")):
                    t = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"\boracle\b|liboracle|orc_", t):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
