"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import pytest

import wumingpic_b200 as wm
from wumingpic_b200.backend import ABI_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    L = wm.load_library()
    hdr = open(os.path.join(ROOT, "include", "wuming_b200.h")).read()
    declared = set(re.findall(r"\b(wm_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"wm_ctx", "wm_params", "wm_stats"}
    assert declared == set(ABI_SYMBOLS), declared ^ set(ABI_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), f"{s} declared in the header but not exported"


def test_para_range_matches_reference_rule():
    L = wm.load_library()
    for n1, n2, size in [(2, 257, 4), (2, 11, 3), (2, 66, 8), (2, 8, 7)]:
        covered = []
        for r in range(size):
            ns, ne = C.c_int(), C.c_int()
            assert L.wm_para_range(n1, n2, size, r, C.byref(ns), C.byref(ne)) == 0
            assert (ns.value, ne.value) == wm.para_range(n1, n2, size, r)
            covered += list(range(ns.value, ne.value + 1))
        assert covered == list(range(n1, n2 + 1))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wm.WmError) as e:
        wm.Backend(3, 96, 2, 9, 2, 9, 2, 9)
    assert "no CPU fallback" in str(e.value)
