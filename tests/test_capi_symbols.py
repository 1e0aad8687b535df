"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/miqp_b200.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

import planner_miqp_b200 as P
from conftest import ROOT, has_gpu


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "miqp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(miqp_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(P.DECLARED_SYMBOLS) == declared
    assert sorted(P.exported_symbols()) == declared


def test_layout_is_host_only(testcase_problem):
    from planner_miqp_b200.capi import layout
    from oracle import oracle as O
    a, b = layout(testcase_problem), O.layout(testcase_problem)
    for n, _ in a._fields_:
        assert getattr(a, n) == getattr(b, n), n


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(P.MiqpB200Error):
        P.Solver()
