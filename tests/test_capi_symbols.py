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


def test_every_function_declared_in_the_headers_is_exported():
    """parses include/*.h for function declarations and looks each one up in the shared libraries"""
    import ctypes
    import os
    import re
    from planner_miqp_b200 import planner_capi as PC
    inc = os.path.join(os.path.dirname(__file__), "..", "include")
    solver = P.load_library()
    planner = PC.load_library()
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "miqp_b200.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(miqp_b200_\w+)\s*\(", text)))
    assert len(names) >= 16
    for n in names:
        assert hasattr(solver, n), n
    assert sorted(names) == sorted(P.DECLARED_SYMBOLS)
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "miqp_planner_c_api.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(\w+CMiqpPlan\w*|GetCollisionRadius)\s*\(", text)))
    assert len(names) == 20, names
    for n in names:
        assert hasattr(planner, n), n
