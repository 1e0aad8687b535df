"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the oracle against the reference's golden vectors, the host logic
and the C-ABI symbol table; `-m gpu` tests are the parity tests proper (CUDA path through
the C-ABI vs. the oracle) and run on a B200.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def testcase_problem():
    from oracle.dat_io import read_dat
    return read_dat(os.path.join(GOLDEN, "cplexmodel_testcase.dat"))


@pytest.fixture(scope="session")
def sos_problem():
    from oracle.dat_io import read_dat
    return read_dat(os.path.join(GOLDEN, "test_sos.dat"))


@pytest.fixture(scope="session")
def golden_solution():
    with open(os.path.join(GOLDEN, "testcase_cplex_solution.json")) as f:
        return json.load(f)


def golden_vector(p, g):
    """Full column vector from the reference's pinned RawResults
    (test/cplex_wrapper_test.cc:283-457).  Eigen's setValues takes nested initializer lists in
    logical index order, so the flattened numbers are row-major over the logical shape."""
    from oracle import oracle as O
    lay = O.layout(p)
    x = np.zeros(lay.ncols)
    v = O.block_views(p, x)
    C, N, R, Ob, L, E = p.C, p.N, p.R, p.O, p.L, p.E

    def colmajor(vals, shape):
        return np.array(vals, dtype=np.float64).reshape(shape)

    for name in O.CORE_BLOCKS:
        v[name][...] = colmajor(g[name], (C, N))
    for name in O.NWE_NAMES:
        v[name][...] = colmajor(g[name], (C, E, N))
    v["active_region"][...] = colmajor(g["active_region"], (C, N, R))
    for name in O.RCNA_NAMES:
        v[name][...] = colmajor(g[name], (C, N))
    v["deltacc"][...] = colmajor(g["deltacc"], (C, Ob, N, L))
    v["deltacc_front"][...] = colmajor(g["deltacc_front"], (C, Ob, N, L, 4))
    return x
