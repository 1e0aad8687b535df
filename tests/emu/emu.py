"""ctypes driver of tests/emu/libmulti_emu.so (host emulation of the multi-car CUDA node code).
TEST INFRASTRUCTURE: checks the device logic against the oracle on the CPU."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
LIB = os.path.join(HERE, "libmulti_emu.so")
SRC = os.path.join(HERE, "multi_emu.cpp")
CSRC = os.path.join(ROOT, "planner-miqp_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("bnb_multi_core.cuh", "node_qp_multi.cuh", "host_pack.hpp",
                                                    "formulation_tables.cuh", "dev_problem.cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-ffp-contract=off",
           "-I/usr/local/cuda/include", SRC, "-o", LIB]
    subprocess.run(cmd, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def solve(p, gap_tol=1e-4, time_limit=60.0, max_nodes=0, verbose=0, warm=None):
    import planner_miqp_b200 as P
    from planner_miqp_b200 import capi
    keep = []
    cp = capi.to_c(p, gap_tol, time_limit, keep)
    lay = capi.layout(p)
    Cn, N = p.C, p.N
    Pn = Cn * (Cn - 1) // 2
    ndec = Cn * N + 5 * Cn * N + 5 * Cn * p.O * N + 4 * Pn * N
    ndec_pad = (ndec + 15) & ~15
    obj, bnd = C.c_double(), C.c_double()
    nodes, iters = C.c_long(), C.c_long()
    traj = np.zeros((Cn, N, 8))
    sig = np.zeros((max(Pn, 1), N, 4))
    dec = np.zeros(ndec_pad, dtype=np.uint8)
    f = lib().emu_multi_solve
    f.restype = C.c_int
    rc = f(C.byref(cp), C.c_double(gap_tol), C.c_double(time_limit), C.c_long(max_nodes), C.byref(obj), C.byref(bnd),
           C.byref(nodes), C.byref(iters), traj.ctypes.data_as(C.c_void_p), sig.ctypes.data_as(C.c_void_p),
           dec.ctypes.data_as(C.c_void_p), C.c_int(verbose),
           None if warm is None else np.ascontiguousarray(warm, dtype=np.float64).ctypes.data_as(C.c_void_p))
    return dict(status=rc, objective=obj.value, bound=bnd.value, nodes=nodes.value, iters=iters.value, traj=traj, sig=sig, dec=dec)
