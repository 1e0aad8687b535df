"""The pybind11 module `miqp` (planner-miqp_b200/host/python_module.cpp) against what the reference's Python tests do
with theirs (python/bindings/tests/python_import_test.py: import; test/py_convexified_map_test.py: decomposition of road
polygons succeeds; python_cplex_wrapper.cpp:19-69: CplexWrapper on a .dat file, SolutionProperties, enums)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "planner-miqp_b200")


@pytest.fixture(scope="module")
def miqp():
    import planner_miqp_b200  # noqa: F401  (builds the libraries if needed)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build_pymodule()
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    import miqp as m
    return m


def area(v):
    x, y = v[:, 0], v[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def is_convex_ccw(v, tol=1e-9):
    n = len(v)
    for k in range(n):
        a, b, c = v[k], v[(k + 1) % n], v[(k + 2) % n]
        if (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]) < -tol:
            return False
    return True


def test_import_enums_and_solution_properties(miqp):
    assert int(miqp.OptimizationStatus.SUCCESS) == 0 and int(miqp.OptimizationStatus.FAILED_TIMEOUT) == 3
    assert miqp.SUCCESS == miqp.OptimizationStatus.SUCCESS                      # export_values, as in the reference
    assert int(miqp.WarmstartType.BOTH_WARMSTART_STRATEGIES) == 3 and int(miqp.ParallelMode.OPPORTUNISTIC) == -1
    sp = miqp.SolutionProperties()
    sp.objective, sp.status, sp.gap, sp.time = 1.5, 101, 0.0, 0.25
    assert (sp.objective, sp.status, sp.gap, sp.time) == (1.5, 101, 0.0, 0.25)


def test_convexified_map_decomposes_an_l_shaped_road(miqp):
    # L-shaped road, 10 m wide arms, given clockwise and closed (the planner accepts both conventions)
    poly = np.array([[0, 0], [0, 40], [10, 40], [10, 10], [50, 10], [50, 0], [0, 0]], dtype=float)
    r = 1.0
    cm = miqp.ConvexifiedMap(None, poly, r, 0.0, 2.0, 1e-9)
    assert cm.HasValidPolygon() and cm.Convert()
    cells = cm.map_convex_polygons
    assert 2 <= len(cells) <= 3                                   # Hertel-Mehlhorn: at most 4x the optimum (2)
    for v in cells.values():
        assert is_convex_ccw(v) and area(v) > 0
    total = sum(area(v) for v in cells.values())
    # the offset region of the L has area 8*38 + 40*8 - the rounded reflex corner; the cells cover it up to that corner piece
    assert 8 * 38 + 40 * 8 - 1.0 <= total <= 8 * 39 + 40 * 8 + 1e-9
    # every cell keeps the collision radius from the boundary of the input polygon
    edges = list(zip(poly[:-1], poly[1:]))
    for v in cells.values():
        for p in v:
            d = min(np.linalg.norm(p - (a + np.clip(np.dot(p - a, b - a) / np.dot(b - a, b - a), 0, 1) * (b - a))) for a, b in edges)
            assert d >= r - 1e-9
    # the cells overlap or touch (a car can pass from one arm into the other): the diagonal between them is not shrunk
    pts = [np.array([5.0, 5.0]), np.array([5.0, 9.5]), np.array([9.0, 5.0])]
    def inside(v, p):
        return all((v[(k + 1) % len(v)][0] - v[k][0]) * (p[1] - v[k][1]) - (v[(k + 1) % len(v)][1] - v[k][1]) * (p[0] - v[k][0]) >= -1e-9 for k in range(len(v)))
    for p in pts:
        assert any(inside(v, p) for v in cells.values())
    # reference along the lower arm only: the upper arm's cell is not selected when it is farther than the buffer
    sel = cm.GetIntersectingConvexPolygons(np.array([[30.0, 5.0], [45.0, 5.0]]))
    assert 1 <= len(sel) <= len(cells) and all(k in cells for k in sel)
    assert np.allclose(cm.map_nonconvex_polygon, poly)


def test_convexified_map_rejects_degenerate_polygons(miqp):
    assert not miqp.ConvexifiedMap(None, np.array([[0, 0], [1, 0], [2, 0]], dtype=float), 0.5, 0.0, 2.0, 1e-9).Convert()
    # narrower than twice the radius: nothing is left
    assert not miqp.ConvexifiedMap(None, np.array([[0, 0], [10, 0], [10, 1], [0, 1]], dtype=float), 1.0, 0.0, 2.0, 1e-9).Convert()
    # a convex polygon is one cell
    cm = miqp.ConvexifiedMap(None, np.array([[0, 0], [10, 0], [10, 6], [0, 6]], dtype=float), 1.0, 0.0, 2.0, 1e-9)
    assert cm.Convert() and len(cm.map_convex_polygons) == 1
    assert np.isclose(area(cm.map_convex_polygons[0]), 8 * 4)


def test_cplex_wrapper_surface_without_device(miqp):
    w = miqp.CplexWrapper("cplexmodel.mod", 12)
    w.setDebugOutputPrint(False)
    w.setDebugOutputFilePrefix("pytest_")
    assert w.getDebugOutputParameterFilePath() == ""
    w.setParameterDatFileAbsolute("/nonexistent/file.dat")
    assert w.callCplex(0.0) == miqp.FAILED_SEG_FAULT              # the solver could not run
    assert "nonexistent" in w.lastError() or w.lastError() != ""


@pytest.mark.gpu
def test_cplex_wrapper_solves_the_reference_fixture(miqp):
    w = miqp.CplexWrapper("cplexmodel.mod", 12)
    w.setParameterDatFileAbsolute(os.path.join(ROOT, "tests", "golden", "cplexmodel_testcase.dat"))
    assert w.callCplex(0.0) == miqp.OptimizationStatus.SUCCESS
    sp = w.getSolutionProperties()
    assert abs(sp.objective - 9.57603) <= 0.1 * 9.57603 and sp.gap <= 0.1 and sp.time > 0    # the file asks for a 10 % gap
    assert w.writeMIPStarts("/tmp/miqp_b200_pybind.mst") and w.readMIPStarts("/tmp/miqp_b200_pybind.mst")
    assert w.exportModel("/tmp/miqp_b200_pybind.lp") and os.path.getsize("/tmp/miqp_b200_pybind.lp") > 100000
