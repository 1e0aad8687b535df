"""Extracts the fitted polynomial tables (results of the offline MATLAB least-squares fits,
pure data) from the reference header common/parameter/fitting_polynomial_parameters.hpp into
planner-miqp_b200/data/fitting_tables.json.  Runs only in the build container; the JSON is
committed.  Tables are stored row-major [R][3] = coefficients of (1, vx, vy); the header holds
them column-major (Eigen::Map of an R x 3 matrix, fitting_polynomial_parameters.hpp:97-168)."""
import json
import os
import re

REF = "/root/reference/common/parameter/fitting_polynomial_parameters.hpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "planner-miqp_b200", "data", "fitting_tables.json")

src = open(REF).read()
vecs = {}
for m in re.finditer(r"const std::vector<double>\s+(POLY_\w+)\s*=\s*\{(.*?)\};", src, re.S):
    vecs[m.group(1)] = [float(t) for t in re.findall(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)", m.group(2))]
maps = {}
for m in re.finditer(r"(POLY_\w+)_map\[\{(\d+),\s*(\d+),\s*(\d+)\}\]\s*=\s*(POLY_\w+);", src):
    kind, R, vmax, vmin, name = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), m.group(5)
    v = vecs[name]
    assert len(v) == 3 * R, (name, len(v), R)
    rowmajor = [[v[k * R + r] for k in range(3)] for r in range(R)]
    maps.setdefault(f"{R},{vmax},{vmin}", {})[kind] = rowmajor
out = {"source": "common/parameter/fitting_polynomial_parameters.hpp:48-89,195-1276", "tables": maps}
with open(OUT, "w") as f:
    json.dump(out, f)
print({k: sorted(v) for k, v in maps.items()})
