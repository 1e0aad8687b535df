"""Extract the pinned CPLEX solution vector of cplexmodel_testcase.dat from the
reference's own test (test/cplex_wrapper_test.cc:283-457) into tests/golden/.

Runs only in the build container (needs /root/reference).  The resulting JSON is the
committed fixture; the GPU box never reads /root/reference.

Also copies the two OPL data fixtures that the reference's tests solve
(cplexmodel/cplexmodel_testcase.dat, cplexmodel/test_sos.dat) -- they are input data, the
golden inputs of SURVEY.md section 8(c) items 1-4.
"""
import json
import os
import re
import shutil
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

NUM = re.compile(r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)")


def main():
    os.makedirs(OUT, exist_ok=True)
    src = open(os.path.join(REF, "test", "cplex_wrapper_test.cc")).read()
    start = src.index("RawResults wv;")
    end = src.index("return {mp, wv};")
    body = src[start:end]
    out = {}
    for m in re.finditer(r"wv\.(\w+)\.setValues\(\s*(.*?)\);", body, re.S):
        name, arr = m.group(1), m.group(2)
        out[name] = [float(t) for t in NUM.findall(arr)]
    out["_source"] = "test/cplex_wrapper_test.cc:283-457 (generateTestDataHelper, RawResults wv)"
    out["_pins"] = {
        "source": "test/cplex_wrapper_test.cc:857-876",
        "NrConstraints": 12361, "NonZeroCoefficients": 29834, "NrBinaryVariables": 1240,
        "NrFloatVariables": 340, "objective": 9.57603, "objective_tol": 1e-5, "gap_setting": 0.1,
    }
    with open(os.path.join(OUT, "testcase_cplex_solution.json"), "w") as f:
        json.dump(out, f, indent=0)
    for fn in ("cplexmodel_testcase.dat", "test_sos.dat"):
        shutil.copyfile(os.path.join(REF, "cplexmodel", fn), os.path.join(OUT, fn))
    print({k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    sys.exit(main())
