"""bench.py contract (CPU part): the reference arm prints exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "0", "--batch", "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MIQP plans/sec at 1e-4 gap" and d["unit"] == "plans/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_device_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--batch", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_both_arms_share_one_config_object_and_every_workload_refuses_without_gpu():
    """`config` must be identical in the GPU arm and the reference arm (the driver compares them); what differs per run lives in
    `run`.  The multi-agent workloads have no CPU fallback either."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("_bench_mod", os.path.join(ROOT, "bench.py"))
    # bench.py redirects fd 1 at import: load it in a child process instead and ask for the config there
    code = ("import json, os, sys, importlib.util; "
            f"spec = importlib.util.spec_from_file_location('b', {os.path.join(ROOT, 'bench.py')!r}); "
            "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b); "
            "h = b.gap_histogram([type('I', (), dict(status=0, gap=g))() for g in (0.0, 5e-5, 5e-4, 0.05, 0.5, 3.0)] + "
            "[type('I', (), dict(status=1, gap=float('nan')))()], 1e-4); "
            "b.emit({'config': b.bench_config(2048), 'hist': h, 'workloads': sorted(b.WORKLOADS)})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert set(d["config"]) == {"workload", "plans_per_gpu_per_step", "gap", "seeds", "cache"} and d["config"]["plans_per_gpu_per_step"] == 2048
    assert d["hist"] == {"<=0.0001": 2, "<=1e-3": 1, "<=1e-2": 0, "<=1e-1": 1, "<=1": 1, ">1": 1, "no incumbent": 1}
    assert d["workloads"] == ["config3", "config4", "config5"]
    if not torch.cuda.is_available():
        for wl in ("config3", "config4", "config5"):
            rr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", "1", "--warmup", "0"],
                                capture_output=True, text=True, timeout=600)
            assert rr.returncode != 0 and "no CPU fallback" in (rr.stderr + rr.stdout), wl
