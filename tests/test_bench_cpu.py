"""bench.py's CPU-side pieces (no GPU): workload definitions of BASELINE.json configs 3-5 and the reference arm."""
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_workload_params_follow_the_configs():
    bench = _bench()
    import lfd_b200
    pb0, pd0, _ = lfd_b200.default_params()
    name, pb, pd = bench.workload_params(types.SimpleNamespace(workload="config3", hough_method=1.0))
    assert name == "config3-camcol-mix" and pd["houghMethod"] == pd0["houghMethod"] == 20 and pd["dilateKernel"].shape == (9, 9)
    name, pb, pd = bench.workload_params(types.SimpleNamespace(workload="config4", hough_method=2.0))
    assert "config4" in name and pd["dilateKernel"].shape == (15, 15) and pd["houghMethod"] == 2.0
    assert pb["houghMethod"] == pb0["houghMethod"] and pd["erodeKernel"].shape == (3, 3)


def test_config5_grid_and_images():
    bench = _bench()
    cases = bench.c5_cases()
    assert len(cases) == 21 and len(bench.c5_cases(quick=True)) == 4
    assert {c[0] for c in cases} == set(bench.C5_DENSITIES) and {c[1] for c in cases} == set(bench.C5_RHOS)
    assert {round(np.pi / c[2]) for c in cases} == {180, 360, 720, 1440}
    img = bench.c5_image(0.01, 7)
    assert img.shape == (4096, 4096) and img.dtype == np.uint8 and 0.008 < (img != 0).mean() < 0.013
    assert np.array_equal(img, bench.c5_image(0.01, 7))


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU; a tiny run must print exactly
    one JSON line with the contract's keys."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"] == "config3-camcol-mix"
