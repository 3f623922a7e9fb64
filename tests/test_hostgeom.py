"""CPU check of the *product's* device geometry source (lfd_b200/csrc/geom.cuh) compiled for the host
(tests/hostgeom.cpp) against cv2: hull-from-row-extremes + minAreaRect + boxPoints + quad fill.
The shipped library never runs this code on the CPU; this only shortens the GPU debug loop."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from lfd_b200 import synth
from oracle import ref_pipeline as rp

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hg(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hg") / "libhostgeom.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(HERE, "hostgeom.cpp"), "-o", out], check=True, stderr=subprocess.DEVNULL)
    lib = ctypes.CDLL(out)
    lib.hg_rect.restype = ctypes.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_rects_match_cv2(cv2mod, hg):
    img, _ = synth.make_frame(77, n_stars=150, h=500, w=640, trails=[
        {"p0": (10, 30), "p1": (620, 470), "sigma": 3.0, "peak": 3.0}])
    taps = {}
    rp.dim_pass(img.copy(), taps=taps, **rp.DEFAULT_DIM)
    n = same = npass = 0
    for c in taps["contours"]:
        pts = np.ascontiguousarray(c.reshape(-1, 2), np.int32)
        out = np.zeros(5, np.float32)
        box = np.zeros(8, np.int32)
        hg.hg_rect(_p(pts), len(pts), _p(out), _p(box))
        ref = cv2mod.minAreaRect(c)
        got = ((float(out[0]), float(out[1])), (float(out[2]), float(out[3])), float(out[4]))
        n += 1
        same += got == ref
        w, h = ref[1]
        L, Wd = (w, h) if w > h else (h, w)
        if L > 1 and Wd > 1 and L / Wd > 5:
            npass += 1
            assert got == ref
            assert np.array_equal(box.reshape(4, 2), np.asarray(cv2mod.boxPoints(ref), np.int32))
    assert n > 50 and npass >= 1
    assert n - same <= max(1, n // 100), (same, n)


def test_fill_quad_matches_cv2(cv2mod, hg):
    rng = np.random.default_rng(8)
    H, W = 180, 240
    for k in range(1500):
        if k % 3 == 0:
            box = rng.integers(-60, 300, (4, 2)).astype(np.int32)   # arbitrary (even self-crossing) quads
        else:
            rect = ((float(rng.uniform(-30, W + 30)), float(rng.uniform(-30, H + 30))),
                    (float(rng.uniform(0.5, 200)), float(rng.uniform(0.5, 30))), float(rng.uniform(-90, 0)))
            box = np.asarray(cv2mod.boxPoints(rect), np.int32)
        a = np.zeros((H, W), np.uint8)
        b = np.zeros((H, W), np.uint8)
        cv2mod.fillPoly(a, [box], (255, 255, 255))
        hg.hg_fill_quad(_p(b), H, W, _p(np.ascontiguousarray(box)))
        assert np.array_equal(a, b), (k, box.tolist())
