"""Pin the oracle (oracle/ref_pipeline.py) against golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py ran /root/reference's processfield / DetectTrails in the build container)."""
import hashlib
import os

import numpy as np
import pytest

from lfd_b200 import synth
from oracle import ref_pipeline as rp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_frame(g):
    peak = float(g["peak"])
    trails = [] if peak == 0 else [{"p0": (10, 20), "p1": (400, 270), "sigma": 2.5, "peak": peak}]
    img, _ = synth.make_frame(int(g["seed"]), n_stars=int(g["nstars"]), h=300, w=420, trails=trails)
    return img


def enc(r):
    return [int(r[0])] + ([r[1]["x1"], r[1]["y1"], r[1]["x2"], r[1]["y2"]] if r[0] else [0, 0, 0, 0])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_small_taps(cv2mod, tag):
    g = np.load(os.path.join(GOLD, "golden_small_%s.npz" % tag))
    assert str(g["cv2_version"]) == cv2mod.__version__, "goldens were generated with another cv2"
    work = np.ascontiguousarray(small_frame(g)[::-1])
    tb, td = {}, {}
    rb = rp.bright_pass(work, taps=tb, **rp.DEFAULT_BRIGHT)
    assert np.array_equal(work, g["clipped_bright"])
    rd = rp.dim_pass(work, taps=td, **rp.DEFAULT_DIM)
    assert np.array_equal(work, g["clipped_dim"])
    assert enc(rb) == g["ret_bright"].tolist() and enc(rd) == g["ret_dim"].tolist()
    assert np.array_equal(tb["equ"], g["1equBRIGHT"])
    assert np.array_equal(tb["morph"], g["2dilateBRIGHT"])
    assert np.array_equal(tb["box_img"], g["3contoursBRIGHT"])
    assert np.array_equal(td["equ"], g["6equDIM"])
    assert np.array_equal(td["eroded"], g["7erodedDIM"])
    assert np.array_equal(td["morph"], g["8openedDIM"])
    assert np.array_equal(td["box_img"], g["9contoursDIM"])


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_full_frame_hashes(cv2mod):
    g = np.load(os.path.join(GOLD, "golden_full.npz"))
    kind, seed = "trail", 1234     # one full-size frame keeps the CPU suite short; the GPU suite checks all
    img, _ = synth.make_case(kind, seed)
    work = np.ascontiguousarray(img[::-1])
    tb, td = {}, {}
    rb = rp.bright_pass(work, taps=tb, **rp.DEFAULT_BRIGHT)
    assert np.array_equal(sha(work), g["%s_%d_clipped_bright_sha1" % (kind, seed)])
    rd = rp.dim_pass(work, taps=td, **rp.DEFAULT_DIM)
    pre = "%s_%d_" % (kind, seed)
    assert enc(rb) == g[pre + "ret_bright"].tolist() and enc(rd) == g[pre + "ret_dim"].tolist()
    for name, arr in (("1equBRIGHT", tb["equ"]), ("2dilateBRIGHT", tb["morph"]), ("3contoursBRIGHT", tb["box_img"]),
                      ("6equDIM", td["equ"]), ("7erodedDIM", td["eroded"]), ("8openedDIM", td["morph"]),
                      ("9contoursDIM", td["box_img"])):
        assert np.array_equal(sha(arr), g[pre + name + "_sha1"]), name


def test_results_txt_of_reference_run(cv2mod, tmp_path):
    """oracle process_frame + result_line reproduce the reference's results.txt byte for byte."""
    g = np.load(os.path.join(GOLD, "golden_run.npz"))
    kinds = {("r", 100): "trail", ("r", 101): "sparse", ("r", 102): "satellite", ("g", 100): "dense_trail",
             ("g", 101): "sparse", ("g", 102): "empty"}
    tree = synth.write_sdss_tree(str(tmp_path), 2888, 1, [100, 101, 102], filters=("r", "g"), kinds=kinds,
                                 startfield=100, endfield=103)
    from lfd_b200 import fitsio_lite
    out = ""
    for flt in ("r", "g"):
        for field in (100, 101, 102):
            path = os.path.join(tree["photoobjpath"], "frames", "301", "2888", "1", "frame-%s-002888-1-%04d.fits" % (flt, field))
            img = fitsio_lite.read(path)
            hdr = fitsio_lite.read_header(path)
            cat = fitsio_lite.read(os.path.join(tree["photoobjpath"], "301", "2888", "1", "photoObj-002888-1-%04d.fits" % field), header=True)[0]
            det, _p, res = rp.process_frame(img, {k: cat[k] for k in cat.dtype.names}, flt)
            if det:
                out += rp.result_line(2888, 1, flt, field, hdr, res)
    assert out == str(g["results_txt"])


def _param_sets():
    ones = lambda a, b: np.ones((a, b), np.uint8)   # noqa: E731  (same sets as oracle/gen_golden.py::param_sets)
    return [("config4_dilate", {"dilateKernel": ones(9, 9)}, {"dilateKernel": ones(15, 15)}),
            ("config4_rho", {"houghMethod": 5}, {"dilateKernel": ones(15, 15), "houghMethod": 2}),
            ("thresholds", {"dilateKernel": ones(3, 3), "nlinesInSet": 5, "lwTresh": 3},
             {"erodeKernel": ones(3, 3), "dilateKernel": ones(9, 9), "minFlux": 0.03, "addFlux": 1.5})]


@pytest.mark.parametrize("pi", range(3))
def test_nondefault_params_against_reference(cv2mod, pi):
    """Non-default parameter sets (BASELINE.json config 4 among them): the oracle reproduces the unmodified
    reference's returns and every debug tap (SHA-1) on the small frames."""
    g = np.load(os.path.join(GOLD, "golden_params.npz"))
    assert str(g["cv2_version"]) == cv2mod.__version__, "goldens were generated with another cv2"
    name, ob, od = _param_sets()[pi]
    for tag in "abc":
        work = np.ascontiguousarray(small_frame(np.load(os.path.join(GOLD, "golden_small_%s.npz" % tag)))[::-1])
        tb, td = {}, {}
        rb = rp.bright_pass(work, taps=tb, **dict(rp.DEFAULT_BRIGHT, **ob))
        pre = "%s_%s_" % (name, tag)
        assert np.array_equal(sha(work), g[pre + "clipped_bright_sha1"])
        rd = rp.dim_pass(work, taps=td, **dict(rp.DEFAULT_DIM, **od))
        assert np.array_equal(sha(work), g[pre + "clipped_dim_sha1"])
        assert enc(rb) == g[pre + "ret_bright"].tolist() and enc(rd) == g[pre + "ret_dim"].tolist(), (name, tag)
        for key, arr in (("1equBRIGHT", tb["equ"]), ("2dilateBRIGHT", tb["morph"]), ("3contoursBRIGHT", tb["box_img"]),
                         ("6equDIM", td["equ"]), ("7erodedDIM", td["eroded"]), ("8openedDIM", td["morph"]),
                         ("9contoursDIM", td["box_img"])):
            assert np.array_equal(sha(arr), g[pre + key + "_sha1"]), (name, tag, key)
