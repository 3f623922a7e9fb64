"""GPU parity against the golden vectors produced by the UNMODIFIED reference (tests/golden, see
oracle/gen_golden.py): the reference's own debug taps, its (bool, dict) returns and its results.txt."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lfd_b200 import synth
from oracle import ref_pipeline as rp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def enc(r):
    return [int(r[0])] + ([r[1]["x1"], r[1]["y1"], r[1]["x2"], r[1]["y2"]] if r[0] else [0, 0, 0, 0])


def sha(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def _run_both(h, work):
    from lfd_b200 import _lib
    from lfd_b200.processfield import result_from_device
    out = {}
    r = h.run_pass(0, work, flags=_lib.KEEP_TAPS, writeback=True)
    out["ret_bright"] = enc(result_from_device(r, 0, work.shape))
    out["clipped_bright"] = work.copy()
    out.update({"1equBRIGHT": h.stage(0, 0, "equ"), "2dilateBRIGHT": h.stage(0, 0, "morph"), "3contoursBRIGHT": h.stage(0, 0, "box")})
    r = h.run_pass(1, work, flags=_lib.KEEP_TAPS, writeback=True)
    out["ret_dim"] = enc(result_from_device(r, 1, work.shape))
    out["clipped_dim"] = work.copy()
    out.update({"6equDIM": h.stage(0, 1, "equ"), "7erodedDIM": h.stage(0, 1, "eroded"), "8openedDIM": h.stage(0, 1, "morph"),
                "9contoursDIM": h.stage(0, 1, "box")})
    return out


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_small_goldens(tag):
    from lfd_b200 import _lib
    g = np.load(os.path.join(GOLD, "golden_small_%s.npz" % tag))
    peak = float(g["peak"])
    trails = [] if peak == 0 else [{"p0": (10, 20), "p1": (400, 270), "sigma": 2.5, "peak": peak}]
    img, _ = synth.make_frame(int(g["seed"]), n_stars=int(g["nstars"]), h=300, w=420, trails=trails)
    h = _lib.Handle(300, 420, max_batch=1)
    try:
        h.set_params(dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM))
        out = _run_both(h, np.ascontiguousarray(img[::-1]))
    finally:
        h.close()
    for k, v in out.items():
        ref = g[k]
        if k.startswith("ret_"):
            assert v == ref.tolist(), k
        else:
            assert np.array_equal(v, ref), "%s: %d px differ" % (k, int(np.sum(v != ref)))


def test_full_goldens():
    from lfd_b200 import _lib
    g = np.load(os.path.join(GOLD, "golden_full.npz"))
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=1)
    try:
        h.set_params(dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM))
        for kind, seed in (("trail", 1234), ("dense_trail", 7), ("satellite", 9), ("sparse", 11)):
            img, _ = synth.make_case(kind, seed)
            out = _run_both(h, np.ascontiguousarray(img[::-1]))
            pre = "%s_%d_" % (kind, seed)
            for k, v in out.items():
                if k.startswith("ret_"):
                    assert v == g[pre + k].tolist(), pre + k
                else:
                    assert np.array_equal(sha(v), g[pre + k + "_sha1"]), pre + k
    finally:
        h.close()


def test_results_txt_golden(tmp_path):
    import lfd_b200
    g = np.load(os.path.join(GOLD, "golden_run.npz"))
    kinds = {("r", 100): "trail", ("r", 101): "sparse", ("r", 102): "satellite", ("g", 100): "dense_trail",
             ("g", 101): "sparse", ("g", 102): "empty"}
    tree = synth.write_sdss_tree(str(tmp_path), 2888, 1, [100, 101, 102], filters=("r", "g"), kinds=kinds,
                                 startfield=100, endfield=103)
    lfd_b200.setup(tree["bosspath"], tree["photoobjpath"], tree["photoreduxpath"], str(tmp_path))
    res, err = str(tmp_path / "results.txt"), str(tmp_path / "errors.txt")
    for flt in ("r", "g"):
        lfd_b200.DetectTrails(run=2888, camcol=1, filter=flt, results=res, errors=err, batch=4).process()
    lfd_b200.DetectTrails(run=2888, camcol=1, filter="r", field=555, results=res, errors=err).process()
    assert open(res).read() == str(g["results_txt"])
    got_err = open(err).read().replace(str(tmp_path), "$TMP")
    ref_err = str(g["errors_txt"])
    # same frame id line, same exception type and message; the traceback's file/line text differs by construction
    assert got_err.splitlines()[0] == ref_err.splitlines()[0] == "2888 1 r 555"
    assert got_err.splitlines()[-2] == ref_err.splitlines()[-2]
    assert "FileNotFoundError" in got_err


def _param_sets():
    ones = lambda a, b: np.ones((a, b), np.uint8)   # noqa: E731  (same sets as oracle/gen_golden.py::param_sets)
    return [("config4_dilate", {"dilateKernel": ones(9, 9)}, {"dilateKernel": ones(15, 15)}),
            ("config4_rho", {"houghMethod": 5}, {"dilateKernel": ones(15, 15), "houghMethod": 2}),
            ("thresholds", {"dilateKernel": ones(3, 3), "nlinesInSet": 5, "lwTresh": 3},
             {"erodeKernel": ones(3, 3), "dilateKernel": ones(9, 9), "minFlux": 0.03, "addFlux": 1.5})]


@pytest.mark.parametrize("pi", range(3))
def test_nondefault_params_goldens(pi):
    """Non-default parameter sets (BASELINE.json config 4: larger dilation kernel, finer rho; other thresholds and
    clip values) against the unmodified reference: returns and SHA-1 of every debug tap on the small frames."""
    from lfd_b200 import _lib
    g = np.load(os.path.join(GOLD, "golden_params.npz"))
    name, ob, od = _param_sets()[pi]
    h = _lib.Handle(300, 420, max_batch=1)
    try:
        h.set_params(dict(rp.DEFAULT_BRIGHT, **ob), dict(rp.DEFAULT_DIM, **od))
        for tag in "abc":
            gs = np.load(os.path.join(GOLD, "golden_small_%s.npz" % tag))
            peak = float(gs["peak"])
            trails = [] if peak == 0 else [{"p0": (10, 20), "p1": (400, 270), "sigma": 2.5, "peak": peak}]
            img, _ = synth.make_frame(int(gs["seed"]), n_stars=int(gs["nstars"]), h=300, w=420, trails=trails)
            out = _run_both(h, np.ascontiguousarray(img[::-1]))
            for k, v in out.items():
                key = "%s_%s_%s" % (name, tag, k)
                if k.startswith("ret_"):
                    assert v == g[key].tolist(), key
                else:
                    assert np.array_equal(sha(v), g[key + "_sha1"]), key
    finally:
        h.close()
