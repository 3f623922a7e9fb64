"""ORACLE tooling: generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference
(build container only) on seeded synthetic inputs.  Inputs are regenerated from seeds by the tests
(lfd_b200.synth), so the fixtures only hold the reference's outputs:

* golden_small_<kind>.npz : 300x420 frame; the reference's own debug taps (processfield.py:349-378,
  :459-496 write 1equBRIGHT/2dilateBRIGHT/3contoursBRIGHT/6equDIM/7erodedDIM/8openedDIM/9contoursDIM PNGs,
  read back losslessly), the (bool, dict) returns of process_field_bright/dim and the clipped float image.
* golden_full.npz : 2048x1489 frames; returns of both passes + SHA-1 of every debug tap.
* golden_params.npz : the small frames with non-default parameter sets (BASELINE.json config 4: larger dilation
  kernel, finer rho; other thresholds / clip values): returns + SHA-1 of every debug tap.
* golden_run.npz  : a synthetic SDSS tree run through the reference's DetectTrails(...).process() with the
  fitsio stand-in: the bytes of results.txt / errors.txt and the frame visiting order for every _pick mode.

Usage (in the build container):  python oracle/gen_golden.py
"""
import hashlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from lfd_b200 import synth  # noqa: E402
from oracle import load_reference as lr  # noqa: E402

cv2.setNumThreads(1)
GOLD = os.path.join(ROOT, "tests", "golden")

SMALL = {"h": 300, "w": 420}
SMALL_CASES = [("a", 101, 60, 3.0), ("b", 202, 400, 8.0), ("c", 303, 30, 0.0)]
FULL_CASES = [("trail", 1234), ("dense_trail", 7), ("satellite", 9), ("sparse", 11)]


def small_frame(seed, nstars, peak):
    trails = [] if peak == 0 else [{"p0": (10, 20), "p1": (400, 270), "sigma": 2.5, "peak": peak}]
    img, _ = synth.make_frame(seed, n_stars=nstars, h=SMALL["h"], w=SMALL["w"], trails=trails)
    return img


# non-default parameter sets (BASELINE.json config 4: larger dilation, finer rho; other thresholds / clip values)
def param_sets():
    ones = lambda a, b: np.ones((a, b), np.uint8)   # noqa: E731
    return [("config4_dilate", {"dilateKernel": ones(9, 9)}, {"dilateKernel": ones(15, 15)}),
            ("config4_rho", {"houghMethod": 5}, {"dilateKernel": ones(15, 15), "houghMethod": 2}),
            ("thresholds", {"dilateKernel": ones(3, 3), "nlinesInSet": 5, "lwTresh": 3},
             {"erodeKernel": ones(3, 3), "dilateKernel": ones(9, 9), "minFlux": 0.03, "addFlux": 1.5})]


def run_with_taps(pf, img, dbg, over_bright=None, over_dim=None):
    """Both passes of the reference with debug=True; returns dict of outputs."""
    os.environ["DEBUG_PATH"] = dbg
    pf.setup_debug()
    from oracle import ref_pipeline as rp
    pb = dict(rp.DEFAULT_BRIGHT, debug=True, **(over_bright or {}))
    pd = dict(rp.DEFAULT_DIM, debug=True, **(over_dim or {}))
    out = {}
    work = np.ascontiguousarray(img[::-1]).copy()
    stdout = sys.stdout
    sys.stdout = io.StringIO()
    try:
        rb = pf.process_field_bright(work, **pb)
        out["clipped_bright"] = work.copy()
        rd = pf.process_field_dim(work, **pd)
        out["clipped_dim"] = work.copy()
    finally:
        sys.stdout = stdout
    for name in ("1equBRIGHT", "2dilateBRIGHT", "3contoursBRIGHT", "6equDIM", "7erodedDIM", "8openedDIM", "9contoursDIM"):
        p = os.path.join(dbg, name + ".png")
        out[name] = cv2.imread(p, cv2.IMREAD_GRAYSCALE)
        os.remove(p)
    for n in os.listdir(dbg):
        if n.endswith(".png"):
            os.remove(os.path.join(dbg, n))
    def enc(r):
        return np.array([int(r[0])] + ([r[1]["x1"], r[1]["y1"], r[1]["x2"], r[1]["y2"]] if r[0] else [0, 0, 0, 0]), np.int64)
    out["ret_bright"] = enc(rb)
    out["ret_dim"] = enc(rd)
    return out


def main():
    assert lr.available(), "/root/reference is not mounted"
    os.makedirs(GOLD, exist_ok=True)
    pf = lr.load_processfield()
    with tempfile.TemporaryDirectory() as dbg:
        for tag, seed, nstars, peak in SMALL_CASES:
            out = run_with_taps(pf, small_frame(seed, nstars, peak), dbg)
            np.savez_compressed(os.path.join(GOLD, "golden_small_%s.npz" % tag), seed=seed, nstars=nstars, peak=peak,
                                cv2_version=cv2.__version__, numpy_version=np.__version__, **out)
        full = {}
        for kind, seed in FULL_CASES:
            img, _cat = synth.make_case(kind, seed)
            out = run_with_taps(pf, img, dbg)
            for k, v in out.items():
                if k.startswith("ret_"):
                    full["%s_%d_%s" % (kind, seed, k)] = v
                else:
                    full["%s_%d_%s_sha1" % (kind, seed, k)] = np.frombuffer(hashlib.sha1(np.ascontiguousarray(v).tobytes()).digest(), np.uint8)
        np.savez_compressed(os.path.join(GOLD, "golden_full.npz"), cv2_version=cv2.__version__, **full)
        # non-default parameters on the small frames: returns + SHA-1 of every tap
        par = {}
        for name, ob, od in param_sets():
            for tag, seed, nstars, peak in SMALL_CASES:
                out = run_with_taps(pf, small_frame(seed, nstars, peak), dbg, ob, od)
                for k, v in out.items():
                    key = "%s_%s_%s" % (name, tag, k)
                    if k.startswith("ret_"):
                        par[key] = v
                    else:
                        par[key + "_sha1"] = np.frombuffer(hashlib.sha1(np.ascontiguousarray(v).tobytes()).digest(), np.uint8)
        np.savez_compressed(os.path.join(GOLD, "golden_params.npz"), cv2_version=cv2.__version__, **par)

    # whole-driver run through the reference's DetectTrails
    dtmod = lr.load_detecttrails()
    with tempfile.TemporaryDirectory() as tmp:
        kinds = {("r", 100): "trail", ("r", 101): "sparse", ("r", 102): "satellite", ("g", 100): "dense_trail",
                 ("g", 101): "sparse", ("g", 102): "empty"}
        tree = synth.write_sdss_tree(tmp, 2888, 1, [100, 101, 102], filters=("r", "g"), kinds=kinds,
                                     startfield=100, endfield=103)
        dtmod.setup(tree["bosspath"], tree["photoobjpath"], tree["photoreduxpath"], tmp)
        res, err = os.path.join(tmp, "results.txt"), os.path.join(tmp, "errors.txt")
        for flt in ("r", "g"):
            dtmod.DetectTrails(run=2888, camcol=1, filter=flt, results=res, errors=err).process()
        dtmod.DetectTrails(run=2888, camcol=1, filter="r", field=555, results=res, errors=err).process()  # missing file
        results_txt = open(res).read()
        errors_txt = open(err).read().replace(tmp, "$TMP")
        # frame visiting order for every selection mode (capture the calls, do no pixel work)
        orders = {}
        real = dtmod.detecttrails.process_field
        calls = []
        dtmod.detecttrails.process_field = lambda results, errors, run, camcol, filter, field, *a: calls.append((int(run), int(camcol), str(filter), int(field)))
        try:
            for name, kw in (("run", dict(run=2888)), ("run-camcol", dict(run=2888, camcol=2)),
                             ("run-filter", dict(run=2888, filter="i")), ("run-camcol-filter", dict(run=2888, camcol=3, filter="z")),
                             ("camcol-filter", dict(camcol=4, filter="u")), ("camcol-frame", dict(run=2888, camcol=5, field=101)),
                             ("field", dict(run=2888, camcol=6, filter="g", field=102))):
                calls.clear()
                d = dtmod.DetectTrails(results=res + ".x", errors=err + ".x", **kw)
                assert d._pick == name, (d._pick, name)
                d.process()
                orders[name] = np.array([(r, c, "ugriz".index(f), fl) for r, c, f, fl in calls], np.int64).reshape(-1, 4)
        finally:
            dtmod.detecttrails.process_field = real
        np.savez_compressed(os.path.join(GOLD, "golden_run.npz"), results_txt=np.array(results_txt), errors_txt=np.array(errors_txt),
                            **{"order_" + k: v for k, v in orders.items()})
    print("golden fixtures written to", GOLD)
    for n in sorted(os.listdir(GOLD)):
        print("  %-28s %8d bytes" % (n, os.path.getsize(os.path.join(GOLD, n))))


if __name__ == "__main__":
    main()
