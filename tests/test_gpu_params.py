"""GPU parity for non-default parameters and frame shapes (BASELINE.json config 4: larger dilation kernel,
finer rho; plus kernel shapes that take the generic shared-memory morphology path and frames whose width is
not a multiple of 8).  Same oracle as test_gpu_stages: cv2 at the reference's call sites."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lfd_b200 import synth
from oracle import ref_pipeline as rp


def _ones(h, w):
    return np.ones((h, w), np.uint8)


def _check(h, pass_, work, params, shape_tag):
    from lfd_b200 import _lib
    from lfd_b200.processfield import result_from_device
    ref_img = work.copy()
    taps = {}
    ref_res = (rp.bright_pass if pass_ == 0 else rp.dim_pass)(ref_img, taps=taps, **params)
    r = h.run_pass(pass_, work, flags=_lib.KEEP_TAPS | _lib.FULL_LINES, writeback=True)
    errs = []
    if not np.array_equal(work.view(np.uint32), ref_img.view(np.uint32)):
        errs.append("clipped")
    for stage, key in (("gray", "gray"), ("equ", "equ"), ("morph", "morph"), ("canny", "canny"), ("box", "box_img")):
        got = h.stage(0, pass_, stage)
        if not np.array_equal(got, taps[key]):
            errs.append("%s: %d px differ" % (stage, int((got != taps[key]).sum())))
    if pass_ == 1 and "eroded" in taps and params.get("erodeKernel") is not None:
        if not np.array_equal(h.stage(0, pass_, "eroded"), taps["eroded"]):
            errs.append("eroded")
    # the retrieved contour set: as many rectangles as cv2.findContours returned contours in this mode, same passing ones
    rects = h.stage(0, pass_, "rects")
    got_rects = [r for r in rects if r["kind"] != 2]
    if len(got_rects) != len(taps["rects"]):
        errs.append("contours: %d retrieved vs %d from cv2" % (len(got_rects), len(taps["rects"])))
    gp = sorted(((float(a["cx"]), float(a["cy"])), (float(a["w"]), float(a["h"])), float(a["angle"])) for a in got_rects if a["passed"])
    if gp != sorted(x for x, _ in taps["passing"]):
        errs.append("passing rects differ: %d vs %d" % (len(gp), len(taps["passing"])))
    if taps["passing"]:
        for which, key in (("equ", "lines_equ"), ("box", "lines_box")):
            ref_lines = taps[key]
            got_lines = h.stage(0, pass_, "lines_" + which)
            if ref_lines is None:
                if len(got_lines):
                    errs.append("lines_%s: expected none" % which)
            elif got_lines.shape != ref_lines.shape or not np.array_equal(got_lines.view(np.uint32), ref_lines.view(np.uint32)):
                errs.append("lines_%s: %s vs %s" % (which, got_lines.shape, ref_lines.shape))
    assert not errs, "%s pass %d: %s" % (shape_tag, pass_, "; ".join(errs))
    assert result_from_device(r, pass_, work.shape) == ref_res


def _cross(n):
    k = np.zeros((n, n), np.uint8); k[n // 2, :] = 1; k[:, n // 2] = 1
    return k


def _ellipse(h, w):
    import cv2
    return cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (w, h))


PARAM_SETS = [
    # (bright overrides, dim overrides)
    ({"dilateKernel": _cross(5)}, {"erodeKernel": _cross(3), "dilateKernel": _ellipse(9, 9)}),           # non-rectangular elements
    ({"dilateKernel": _ones(27, 25)}, {"erodeKernel": _ones(5, 5), "dilateKernel": _ones(31, 31)}),          # rectangles beyond the tile halo
    ({"dilateKernel": _ellipse(7, 4)}, {"erodeKernel": _ones(3, 3), "dilateKernel": (np.random.default_rng(5).random((11, 8)) < 0.4).astype(np.uint8) | _cross(11)[:, :8]}),
    ({"dilateKernel": _ones(9, 9)}, {"dilateKernel": _ones(15, 15)}),                                    # config 4: larger dilation
    ({"houghMethod": 5}, {"dilateKernel": _ones(15, 15), "houghMethod": 2}),                            # finer rho
    ({"dilateKernel": _ones(5, 3)}, {"erodeKernel": _ones(2, 2), "dilateKernel": _ones(6, 7)}),          # generic tile kernel
    ({"contoursMode": 0}, {"contoursMode": 0}),                                                          # cv2.RETR_EXTERNAL
    ({"contoursMode": 0, "dilateKernel": _ones(9, 9), "lwTresh": 2}, {"contoursMode": 0, "contoursMethod": 2, "lwTresh": 3}),
    ({"dilateKernel": _ones(3, 3), "nlinesInSet": 5, "lwTresh": 3}, {"erodeKernel": _ones(3, 3), "dilateKernel": _ones(9, 9), "minFlux": 0.03, "addFlux": 1.5}),
]


@pytest.mark.parametrize("pi", range(len(PARAM_SETS)))
def test_nondefault_params(cv2mod, pi):
    from lfd_b200 import _lib
    ob, od = PARAM_SETS[pi]
    pb, pd = dict(rp.DEFAULT_BRIGHT, **ob), dict(rp.DEFAULT_DIM, **od)
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=1)
    try:
        h.set_params(pb, pd)
        for kind, seed in (("trail", 31), ("dense_trail", 32)):
            img, _cat = synth.make_case(kind, seed)
            work = np.ascontiguousarray(img[::-1])
            _check(h, 0, work.copy(), pb, "%s/%d" % (kind, pi))
            w1 = work.copy()
            w1[w1 < 0] = 0
            _check(h, 1, w1, pd, "%s/%d" % (kind, pi))
    finally:
        h.close()


@pytest.mark.parametrize("shape", [(200, 256), (123, 100), (64, 36), (301, 520)])
def test_other_frame_shapes(cv2mod, shape):
    """Crops of a synthetic frame: widths that are / are not multiples of 8 and 32, heights that are not multiples
    of the 64-row marching chunk or the 32-row CCL band."""
    from lfd_b200 import _lib
    H, W = shape
    img, _cat = synth.make_case("trail", 77)
    crop = np.ascontiguousarray(img[::-1][100:100 + H, 150:150 + W])
    # a short bright streak so that rectangles pass in small crops too
    yy = np.arange(H)
    for t in range(-2, 3):
        xs = np.clip((yy * (W - 20) // max(H, 1)) + 10 + t, 0, W - 1)
        crop[yy, xs] += 40.0
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h = _lib.Handle(H, W, max_batch=1)
    try:
        h.set_params(pb, pd)
        _check(h, 0, crop.copy(), pb, str(shape))
        w1 = crop.copy()
        w1[w1 < 0] = 0
        _check(h, 1, w1, pd, str(shape))
    finally:
        h.close()


def test_special_float_values(cv2mod):
    """NaN, +-inf, |v| >= 2^31, exact .5 ties and the 255 saturation edge go through convertScaleAbs the way cv2
    does it: the fast conversion path must hand these to its exact slow path."""
    from lfd_b200 import _lib
    H, W = 96, 128
    rng = np.random.default_rng(3)
    img = rng.normal(0, 0.03, (H, W)).astype(np.float32)
    specials = np.array([np.nan, np.inf, -np.inf, 2.0 ** 31, 2.0 ** 31 + 4096, 3.0e9, -3.0e9, 2147483520.0, 254.5, 255.0, 255.49, 255.5, 256.0,
                         1000.0, 0.5, 1.5, 2.5, 3.5, 127.5, 128.5, -0.5, -1.5, 0.02, 0.019999, 1e-30, 65535.0, 16777216.0, 8388608.0, 8388607.5],
                        np.float32)
    # huge values whose bit pattern looks like a small integer in the low mantissa byte (0x5700003c = 1.4e14 ...):
    # the magic-add conversion must not mistake them for in-range results
    specials = np.concatenate([specials, np.array([0x5700003c, 0x5700007d, 0x60000001, 0x7f000055, 0x4c0000ff, 0x4b8000aa],
                                                  np.uint32).view(np.float32)])
    ys = rng.integers(0, H, len(specials) * 3)
    xs = rng.integers(0, W, len(specials) * 3)
    img[ys, xs] = np.tile(specials, 3)
    img[40:44, 10:110] += 30.0        # something for the later stages to find
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h = _lib.Handle(H, W, max_batch=2)
    try:
        h.set_params(pb, pd)
        # standalone passes (run_pass, generic prep kernel with write-back)
        for pass_, params in ((0, pb), (1, pd)):
            work = img.copy()
            if pass_ == 1:
                with np.errstate(invalid="ignore"):
                    work[work < 0] = 0
            ref = work.copy()
            taps = {}
            with np.errstate(invalid="ignore"):
                (rp.bright_pass if pass_ == 0 else rp.dim_pass)(ref, taps=taps, **params)
            h.run_pass(pass_, work, flags=_lib.KEEP_TAPS, writeback=True)
            assert np.array_equal(h.stage(0, pass_, "gray"), taps["gray"]), pass_
            assert np.array_equal(work.view(np.uint32), ref.view(np.uint32)), pass_
        # whole-frame path (fast prep kernel): the un-flipped frame goes in, gray comes out flipped
        h.submit(np.stack([img, img]), [np.zeros((0, 4), np.int32)] * 2, flags=_lib.KEEP_TAPS | _lib.SERIAL_PASSES)
        h.wait()
        work = np.ascontiguousarray(img[::-1]).copy()
        tb, td = {}, {}
        with np.errstate(invalid="ignore"):
            rp.bright_pass(work, taps=tb, **pb)
            rp.dim_pass(work, taps=td, **pd)
        for fr in (0, 1):
            assert np.array_equal(h.stage(fr, 0, "gray"), tb["gray"])
            assert np.array_equal(h.stage(fr, 1, "gray"), td["gray"])
            assert np.array_equal(h.stage(fr, 0, "hist").astype(np.int64), np.bincount(tb["gray"].ravel(), minlength=256))
            assert np.array_equal(h.stage(fr, 1, "hist").astype(np.int64), np.bincount(td["gray"].ravel(), minlength=256))
    finally:
        h.close()


def test_config5_microbench_hough_canny(cv2mod):
    """BASELINE.json config 5: 4096x4096 frames, non-zero density sweep, rho / theta sweep, against cv2 directly
    (bit-exact line lists and edge maps)."""
    from lfd_b200 import _lib
    S = 4096
    rng = np.random.default_rng(55)
    h = _lib.Handle(S, S, max_batch=1, max_runs=1 << 22)
    try:
        pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
        h.set_params(pb, pd)
        for density, rho, theta in ((0.001, 20, np.pi / 180), (0.01, 5, np.pi / 360), (0.03, 20, np.pi / 720), (0.10, 20, np.pi / 180),
                                    (0.003, 1, np.pi / 180)):
            img = (rng.random((S, S)) < density).astype(np.uint8) * 255
            cv2mod.line(img, (100, 50), (3900, 3000), 255, 2)
            cv2mod.line(img, (4000, 10), (30, 4090), 255, 1)
            ref = cv2mod.HoughLines(img, rho, theta, 1)
            got, _ = h.hough_lines(img, rho, theta, 1)
            assert (ref is None) == (got is None)
            assert got.shape == ref.shape, (density, rho, theta, got.shape, ref.shape)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (density, rho, theta)
        # Canny on grey images of increasing busyness
        for sigma, lo, hi in ((2.0, 0, 255), (6.0, 0, 255), (12.0, 50, 150)):
            base = np.zeros((S, S), np.float32)
            for _ in range(400):
                y, x = rng.integers(50, S - 50, 2)
                base[y - 6:y + 6, x - 6:x + 6] += rng.uniform(20, 200)
            noisy = np.clip(base + rng.normal(0, sigma, (S, S)), 0, 255).astype(np.uint8)
            ref = cv2mod.Canny(noisy, lo, hi)
            got = h.canny(noisy, lo, hi)
            assert np.array_equal(got, ref), (sigma, lo, hi, int((got != ref).sum()))
    finally:
        h.close()


def test_fit_min_area_rect_standalone(cv2mod):
    """lfd_b200.processfield.fit_minAreaRect == the reference's fit_minAreaRect (processfield.py:201-263) on uint8 images."""
    from lfd_b200.processfield import fit_minAreaRect
    img, _ = synth.make_case("dense_trail", 91)
    taps = {}
    rp.bright_pass(np.ascontiguousarray(img[::-1]).copy(), taps=taps, **rp.DEFAULT_BRIGHT)
    equ = taps["morph"]
    for mode, method, minlen, lw in ((1, 1, 1, 5), (0, 2, 1, 3), (2, 1, 3, 2)):
        ref_det, ref_box = rp.fit_rects(equ, mode, method, minlen, lw)
        det, box = fit_minAreaRect(equ, mode, method, minlen, lw, False)
        assert det == ref_det and np.array_equal(box, ref_box), (mode, method, minlen, lw)


def test_capacity_overflow_is_reported_not_silent(cv2mod):
    """A frame with more runs / contours than the configured capacity comes back flagged (LFD_FRAME_OVERFLOW) instead of
    with a wrong verdict; its batch neighbours are unaffected."""
    from lfd_b200 import _lib
    from lfd_b200.processfield import result_from_device
    dense, _ = synth.make_case("dense", 6)
    sparse, cat = synth.make_case("trail", 1234)
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=2, max_runs=20000, max_components=2000)
    try:
        h.set_params(pb, pd)
        h.submit(np.stack([dense, sparse]), [np.zeros((0, 4), np.int32)] * 2)
        r = h.wait()
        assert r[0].status & _lib.FRAME_OVERFLOW
        with pytest.raises(_lib.LfdError):
            result_from_device(r[0], 0, dense.shape)
        assert not (r[1].status & _lib.FRAME_OVERFLOW)
        ref = rp.process_frame(sparse.copy(), {k: v[:0] for k, v in cat.items()}, "r")
        got = (False, -1, None)
        for p in (0, 1):
            if r[1].rect_detection[p] >= 0:
                d_, o_ = result_from_device(r[1], p, sparse.shape)
                if d_:
                    got = (True, p, o_)
                    break
        assert got == ref
    finally:
        h.close()
