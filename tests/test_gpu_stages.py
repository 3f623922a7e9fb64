"""GPU parity, stage by stage, through the C ABI (lfd_run_pass + lfd_get_stage) against the oracle:
cv2 at the reference's call sites (oracle/ref_pipeline.py) and the restated internals
(oracle/cv_restate.py).  Bit-exact for every integer/byte stage; rect floats exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lfd_b200 import synth
from oracle import cv_restate as cr
from oracle import ref_pipeline as rp


@pytest.fixture(scope="module")
def handle():
    from lfd_b200 import _lib
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=2)
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h.set_params(pb, pd)
    yield h
    h.close()


def _diff(name, got, ref):
    if got.shape != ref.shape:
        return "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    bad = np.argwhere(got != ref)
    if len(bad) == 0:
        return None
    first = [tuple(int(v) for v in b) for b in bad[:6]]
    vals = [(got[t].item(), ref[t].item()) for t in first]
    return "%s: %d mismatches, first at %s (got, ref)=%s" % (name, len(bad), first, vals)


CASES = [("trail", 1234), ("sparse", 11), ("dense_trail", 7), ("satellite", 9), ("empty", 5)]


@pytest.mark.parametrize("kind,seed", CASES)
@pytest.mark.parametrize("pass_", [0, 1])
def test_pass_stages(handle, cv2mod, kind, seed, pass_):
    from lfd_b200 import _lib
    img, _cat = synth.make_case(kind, seed)
    work = np.ascontiguousarray(img[::-1])        # cv2.flip(img, 0)
    ref_img = work.copy()
    taps = {}
    if pass_ == 0:
        ref_res = rp.bright_pass(ref_img, taps=taps, **rp.DEFAULT_BRIGHT)
    else:
        ref_img[ref_img < 0] = 0                   # what bright leaves behind (detecttrails.py:125-129)
        work = ref_img.copy()
        ref_res = rp.dim_pass(ref_img, taps=taps, **rp.DEFAULT_DIM)
    r = handle.run_pass(pass_, work, flags=_lib.KEEP_TAPS | _lib.FULL_LINES, writeback=True)
    errs = []
    # in-place clip
    e = _diff("clipped", work.view(np.uint32), ref_img.view(np.uint32))
    if e: errs.append(e)
    for stage, key in (("gray", "gray"), ("equ", "equ"), ("morph", "morph"), ("canny", "canny"), ("box", "box_img")):
        e = _diff(stage, handle.stage(0, pass_, stage), taps[key])
        if e: errs.append(e)
    if pass_ == 1:
        e = _diff("eroded", handle.stage(0, pass_, "eroded"), taps["eroded"])
        if e: errs.append(e)
    hist = handle.stage(0, pass_, "hist")
    e = _diff("hist", hist.astype(np.int64), np.bincount(taps["gray"].ravel(), minlength=256))
    if e: errs.append(e)
    cls, _ = cr.canny_classes(taps["morph"], 0, 255)
    e = _diff("nms", handle.stage(0, pass_, "nms"), cls)
    if e: errs.append(e)
    # labels
    e = _diff("fg_labels", handle.stage(0, pass_, "fg_labels"), cr.label_fg8(taps["canny"]))
    if e: errs.append(e)
    e = _diff("bg_labels", handle.stage(0, pass_, "bg_labels"), cr.label_bg4(taps["canny"]))
    if e: errs.append(e)
    # rectangles: same multiset as cv2.minAreaRect over cv2.findContours (tie cases excepted), passing set exact
    rects = handle.stage(0, pass_, "rects")
    if len(rects) != len(taps["rects"]):
        errs.append("rect count %d vs %d" % (len(rects), len(taps["rects"])))
    else:
        import collections
        got = collections.Counter(((float(a["cx"]), float(a["cy"])), (float(a["w"]), float(a["h"])), float(a["angle"])) for a in rects)
        ref = collections.Counter(taps["rects"])
        miss = sum((ref - got).values())
        if miss > max(2, len(taps["rects"]) // 200):
            errs.append("rects: %d of %d differ, e.g. ref %s got %s" % (miss, len(taps["rects"]), list((ref - got))[:3], list((got - ref))[:3]))
        gp = sorted(((float(a["cx"]), float(a["cy"])), (float(a["w"]), float(a["h"])), float(a["angle"])) for a in rects if a["passed"])
        rpass = sorted(x for x, _ in taps["passing"])
        if gp != rpass:
            errs.append("passing rects differ: %d vs %d" % (len(gp), len(rpass)))
    assert r.rect_detection[pass_] == (1 if len(taps["passing"]) else 0), errs
    if taps["passing"]:
        for which, key in (("equ", "lines_equ"), ("box", "lines_box")):
            ref_lines = taps[key]
            got_lines = handle.stage(0, pass_, "lines_" + which)
            if ref_lines is None:
                if len(got_lines): errs.append("lines_%s: expected none, got %d" % (which, len(got_lines)))
            else:
                e = _diff("lines_" + which, got_lines.view(np.uint32), ref_lines.view(np.uint32))
                if e: errs.append(e)
            src = taps["morph"] if which == "equ" else taps["box_img"]
            _, acc, _ = cr.hough_lines(src, rp.DEFAULT_BRIGHT["houghMethod"], np.pi / 180, 1)
            e = _diff("accum_" + which, handle.stage(0, pass_, "accum_" + which).reshape(acc.shape), acc)
            if e: errs.append(e)
    assert not errs, "\n".join(errs)
    # verdict
    from lfd_b200.processfield import result_from_device
    assert result_from_device(r, pass_, work.shape) == ref_res


def test_hough_standalone(handle, cv2mod):
    rng = np.random.default_rng(4)
    for (H, W) in ((120, 160), (333, 517)):
        img = (rng.random((H, W)) < 0.02).astype(np.uint8) * 255
        cv2mod.line(img, (5, 7), (W - 10, H - 20), 255, 1)
        for rho in (20, 1, 0.5, 3):
            for theta in (np.pi / 180, np.pi / 360, np.pi / 90):
                ref = cv2mod.HoughLines(img, rho, theta, 1)
                got, accum = handle.hough_lines(img, rho, theta, 1, want_accum=True)
                _, racc, _ = cr.hough_lines(img, rho, theta, 1)
                assert np.array_equal(accum, racc), (H, W, rho, theta)
                assert (ref is None) == (got is None)
                if ref is not None:
                    assert got.shape == ref.shape, (H, W, rho, theta, got.shape, ref.shape)
                    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (H, W, rho, theta)


def test_hough_fine_rho_on_blobs(handle, cv2mod):
    """Fine rho on blob-like full-size masks (what config 4 feeds HoughLines: rho = 1 px on the 15x15-dilated plane;
    2-4 angles per CTA, words walked bit by bit, single privatised copy flushed with plain stores): the accumulator
    equals the restated one cell for cell and the line list equals cv2.HoughLines bit for bit."""
    rng = np.random.default_rng(11)
    H, W = synth.FRAME_H, synth.FRAME_W
    img = np.zeros((H, W), np.uint8)
    for _ in range(140):                                        # filled boxes: mask words are mostly full (>= 6 px per word)
        x, y = int(rng.integers(0, W - 90)), int(rng.integers(0, H - 60))
        img[y:y + int(rng.integers(12, 60)), x:x + int(rng.integers(30, 90))] = int(rng.integers(1, 256))
    cv2mod.line(img, (40, 30), (W - 60, H - 90), 255, 9)
    words = img.reshape(H, W // 32, 32)
    nzw = (words != 0).any(axis=2).sum()
    assert np.count_nonzero(img) >= 6 * nzw, "test image is not blob-like (mask words should be mostly full)"
    for rho in (1, 2, 0.5):
        ref = cv2mod.HoughLines(img, rho, np.pi / 180, 1)
        got, accum = handle.hough_lines(img, rho, np.pi / 180, 1, want_accum=True)
        _, racc, _ = cr.hough_lines(img, rho, np.pi / 180, 1)
        assert np.array_equal(accum, racc), rho
        assert ref is not None and got is not None and got.shape == ref.shape, rho
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), rho
