"""GPU parity for non-default parameters and frame shapes (BASELINE.json config 4: larger dilation kernel,
finer rho; plus kernel shapes that take the generic shared-memory morphology path and frames whose width is
not a multiple of 8).  Same oracle as test_gpu_stages: cv2 at the reference's call sites."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lfd_b200 import synth
from oracle import cv_restate as cr
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


PARAM_SETS = [
    # (bright overrides, dim overrides)
    ({"dilateKernel": _ones(9, 9)}, {"dilateKernel": _ones(15, 15)}),                                    # config 4: larger dilation
    ({"houghMethod": 5}, {"dilateKernel": _ones(15, 15), "houghMethod": 2}),                            # finer rho
    ({"dilateKernel": _ones(5, 3)}, {"erodeKernel": _ones(2, 2), "dilateKernel": _ones(6, 7)}),          # generic tile kernel
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
