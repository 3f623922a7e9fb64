"""ORACLE (test infrastructure, not product code): CPU restatement of lfd.detecttrails' per-frame path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
The product path (``lfd_b200``) never does, and fails loudly without its CUDA library.

What this restates, with the reference line each function follows (paths under /root/reference/):

* ``star_rects`` / ``blot``      lfd/detecttrails/removestars.py:111-130 (ceil), :212-231 (filter + blot)
* ``bright_pass`` / ``dim_pass`` lfd/detecttrails/processfield.py:291-388 / :391-506
* ``fit_rects``                  lfd/detecttrails/processfield.py:201-263
* ``check_theta``                lfd/detecttrails/processfield.py:36-150
* ``dictify_hough``              lfd/detecttrails/processfield.py:266-288
* ``process_frame``              lfd/detecttrails/detecttrails.py:113-131 (remove_stars -> flip -> bright -> dim)
* ``result_line``                lfd/detecttrails/detecttrails.py:115-117,127,131

The pixel arithmetic of the reference lives in a third-party dependency that is not under
/root/reference: OpenCV (``opencv-python``, unpinned in setup.py:18-28; 3.4.2 in environment.yml:70)
and NumPy.  The parity target is the binary installed in this image, **cv2 4.13.0 / numpy 2.3**, and
this module calls that binary at the same call sites the reference does, returning every intermediate
("stage taps") so each CUDA kernel can be compared alone.  ``oracle/cv_restate.py`` restates the
OpenCV routines themselves (needed for what cv2 does not expose: Hough accumulators, hull order,
calipers internals) and is pinned bit-for-bit against cv2 in tests/test_oracle_cv.py.

Pinning: the reference has no tests or golden vectors, so this module is pinned
against the *reference code itself* imported from /root/reference in this container
(oracle/gen_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).
"""
import math

import numpy as np

try:  # the oracle needs cv2; the product does not
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

RETR_LIST = 1
CHAIN_APPROX_NONE = 1

DEFAULT_BRIGHT = {
    "lwTresh": 5, "thetaTresh": 0.15, "dilateKernel": np.ones((4, 4), np.uint8),
    "contoursMode": RETR_LIST, "contoursMethod": CHAIN_APPROX_NONE, "minAreaRectMinLen": 1,
    "houghMethod": 20, "nlinesInSet": 3, "lineSetTresh": 0.15, "dro": 25, "debug": False,
}
DEFAULT_DIM = {
    "minFlux": 0.02, "addFlux": 0.5, "lwTresh": 5, "thetaTresh": 0.15,
    "erodeKernel": np.ones((3, 3), np.uint8), "dilateKernel": np.ones((9, 9), np.uint8),
    "contoursMode": RETR_LIST, "contoursMethod": CHAIN_APPROX_NONE, "minAreaRectMinLen": 1,
    "houghMethod": 20, "nlinesInSet": 3, "lineSetTresh": 0.15, "dro": 20, "debug": False,
}
DEFAULT_REMOVESTARS = {
    "pixscale": 0.396, "defaultxy": 20, "maxxy": 60,
    "filter_caps": {"u": 22.0, "g": 22.2, "r": 22.2, "i": 21.3, "z": 20.5},
    "magcount": 3, "maxmagdiff": 3, "debug": False,
}
_BANDS = ("u", "g", "r", "i", "z")


# ----------------------------------------------------------------------------------------------
# remove_stars
# ----------------------------------------------------------------------------------------------
def star_rects(cat, filt, defaultxy, filter_caps, maxxy, pixscale, magcount, maxmagdiff, debug=False):
    """Objects that get blotted, as (x, y, dxy) python ints, in catalog order.

    removestars.py:111-130 turns every band value into ``math.ceil``; :212-230 is the filter:
    psfMag < cap, at most ``magcount`` of the 10 band-pair differences above ``maxmagdiff``,
    nObserve == nDetect, half-side from petro90 (or the default when <= 0 or > maxxy).
    """
    b = _BANDS.index(filt)
    out = []
    n = len(cat["ROWC"])
    for i in range(n):
        x = int(math.ceil(cat["COLC"][i][b]))
        y = int(math.ceil(cat["ROWC"][i][b]))
        mags = [math.ceil(v) for v in cat["PSFMAG"][i]]
        if not mags[b] < filter_caps[filt]:
            continue
        big = 0
        for j in range(5):
            for k in range(j + 1, 5):
                if abs(mags[j] - mags[k]) > maxmagdiff:
                    big += 1
        if magcount < big:
            continue
        dxy = defaultxy
        p90 = math.ceil(cat["PETROTH90"][i][b])
        if p90 > 0:
            dxy = int(p90 / pixscale) + 10
        if dxy > maxxy:
            dxy = defaultxy
        if cat["NOBSERVE"][i] == cat["NDETECT"][i]:
            out.append((x, y, dxy))
    return out


def blot(img, rects):
    """removestars.py:231 - ``img[x-dxy:x+dxy, y-dxy:y+dxy].fill(0.0)`` with axis 0 indexed by x.

    Python slice semantics are the observable behaviour (negative start wraps, so objects closer
    than dxy to the low edges are silently skipped); NumPy does that for us.
    """
    for x, y, dxy in rects:
        img[x - dxy:x + dxy, y - dxy:y + dxy].fill(0.0)
    return img


# ----------------------------------------------------------------------------------------------
# processfield
# ----------------------------------------------------------------------------------------------
def check_theta(h1, h2, navg, dro, thetaTresh, lineSetTresh):
    """processfield.py:36-150.  True = reject, None = accept.  (navg,1) float64 work arrays; the
    four assignments sit in one try so a short ``h2`` leaves theta1[i] at 0 as well (:93-102)."""
    ro1 = np.zeros((navg, 1))
    ro2 = np.zeros((navg, 1))
    th1 = np.zeros((navg, 1))
    th2 = np.zeros((navg, 1))
    for i in range(navg):
        try:
            ro1[i] = h1[i][0][0]
            ro2[i] = h2[i][0][0]
            th1[i] = h1[i][0][1]
            th2[i] = h2[i][0][1]
        except IndexError:
            pass
    if abs(np.average(ro1) - np.average(ro2)) > dro:
        return True
    if abs(th1.max() - th1.min()) > thetaTresh:
        return True
    if abs(th2.max() - th2.min()) > thetaTresh:
        return True
    if np.average(abs(th1 - th2)) > lineSetTresh:
        return True
    return None


def dictify_hough(shape, line):
    """processfield.py:266-288 - float32 NumPy scalar arithmetic, then int() truncation."""
    rho, theta = line
    n_x, n_y = shape
    x0 = np.cos(theta) * rho
    y0 = np.sin(theta) * rho
    return {"x1": int(x0 - (n_x + n_y) * np.sin(theta)), "y1": int(y0 + (n_x + n_y) * np.cos(theta)),
            "x2": int(x0 + (n_x + n_y) * np.sin(theta)), "y2": int(y0 - (n_x + n_y) * np.cos(theta))}


def fit_rects(img8, contoursMode, contoursMethod, minAreaRectMinLen, lwTresh, taps=None):
    """processfield.py:201-263.  Returns (detection, box_img)."""
    box_img = np.zeros(img8.shape, np.uint8)
    canny = cv2.Canny(img8, 0, 255)
    found = cv2.findContours(canny, contoursMode, contoursMethod)
    contours = found[0] if len(found) == 2 else found[1]
    detection = False
    rects, passing = [], []
    for cnt in contours:
        rect = cv2.minAreaRect(cnt)
        rects.append(rect)
        w, h = rect[1]
        length, width = (w, h) if w > h else (h, w)
        if length > minAreaRectMinLen and width > minAreaRectMinLen:
            if length / width > lwTresh:
                detection = True
                box = np.asarray(cv2.boxPoints(rect), dtype=np.int32)
                cv2.fillPoly(box_img, [box], (255, 255, 255))
                passing.append((rect, box))
    if taps is not None:
        taps["canny"] = canny
        taps["contours"] = contours
        taps["rects"] = rects
        taps["passing"] = passing
        taps["box_img"] = box_img
    return detection, box_img


def _finish(equ, box_img, detection, houghMethod, nlinesInSet, dro, thetaTresh, lineSetTresh, taps):
    if not detection:
        return (False, None)
    equhough = cv2.HoughLines(equ, houghMethod, np.pi / 180, 1)
    boxhough = cv2.HoughLines(box_img, houghMethod, np.pi / 180, 1)
    if taps is not None:
        taps["lines_equ"] = equhough
        taps["lines_box"] = boxhough
    if check_theta(equhough, boxhough, nlinesInSet, dro, thetaTresh, lineSetTresh):
        return (False, None)
    return (True, dictify_hough(equ.shape, equhough[0][0]))


def bright_pass(img, lwTresh, thetaTresh, dilateKernel, contoursMode, contoursMethod,
                minAreaRectMinLen, houghMethod, nlinesInSet, lineSetTresh, dro, debug=False, taps=None):
    """processfield.py:291-388 (clip in place :342, convertScaleAbs :346, equalizeHist :347,
    dilate :354, rect fit :361, two HoughLines :370-371, check_theta :380, dictify :384)."""
    img[img < 0] = 0
    gray = cv2.convertScaleAbs(img)
    equ0 = cv2.equalizeHist(gray)
    equ = cv2.dilate(equ0, dilateKernel)
    if taps is not None:
        taps.update(gray=gray, equ=equ0, morph=equ)
    det, box_img = fit_rects(equ, contoursMode, contoursMethod, minAreaRectMinLen, lwTresh, taps)
    return _finish(equ, box_img, det, houghMethod, nlinesInSet, dro, thetaTresh, lineSetTresh, taps)


def dim_pass(img, minFlux, addFlux, lwTresh, thetaTresh, erodeKernel, dilateKernel, contoursMode,
             contoursMethod, minAreaRectMinLen, houghMethod, nlinesInSet, dro, lineSetTresh,
             debug=False, taps=None):
    """processfield.py:391-506 (threshold/offset in place :453-454, erode :464, dilate :471)."""
    img[img < minFlux] = 0
    img[img > 0] += addFlux
    gray = cv2.convertScaleAbs(img)
    equ0 = cv2.equalizeHist(gray)
    opened = cv2.erode(equ0, erodeKernel)
    equ = cv2.dilate(opened, dilateKernel)
    if taps is not None:
        taps.update(gray=gray, equ=equ0, eroded=opened, morph=equ)
    det, box_img = fit_rects(equ, contoursMode, contoursMethod, minAreaRectMinLen, lwTresh, taps)
    return _finish(equ, box_img, det, houghMethod, nlinesInSet, dro, thetaTresh, lineSetTresh, taps)


def process_frame(img, cat, filt, params_bright=None, params_dim=None, params_removestars=None,
                  taps=None):
    """detecttrails.py:119-131: remove_stars (in place) -> cv2.flip(img, 0) -> bright -> dim.

    ``img`` is the un-flipped float32 frame as read from FITS and is modified in place up to the
    flip.  Returns (detected, pass_index 0|1|-1, result dict|None).
    """
    pb = dict(DEFAULT_BRIGHT if params_bright is None else params_bright)
    pd = dict(DEFAULT_DIM if params_dim is None else params_dim)
    pr = dict(DEFAULT_REMOVESTARS if params_removestars is None else params_removestars)
    rects = star_rects(cat, filt, **pr) if cat is not None else []
    blot(img, rects)
    work = cv2.flip(img, 0)
    tb = {} if taps is not None else None
    td = {} if taps is not None else None
    if taps is not None:
        taps["star_rects"] = rects
        taps["masked"] = work.copy()
        taps["bright"] = tb
        taps["dim"] = td
    det, res = bright_pass(work, taps=tb, **pb)
    if det:
        return True, 0, res
    det, res = dim_pass(work, taps=td, **pd)
    if det:
        return True, 1, res
    return False, -1, None


def result_line(run, camcol, filt, field, header, res):
    """detecttrails.py:115-117,127,131.  Only the first fragment is an f-string in the reference;
    the seven brace groups of the other two fragments are written out literally."""
    head = (f"{run} {camcol} {filt} {field} {header['TAI']} {header['CRPIX1']} "
            "{h['CRPIX2']} {h['CRVAL1']} {h['CRVAL2']} {h['CD1_1']} "
            "{h['CD1_2']} {h['CD2_1']} {h['CD2_2']} ")
    return head + f"{res['x1']} {res['y1']} {res['x2']} {res['y2']}\n"
