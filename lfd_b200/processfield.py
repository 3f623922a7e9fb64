"""Drop-in for ``lfd.detecttrails.processfield`` with the pixel work on the GPU.

Mirrors /root/reference/lfd/detecttrails/processfield.py: same function names, argument names and
order, return values ``(bool, {"x1","y1","x2","y2"} | None)`` and the same in-place mutation of
``img`` (processfield.py:342, :453-454).  The image arithmetic (clip, convertScaleAbs, equalizeHist,
erode/dilate, Canny, contour rectangles, box fill, both HoughLines, check_theta) runs in
liblfd_b200.so; only ``dictify_hough`` stays on the host in NumPy float32, which is exactly what the
reference executes (processfield.py:281-286) and costs microseconds.
"""
import os

import numpy as np

from . import _lib

__all__ = ["process_field_bright", "process_field_dim", "pathBright", "check_theta", "dictify_hough",
           "fit_minAreaRect", "setup_debug", "draw_lines"]

pathBright = None
pathDim = None


def setup_debug():
    """processfield.py:22-33: read DEBUG_PATH into the module globals."""
    global pathBright, pathDim
    try:
        pathBright = os.environ["DEBUG_PATH"]
        pathDim = os.environ["DEBUG_PATH"]
    except KeyError:
        pass


def check_theta(hough1, hough2, navg, dro, thetaTresh, lineSetTresh, debug):
    """Host mirror of processfield.py:36-150 for callers that hold line lists (the device runs the
    same test inside the pipeline, csrc/k_hough.cuh::k_check_theta).  True = reject, None = accept."""
    ro1 = np.zeros((navg, 1)); ro2 = np.zeros((navg, 1))
    theta1 = np.zeros((navg, 1)); theta2 = np.zeros((navg, 1))
    for i in range(navg):
        try:
            ro1[i] = hough1[i][0][0]
            ro2[i] = hough2[i][0][0]
            theta1[i] = hough1[i][0][1]
            theta2[i] = hough2[i][0][1]
        except IndexError:
            pass
    if debug:
        print("RO: Ro_tresh: %s avg(ro1): %s avg(ro2): %s" % (dro, np.average(ro1), np.average(ro2)))
    if abs(np.average(ro1) - np.average(ro2)) > dro:
        return True
    if abs(theta1.max() - theta1.min()) > thetaTresh:
        return True
    if abs(theta2.max() - theta2.min()) > thetaTresh:
        return True
    if np.average(abs(theta1 - theta2)) > lineSetTresh:
        return True


def dictify_hough(shape, houghVals):
    """processfield.py:266-288 (float32 NumPy scalar arithmetic, int() truncation)."""
    rho, theta = houghVals
    n_x, n_y = shape
    x0 = np.cos(theta) * rho
    y0 = np.sin(theta) * rho
    x1 = int(x0 - (n_x + n_y) * np.sin(theta))
    y1 = int(y0 + (n_x + n_y) * np.cos(theta))
    x2 = int(x0 + (n_x + n_y) * np.sin(theta))
    y2 = int(y0 - (n_x + n_y) * np.cos(theta))
    return {"x1": x1, "y1": y1, "x2": x2, "y2": y2}


def _write_png(path, arr, compression=0):
    """Minimal PNG encoder (8-bit gray or BGR colour arrays) so that the debug taps need no image library."""
    import struct
    import zlib
    a = np.ascontiguousarray(arr, np.uint8)
    if a.ndim == 3:
        a = np.ascontiguousarray(a[:, :, ::-1])          # BGR (cv2 convention) -> RGB
        ctype = 2
    else:
        ctype = 0
    hgt, wid = a.shape[:2]
    raw = np.concatenate([np.zeros((hgt, 1), np.uint8), a.reshape(hgt, -1)], axis=1).tobytes()    # filter byte 0 per row

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", wid, hgt, 8, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, int(compression))) + chunk(b"IEND", b""))


def draw_lines(hough, image, nlines, name, path=None, compression=0, color=(255, 0, 0)):
    """processfield.py:153-198: draw the first ``nlines`` Hough lines on a colour copy of ``image`` and save
    ``<path>/<name>.png``.  Debug visualisation only (2-px lines rasterised with NumPy, not cv2.line's exact pixels)."""
    path = pathDim if path is None else path
    n_x, n_y = image.shape
    draw = np.repeat(np.ascontiguousarray(image, np.uint8)[:, :, None], 3, axis=2)
    for houghparams in (hough if hough is not None else [])[:nlines]:
        try:
            rho, theta = houghparams[0]
            x0 = np.cos(theta) * rho
            y0 = np.sin(theta) * rho
            x1 = int(x0 - (n_x + n_y) * np.sin(theta)); y1 = int(y0 + (n_x + n_y) * np.cos(theta))
            x2 = int(x0 + (n_x + n_y) * np.sin(theta)); y2 = int(y0 - (n_x + n_y) * np.cos(theta))
            n = max(abs(x2 - x1), abs(y2 - y1)) + 1
            t = np.linspace(0.0, 1.0, n)
            xs = np.rint(x1 + (x2 - x1) * t).astype(np.int64)
            ys = np.rint(y1 + (y2 - y1) * t).astype(np.int64)
            for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
                ok = (xs + dx >= 0) & (xs + dx < n_y) & (ys + dy >= 0) & (ys + dy < n_x)
                draw[ys[ok] + dy, xs[ok] + dx] = color
        except Exception:   # noqa: BLE001 - the reference ignores lines it cannot draw
            pass
    _write_png(os.path.join(path, name + ".png"), draw, compression)


def _dump_debug(handle, pass_, bright, params):
    """The reference's debug taps (processfield.py:349-378, :459-496): the same PNG file names in $DEBUG_PATH, and the
    check_theta printout (:104-132) through the host mirror of that function."""
    path = pathBright if bright else pathDim
    if path is None:
        return
    names = ([("1equBRIGHT", "equ"), ("2dilateBRIGHT", "morph"), ("3contoursBRIGHT", "box")] if bright else
             [("6equDIM", "equ"), ("7erodedDIM", "eroded"), ("8openedDIM", "morph"), ("9contoursDIM", "box")])
    imgs = {}
    for fname, stage in names:
        try:
            imgs[stage] = handle.stage(0, pass_, stage)
            _write_png(os.path.join(path, fname + ".png"), imgs[stage])
        except _lib.LfdError:
            pass
    try:
        leq, lbox = handle.stage(0, pass_, "lines_equ"), handle.stage(0, pass_, "lines_box")
    except _lib.LfdError:
        return
    if len(leq) and len(lbox) and "morph" in imgs and "box" in imgs:
        nl = int(params["nlinesInSet"])
        if bright:
            draw_lines(lbox, imgs["box"], nl, "4boxhoughBRIGHT", path=path)
            draw_lines(leq, imgs["morph"], nl, "5equhoughBRIGHT", path=path)
        else:
            draw_lines(leq, imgs["morph"], nl, "10equhoughDIM", path=path)
            draw_lines(lbox, imgs["box"], nl, "11boxhoughDIM", path=path)
        check_theta(leq, lbox, nl, params["dro"], params["thetaTresh"], params["lineSetTresh"], True)


def result_from_device(r, pass_, shape):
    """lfd_result -> the reference's (bool, dict|None), raising what the reference would raise."""
    if r.status & _lib.FRAME_OVERFLOW:
        raise _lib.LfdError(_lib.LFD_E_CAPACITY, "per-frame work list overflow (raise max_runs/max_components)")
    if r.rect_detection[pass_] == 1:
        if r.status & (_lib.FRAME_NO_LINES_EQU | _lib.FRAME_NO_LINES_BOX):
            # cv2.HoughLines returned None; check_theta then fails on hough[i] (processfield.py:97)
            raise TypeError("'NoneType' object is not subscriptable")
        if r.rejected[pass_]:
            return (False, None)
        return (True, dictify_hough(shape, (np.float32(r.top_equ[pass_][0][0]), np.float32(r.top_equ[pass_][0][1]))))
    return (False, None)


def _run(pass_, img, params, dim):
    if not isinstance(img, np.ndarray) or img.ndim != 2:
        raise TypeError("img must be a 2-D numpy array")
    work = img
    if img.dtype != np.float32 or not img.flags["C_CONTIGUOUS"]:
        work = np.ascontiguousarray(img, np.float32)
    h = _lib.default_handle(work.shape[0], work.shape[1])
    # the other pass's params are irrelevant for a single-pass call; reuse this one for both slots
    from .detecttrails import default_params
    pb, pd, _ = default_params()
    if dim:
        pd = params
    else:
        pb = params
    h.set_params(pb, pd)
    debug = bool(params.get("debug", False))
    r = h.run_pass(pass_, work, flags=(_lib.KEEP_TAPS | _lib.FULL_LINES) if debug else 0, writeback=True)
    if work is not img:
        img[...] = work      # the reference mutates its argument in place
    if debug:
        _dump_debug(h, pass_, not dim, params)
    return result_from_device(r, pass_, img.shape)


def process_field_bright(img, lwTresh, thetaTresh, dilateKernel, contoursMode, contoursMethod,
                         minAreaRectMinLen, houghMethod, nlinesInSet, lineSetTresh, dro, debug):
    """processfield.py:291-388 on the GPU.  ``img`` (float32 2-D) is clipped in place like the original."""
    params = dict(lwTresh=lwTresh, thetaTresh=thetaTresh, dilateKernel=dilateKernel, contoursMode=contoursMode,
                  contoursMethod=contoursMethod, minAreaRectMinLen=minAreaRectMinLen, houghMethod=houghMethod,
                  nlinesInSet=nlinesInSet, lineSetTresh=lineSetTresh, dro=dro, debug=debug)
    return _run(_lib.PASS_BRIGHT, img, params, False)


def process_field_dim(img, minFlux, addFlux, lwTresh, thetaTresh, erodeKernel, dilateKernel, contoursMode,
                      contoursMethod, minAreaRectMinLen, houghMethod, nlinesInSet, dro, lineSetTresh, debug):
    """processfield.py:391-506 on the GPU."""
    params = dict(minFlux=minFlux, addFlux=addFlux, lwTresh=lwTresh, thetaTresh=thetaTresh, erodeKernel=erodeKernel,
                  dilateKernel=dilateKernel, contoursMode=contoursMode, contoursMethod=contoursMethod,
                  minAreaRectMinLen=minAreaRectMinLen, houghMethod=houghMethod, nlinesInSet=nlinesInSet,
                  dro=dro, lineSetTresh=lineSetTresh, debug=debug)
    return _run(_lib.PASS_DIM, img, params, True)


def fit_minAreaRect(img, contoursMode, contoursMethod, minAreaRectMinLen, lwTresh, debug):
    """processfield.py:201-263 on the GPU: ``(detection, box_img)`` for a uint8 image (Canny(0,255) -> findContours ->
    minAreaRect filter -> fillPoly of the int-truncated boxPoints)."""
    if not isinstance(img, np.ndarray) or img.ndim != 2:
        raise TypeError("img must be a 2-D numpy array")
    h = _lib.default_handle(img.shape[0], img.shape[1])
    if h._params_key is None:
        from .detecttrails import default_params
        pb, pd, _ = default_params()
        h.set_params(pb, pd)
    detection, box_img = h.fit_min_area_rect(img, contoursMode, contoursMethod, minAreaRectMinLen, lwTresh)
    if debug and pathBright is not None:
        _write_png(os.path.join(pathBright, "3contours.png"), box_img)
    return detection, box_img
