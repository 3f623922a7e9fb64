"""Drop-in for ``lfd.detecttrails.removestars``: catalog filter on the host (vectorised NumPy with the
reference's exact ceil/compare rules), square blots on the device.

Mirrors /root/reference/lfd/detecttrails/removestars.py: ``read_photoObj`` (:63-132) and
``remove_stars`` (:148-233) keep their names, arguments and return values; the deprecated CSV
functions (:19-60, :135-145) are unreachable from ``process_field`` and are not provided.
"""

import numpy as np

from . import _lib, sdssfiles as files

try:  # the reference imports fitsio unconditionally; it is optional here
    import fitsio
except ImportError:  # pragma: no cover - depends on the image
    from . import fitsio_lite as fitsio

__all__ = ["read_photoObj", "remove_stars", "star_rects", "star_rects_batch", "read_photoObj_arrays"]

_BANDS = ("u", "g", "r", "i", "z")


def read_photoObj_arrays(path_to_photoOBJ):
    """The eight columns removestars.py:97-104 reads, as arrays."""
    cols = ("OBJC_TYPE", "TYPE", "ROWC", "COLC", "PETROTH90", "PSFMAG", "NOBSERVE", "NDETECT")
    if hasattr(fitsio, "read_columns"):          # fitsio_lite: converts only these columns
        return fitsio.read_columns(path_to_photoOBJ, cols)
    data, _hdr = fitsio.read(path_to_photoOBJ, header="True")
    return {k: np.asarray(data[k]) for k in cols}


def _ceil_int(a):
    """math.ceil element-wise -> int64; non-finite input raises like math.ceil does (removestars.py:113-130)."""
    a = np.asarray(a, np.float64)
    if not np.all(np.isfinite(a)):
        if np.any(np.isnan(a)):
            raise ValueError("cannot convert float NaN to integer")
        raise OverflowError("cannot convert float infinity to integer")
    return np.ceil(a).astype(np.int64)


def read_photoObj(path_to_photoOBJ):
    """removestars.py:63-132: lists of per-band dicts of ceil'd values plus the raw type/count columns."""
    c = read_photoObj_arrays(path_to_photoOBJ)
    def dicts(a):
        a = _ceil_int(a)
        return [dict(zip(_BANDS, (int(v) for v in row))) for row in a]
    return (dicts(c["ROWC"]), dicts(c["COLC"]), dicts(c["PSFMAG"]), dicts(c["PETROTH90"]),
            c["OBJC_TYPE"], c["TYPE"], c["NOBSERVE"], c["NDETECT"])


def star_rects(cat, _filter, shape, defaultxy, filter_caps, maxxy, pixscale, magcount, maxmagdiff, debug=False):
    """Blot rectangles (row_start, row_stop, col_start, col_stop) on the un-flipped image.

    Filter logic of removestars.py:212-230, vectorised; the slice ``img[x-dxy:x+dxy, y-dxy:y+dxy]``
    (:231) is resolved with Python's slice arithmetic (vectorised ``slice.indices``) so that objects whose start
    goes negative wrap to an empty slice exactly like NumPy's indexing does."""
    b = _BANDS.index(_filter)
    H, W = shape
    n = len(cat["ROWC"])
    if n == 0:
        return np.zeros((0, 4), np.int32)
    # one ceil over the four (n, 5) columns (math.ceil of every band of every object, removestars.py:113-130:
    # a non-finite value anywhere raises, whatever the filter)
    c = _ceil_int(np.stack([cat["ROWC"], cat["COLC"], cat["PSFMAG"], cat["PETROTH90"]]))
    rows, cols, mags, p90 = c[0][:, b], c[1][:, b], c[2], c[3][:, b]
    keep = mags[:, b] < filter_caps[_filter]
    big = (np.abs(mags[:, _PAIR_J] - mags[:, _PAIR_K]) > maxmagdiff).sum(axis=1)      # the 10 band pairs (:217-224)
    keep &= magcount >= big
    keep &= np.asarray(cat["NOBSERVE"]) == np.asarray(cat["NDETECT"])
    dxy = np.full(n, defaultxy, np.int64)
    pos = p90 > 0
    dxy[pos] = (p90[pos] / pixscale).astype(np.int64) + 10     # int() truncation of a positive float
    dxy[dxy > maxxy] = defaultxy
    x, y, dk = cols[keep], rows[keep], dxy[keep]
    # slice.indices() for step 1: a negative bound counts from the end, then both are clamped to [0, length]
    r = np.stack([x - dk, x + dk])
    r = np.clip(np.where(r < 0, r + H, r), 0, H)
    cc = np.stack([y - dk, y + dk])
    cc = np.clip(np.where(cc < 0, cc + W, cc), 0, W)
    ok = (r[0] < r[1]) & (cc[0] < cc[1])
    return np.stack([r[0][ok], r[1][ok], cc[0][ok], cc[1][ok]], axis=1).astype(np.int32).reshape(-1, 4)


_PAIR_J, _PAIR_K = (np.array(v) for v in zip(*[(j, k) for j in range(5) for k in range(j + 1, 5)]))


def star_rects_batch(cats, filters, shape, defaultxy, filter_caps, maxxy, pixscale, magcount, maxmagdiff, debug=False):
    """``star_rects`` for the frames of one GPU batch in a single set of NumPy calls (the per-call interpreter
    overhead, not the arithmetic, is what a loader thread spends its time on at > 1 k frames/s).  ``cats[i]`` is the
    column dict of frame i (or an exception instance: passed through), ``filters[i]`` its filter.  Returns one
    entry per frame: the (n, 4) int32 rectangles, or the exception ``star_rects`` would have raised for it."""
    out = [None] * len(cats)
    idx = [i for i, c in enumerate(cats) if not isinstance(c, BaseException)]
    for i, c in enumerate(cats):
        if isinstance(c, BaseException):
            out[i] = c
    H, W = shape
    if idx:
        try:
            lens = np.array([len(cats[i]["ROWC"]) for i in idx])
            cols4 = np.stack([np.concatenate([np.asarray(cats[i][k], np.float64).reshape(-1, 5) for i in idx])
                              for k in ("ROWC", "COLC", "PSFMAG", "PETROTH90")])
            if not np.all(np.isfinite(cols4)):
                raise ValueError("non-finite catalog value")            # resolved per frame below
            c = np.ceil(cols4).astype(np.int64)
            fid = np.repeat(np.arange(len(idx)), lens)                      # frame (position in idx) of every object
            bo = np.repeat(np.array([_BANDS.index(filters[i]) for i in idx]), lens)
            ar = np.arange(len(fid))
            rows, colsb, mags, p90 = c[0][ar, bo], c[1][ar, bo], c[2], c[3][ar, bo]
            caps = np.array([filter_caps[f] for f in _BANDS], np.float64)
            keep = mags[ar, bo] < caps[bo]
            keep &= magcount >= (np.abs(mags[:, _PAIR_J] - mags[:, _PAIR_K]) > maxmagdiff).sum(axis=1)
            keep &= (np.concatenate([np.asarray(cats[i]["NOBSERVE"]).reshape(-1) for i in idx]) ==
                     np.concatenate([np.asarray(cats[i]["NDETECT"]).reshape(-1) for i in idx]))
            dxy = np.full(len(fid), defaultxy, np.int64)
            pos = p90 > 0
            dxy[pos] = (p90[pos] / pixscale).astype(np.int64) + 10
            dxy[dxy > maxxy] = defaultxy
            x, y, dk, fk = colsb[keep], rows[keep], dxy[keep], fid[keep]
            r = np.stack([x - dk, x + dk]); r = np.clip(np.where(r < 0, r + H, r), 0, H)
            cc = np.stack([y - dk, y + dk]); cc = np.clip(np.where(cc < 0, cc + W, cc), 0, W)
            ok = (r[0] < r[1]) & (cc[0] < cc[1])
            rects = np.stack([r[0][ok], r[1][ok], cc[0][ok], cc[1][ok]], axis=1).astype(np.int32).reshape(-1, 4)
            cuts = np.searchsorted(fk[ok], np.arange(1, len(idx)))          # objects stay in frame order
            for i, part in zip(idx, np.split(rects, cuts)):
                out[i] = part
        except Exception:   # noqa: BLE001 - a bad catalog somewhere in the batch: frame by frame, each with its own error
            for i in idx:
                try:
                    out[i] = star_rects(cats[i], filters[i], shape, defaultxy, filter_caps, maxxy, pixscale, magcount,
                                        maxmagdiff, debug)
                except Exception as e:   # noqa: BLE001
                    out[i] = e
    return out


def remove_stars(img, _run, _camcol, _filter, _field, defaultxy, filter_caps, maxxy, pixscale, magcount,
                 maxmagdiff, debug):
    """removestars.py:148-233: zero a square per catalog object, in place, and return ``img``."""
    cat = read_photoObj_arrays(files.filename("photoObj", run=_run, camcol=_camcol, field=_field))
    rects = star_rects(cat, _filter, img.shape, defaultxy, filter_caps, maxxy, pixscale, magcount, maxmagdiff, debug)
    if len(rects) == 0:
        return img
    work = img
    if img.dtype != np.float32 or not img.flags["C_CONTIGUOUS"]:
        work = np.ascontiguousarray(img, np.float32)
    h = _lib.default_handle(work.shape[0], work.shape[1])
    h.blot(work, rects)
    if work is not img:
        img[...] = work
    return img
