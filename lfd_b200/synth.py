"""Seeded synthetic SDSS-shaped inputs (frames, photoObj catalogs, a runList.par tree).

Shapes and statistics follow SURVEY.md section 8(d): frames are float32 (1489, 2048) in
sky-subtracted nanomaggy-like units (Gaussian sky noise sigma 0.025), stars are Gaussian PSFs
(sigma 1.5 px, mag U(14, 22), flux = 10**((22.5 - mag) / 2.5)), trails have a Gaussian cross
section.  photoObj tables carry the eight columns the reference reads
(/root/reference/lfd/detecttrails/removestars.py:97-104).

Nothing in here is on the measured path; it feeds the oracle, the tests and bench.py alike.
"""
import os

import numpy as np

from . import fitsio_lite

__all__ = ["FRAME_H", "FRAME_W", "FILTERS", "make_frame", "make_catalog", "frame_seed",
           "make_case", "write_sdss_tree", "DEFAULT_HEADER"]

FRAME_H, FRAME_W = 1489, 2048
FILTERS = ("u", "g", "r", "i", "z")

DEFAULT_HEADER = {
    "TAI": 4575925956.48, "CRPIX1": 1025.0, "CRPIX2": 745.0, "CRVAL1": 52.7013, "CRVAL2": -0.8272,
    "CD1_1": 4.7e-05, "CD1_2": 9.9e-05, "CD2_1": 9.9e-05, "CD2_2": -4.7e-05,
}


def frame_seed(run, camcol, filter, field):
    """Deterministic seed f(run, camcol, filter, field)."""
    fi = FILTERS.index(filter) if isinstance(filter, str) else int(filter)
    return (int(run) * 1000003 + int(camcol) * 10007 + fi * 1009 + int(field)) & 0x7FFFFFFF


def _add_stars(img, xs, ys, flux, sigma):
    h, w = img.shape
    r = int(np.ceil(5 * sigma))
    ax = np.arange(-r, r + 1, dtype=np.float32)
    norm = np.float32(1.0 / (2 * np.pi * sigma * sigma))
    for x, y, f in zip(xs, ys, flux):
        xi, yi = int(round(x)), int(round(y))
        x0, x1 = max(xi - r, 0), min(xi + r + 1, w)
        y0, y1 = max(yi - r, 0), min(yi + r + 1, h)
        if x0 >= x1 or y0 >= y1:
            continue
        gx = np.exp(-0.5 * ((np.arange(x0, x1, dtype=np.float32) - np.float32(x)) / sigma) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1, dtype=np.float32) - np.float32(y)) / sigma) ** 2)
        img[y0:y1, x0:x1] += (np.float32(f) * norm) * np.outer(gy, gx).astype(np.float32)
    del ax


def _add_trail(img, p0, p1, sigma, peak):
    """Gaussian-profile line segment from p0=(x,y) to p1=(x,y), float32 arithmetic."""
    h, w = img.shape
    x0, y0 = p0
    x1, y1 = p1
    dx, dy = x1 - x0, y1 - y0
    L = float(np.hypot(dx, dy))
    if L == 0:
        return
    ux, uy = dx / L, dy / L
    pad = int(np.ceil(5 * sigma)) + 1
    ya, yb = max(int(min(y0, y1)) - pad, 0), min(int(max(y0, y1)) + pad + 1, h)
    xa, xb = max(int(min(x0, x1)) - pad, 0), min(int(max(x0, x1)) + pad + 1, w)
    if ya >= yb or xa >= xb:
        return
    yy, xx = np.mgrid[ya:yb, xa:xb].astype(np.float32)
    t = (xx - x0) * ux + (yy - y0) * uy
    d = -(xx - x0) * uy + (yy - y0) * ux
    prof = np.exp(-0.5 * (d / np.float32(sigma)) ** 2)
    inside = (t >= 0) & (t <= L)
    img[ya:yb, xa:xb] += np.where(inside, np.float32(peak) * prof, np.float32(0)).astype(np.float32)


def make_frame(seed, n_stars=300, trails=(), sky_sigma=0.025, psf_sigma=1.5, h=FRAME_H, w=FRAME_W,
               mag_range=(14.0, 22.0)):
    """Return ``(img float32 (h, w), stars)`` where stars = dict(x, y, mag) arrays.

    ``trails`` is a sequence of dicts ``{p0:(x,y), p1:(x,y), sigma:float, peak:float}``.
    """
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((h, w), dtype=np.float32) * np.float32(sky_sigma)
    xs = rng.uniform(0, w, n_stars)
    ys = rng.uniform(0, h, n_stars)
    mags = rng.uniform(mag_range[0], mag_range[1], n_stars)
    flux = 10.0 ** ((22.5 - mags) / 2.5)
    _add_stars(img, xs, ys, flux, psf_sigma)
    for t in trails:
        _add_trail(img, t["p0"], t["p1"], t.get("sigma", 3.0), t.get("peak", 3.0))
    return img, {"x": xs, "y": ys, "mag": mags}


def make_catalog(seed, stars, extra_fake=10):
    """photoObj-like columns for the given stars (all five bands share the centre)."""
    rng = np.random.default_rng(seed ^ 0x5EED)
    n = len(stars["x"]) + extra_fake
    rowc = np.empty((n, 5), np.float32)
    colc = np.empty((n, 5), np.float32)
    psfmag = np.empty((n, 5), np.float32)
    ns = len(stars["x"])
    rowc[:ns] = (stars["y"][:, None] + rng.normal(0, 0.2, (ns, 5))).astype(np.float32)
    colc[:ns] = (stars["x"][:, None] + rng.normal(0, 0.2, (ns, 5))).astype(np.float32)
    psfmag[:ns] = (stars["mag"][:, None] + rng.normal(0, 0.4, (ns, 5))).astype(np.float32)
    # fake single-band detections (large colour spread -> exercised by magcount)
    rowc[ns:] = rng.uniform(0, FRAME_H, (extra_fake, 1)).astype(np.float32)
    colc[ns:] = rng.uniform(0, FRAME_W, (extra_fake, 1)).astype(np.float32)
    psfmag[ns:] = rng.uniform(14, 30, (extra_fake, 5)).astype(np.float32)
    # a few objects with outliers in 1..4 bands
    nout = min(ns, 12)
    for k in range(nout):
        bands = rng.choice(5, size=1 + k % 4, replace=False)
        psfmag[k, bands] += np.float32(8.0)
    petro = np.where(rng.random((n, 5)) < 0.3, np.float32(-9999.0),
                     rng.uniform(1, 20, (n, 5)).astype(np.float32)).astype(np.float32)
    petro[rng.random(n) < 0.05] = np.float32(40.0)  # > maxxy after scaling -> default size
    nobs = rng.integers(1, 4, n).astype(np.int32)
    ndet = nobs.copy()
    mism = rng.random(n) < 0.08
    ndet[mism] = np.maximum(nobs[mism] - 1, 0)
    return {
        "OBJC_TYPE": np.full(n, 6, np.int32),
        "TYPE": np.full((n, 5), 6, np.int32),
        "ROWC": rowc, "COLC": colc, "PETROTH90": petro, "PSFMAG": psfmag,
        "NOBSERVE": nobs, "NDETECT": ndet,
    }


_KINDS = ("sparse", "trail", "dense", "satellite", "dense_trail", "faint_trail", "empty", "trail_var", "trail_axis",
          "dense_heavy")


def _random_trail(rng, angle=None):
    """One trail of SURVEY.md 8(d) config 3: any angle, width 2-15 px (FWHM of the Gaussian cross section), peak
    0.3-50 (log-uniform), a random chord of the frame at least 600 px long."""
    ang = rng.uniform(0, np.pi) if angle is None else angle
    width = rng.uniform(2.0, 15.0)
    peak = float(np.exp(rng.uniform(np.log(0.3), np.log(50.0))))
    cx, cy = rng.uniform(0.25 * FRAME_W, 0.75 * FRAME_W), rng.uniform(0.25 * FRAME_H, 0.75 * FRAME_H)
    L = rng.uniform(600.0, 2600.0)
    return {"p0": (cx - L / 2 * np.cos(ang), cy - L / 2 * np.sin(ang)),
            "p1": (cx + L / 2 * np.cos(ang), cy + L / 2 * np.sin(ang)),
            "sigma": float(width / 2.3548), "peak": peak}


def make_case(kind, seed):
    """Named frame recipes used across tests/bench.  Returns (img, catalog)."""
    rng = np.random.default_rng(seed ^ 0xABCDEF)
    if kind == "sparse":
        img, st = make_frame(seed, n_stars=int(rng.integers(100, 500)))
    elif kind == "trail":
        img, st = make_frame(seed, n_stars=300, trails=[
            {"p0": (100, 200), "p1": (1900, 1300), "sigma": 3.0, "peak": 3.0}])
    elif kind == "faint_trail":
        ang = rng.uniform(0, np.pi)
        cx, cy, L = FRAME_W / 2, FRAME_H / 2, 1800.0
        img, st = make_frame(seed, n_stars=200, trails=[
            {"p0": (cx - L / 2 * np.cos(ang), cy - L / 2 * np.sin(ang)),
             "p1": (cx + L / 2 * np.cos(ang), cy + L / 2 * np.sin(ang)),
             "sigma": float(rng.uniform(2, 6)), "peak": float(rng.uniform(0.08, 0.3))}])
    elif kind == "dense":
        img, st = make_frame(seed, n_stars=int(rng.integers(3000, 6000)))
    elif kind == "dense_trail":
        img, st = make_frame(seed, n_stars=3000, trails=[
            {"p0": (50, 1400), "p1": (2000, 100), "sigma": 4.0, "peak": 8.0}])
    elif kind == "satellite":
        off = float(rng.uniform(25, 60))
        img, st = make_frame(seed, n_stars=300, trails=[
            {"p0": (0, 300), "p1": (2047, 900), "sigma": 2.5, "peak": 5.0},
            {"p0": (0, 300 + off), "p1": (2047, 900 + off), "sigma": 2.5, "peak": 5.0}])
    elif kind == "empty":
        img, st = make_frame(seed, n_stars=0)
    elif kind == "trail_var":
        img, st = make_frame(seed, n_stars=int(rng.integers(100, 500)), trails=[_random_trail(rng)])
    elif kind == "trail_axis":
        # near-horizontal / near-vertical trails: 0 or 90 degrees +- 1 degree
        base = (0.0, np.pi / 2)[int(rng.integers(0, 2))]
        img, st = make_frame(seed, n_stars=int(rng.integers(100, 500)),
                             trails=[_random_trail(rng, base + np.deg2rad(rng.uniform(-1.0, 1.0)))])
    elif kind == "dense_heavy":
        img, st = make_frame(seed, n_stars=int(rng.integers(6000, 10001)))
    else:
        raise ValueError("unknown case kind %r (known: %s)" % (kind, ", ".join(_KINDS)))
    return img, make_catalog(seed, st)


def case_for_frame(run, camcol, filter, field):
    """Config-3 mix: ~7 % trails (any angle incl. near-horizontal / near-vertical, width 2-15 px, peak
    0.3-50, plus faint ones), ~2 % satellites (two parallel trails), ~10 % dense fields (3 000-10 000 stars), rest
    sparse (100-500 stars)."""
    seed = frame_seed(run, camcol, filter, field)
    u = np.random.default_rng(seed ^ 0x77).random()
    if u < 0.01:
        kind = "trail"
    elif u < 0.04:
        kind = "trail_var"
    elif u < 0.055:
        kind = "trail_axis"
    elif u < 0.07:
        kind = "faint_trail"
    elif u < 0.09:
        kind = "satellite"
    elif u < 0.14:
        kind = "dense"
    elif u < 0.19:
        kind = "dense_heavy"
    else:
        kind = "sparse"
    return kind, seed


def write_sdss_tree(root, run, camcol, fields, filters=FILTERS, rerun="301", kinds=None,
                    startfield=None, endfield=None):
    """Write a minimal $BOSS_PHOTOOBJ / $PHOTO_REDUX tree the drop-in (and the reference) can read.

    Returns dict(bosspath, photoobjpath, photoreduxpath, frames={(filter, field): (img, catalog)}).
    Layout per /root/reference/lfd/detecttrails/sdss/share/sdssFileTypes.par:43 (frame) and :75
    (photoObj); runList.par per sdss/files.py:623-635.
    """
    boss = os.path.join(root, "boss")
    photoobj = os.path.join(boss, "photoObj")
    redux = os.path.join(boss, "photo", "redux")
    fdir = os.path.join(photoobj, "frames", rerun, str(run), str(camcol))
    odir = os.path.join(photoobj, rerun, str(run), str(camcol))
    for d in (fdir, odir, redux):
        os.makedirs(d, exist_ok=True)
    fields = list(fields)
    sf = min(fields) if startfield is None else startfield
    ef = max(fields) if endfield is None else endfield
    with open(os.path.join(redux, "runList.par"), "w") as f:
        f.write("typedef struct {\n int run;\n char rerun[];\n int exist;\n int done;\n int calib;\n"
                " int startfield;\n int endfield;\n char machine[];\n char disk[];\n} RUNDATA;\n\n")
        f.write("RUNDATA %d %s 1 1 1 %d %d synth /synth\n" % (run, rerun, sf, ef))
    out = {}
    for field in fields:
        cat_written = False
        for flt in filters:
            if kinds is not None:
                kind = kinds[(flt, field)] if isinstance(kinds, dict) else kinds
                seed = frame_seed(run, camcol, flt, field)
            else:
                kind, seed = case_for_frame(run, camcol, flt, field)
            img, cat = make_case(kind, seed)
            hdr = dict(DEFAULT_HEADER)
            hdr["TAI"] = DEFAULT_HEADER["TAI"] + field * 36.0
            fitsio_lite.write_image(
                os.path.join(fdir, "frame-%s-%06d-%d-%04d.fits" % (flt, run, camcol, field)), img, hdr)
            if not cat_written:
                fitsio_lite.write_bintable(
                    os.path.join(odir, "photoObj-%06d-%d-%04d.fits" % (run, camcol, field)), cat)
                cat_written = True
                field_cat = cat
            out[(flt, field)] = (img, field_cat)
    return {"bosspath": boss, "photoobjpath": photoobj, "photoreduxpath": redux, "frames": out}
