"""Size-independent properties at the BASELINE frame size (2048x1489) and batch sizes the oracle would take minutes
for: determinism, batch-position / batch-size invariance, Hough vote conservation and linearity, label invariants,
raw big-endian input == native input, serialised passes == overlapped passes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from lfd_b200 import synth
from oracle import ref_pipeline as rp


def _pool(n):
    from lfd_b200.removestars import star_rects
    kinds = ["trail", "sparse", "dense", "satellite", "faint_trail", "dense_trail", "empty", "sparse"]
    frames, rects = [], []
    for i in range(n):
        img, cat = synth.make_case(kinds[i % len(kinds)], 500 + i)
        frames.append(img)
        rects.append(star_rects(cat, "r", img.shape, **rp.DEFAULT_REMOVESTARS))
    return np.stack(frames), rects


def test_batch_invariance_determinism_and_input_paths():
    from lfd_b200 import _lib
    n = 12
    frames, rects = _pool(n)
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=n)
    h1 = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=1)
    try:
        h.set_params(pb, pd); h1.set_params(pb, pd)
        h.submit(frames, rects); a = [bytes(r) for r in h.wait()]
        # determinism (graph replay, atomics in CCL / Hough / contour allocation do not leak into results)
        for _ in range(3):
            h.submit(frames, rects)
            assert [bytes(r) for r in h.wait()] == a
        # position in the batch does not matter
        perm = np.random.default_rng(1).permutation(n)
        h.submit(frames[perm], [rects[i] for i in perm])
        b = [bytes(r) for r in h.wait()]
        assert [b[list(perm).index(i)] for i in range(n)] == a
        # batch of one == batch of n
        for i in (0, 2, 5):
            h1.submit(frames[i:i + 1], [rects[i]])
            assert bytes(h1.wait()[0]) == a[i]
        # passes on one stream == passes overlapped on two streams
        h.submit(frames, rects, flags=_lib.SERIAL_PASSES)
        assert [bytes(r) for r in h.wait()] == a
        # raw big-endian FITS payload (byte-swapped on the device) == native floats
        be = frames.astype(">f4").view(np.uint32)
        h.submit(be, rects, flags=_lib.INPUT_BIGENDIAN)
        assert [bytes(r) for r in h.wait()] == a
        assert sum(r.detected for r in h._results[:n]) >= 3
    finally:
        h.close(); h1.close()


def test_hough_conservation_and_linearity():
    """Every non-zero pixel votes once per angle; votes of disjoint images add up."""
    from lfd_b200 import _lib
    H, W = synth.FRAME_H, synth.FRAME_W
    rng = np.random.default_rng(9)
    h = _lib.Handle(H, W, max_batch=1)
    try:
        a = (rng.random((H, W)) < 0.02).astype(np.uint8) * 255
        b = (rng.random((H, W)) < 0.05).astype(np.uint8) * 200
        b[a > 0] = 0
        for rho, theta in ((20, np.pi / 180), (3, np.pi / 360)):
            _, acc_a = h.hough_lines(a, rho, theta, 1, want_accum=True, max_lines=16)
            _, acc_b = h.hough_lines(b, rho, theta, 1, want_accum=True, max_lines=16)
            _, acc_ab = h.hough_lines(a | b, rho, theta, 1, want_accum=True, max_lines=16)
            na = acc_a.shape[0] - 2
            assert acc_a.sum() == int((a > 0).sum()) * na
            assert acc_ab.sum() == int(((a | b) > 0).sum()) * na
            assert np.array_equal(acc_a.astype(np.int64) + acc_b, acc_ab)
            assert (acc_a[0] == 0).all() and (acc_a[-1] == 0).all() and (acc_a[:, 0] == 0).all() and (acc_a[:, -1] == 0).all()
    finally:
        h.close()


def test_label_and_edge_invariants_full_frame():
    """fg labels: constant on 8-connected neighbours, label = raster-first pixel of the component; Canny output is a
    subset of the NMS candidates and contains every strong pixel; dilation never lowers, erosion never raises."""
    from lfd_b200 import _lib
    img, _ = synth.make_case("dense_trail", 77)
    work = np.ascontiguousarray(img[::-1])
    h = _lib.Handle(synth.FRAME_H, synth.FRAME_W, max_batch=1)
    try:
        h.set_params(dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM))
        for pass_ in (0, 1):
            w = work.copy()
            if pass_ == 1:
                w[w < 0] = 0
            h.run_pass(pass_, w, flags=_lib.KEEP_TAPS, writeback=False)
            canny = h.stage(0, pass_, "canny") > 0
            nms = h.stage(0, pass_, "nms")
            lab = h.stage(0, pass_, "fg_labels")
            equ, morph = h.stage(0, pass_, "equ"), h.stage(0, pass_, "morph")
            assert not (canny & (nms == 0)).any()
            assert not ((nms == 2) & ~canny).any()
            assert ((lab >= 0) == canny).all()
            # 8-neighbours that are both edges carry the same label
            for dy, dx in ((0, 1), (1, -1), (1, 0), (1, 1)):
                a = lab[max(0, -dy):lab.shape[0] - max(0, dy), max(0, -dx):lab.shape[1] - max(0, dx)]
                b = lab[max(0, dy):, max(0, dx):][:a.shape[0], :a.shape[1]]
                both = (a >= 0) & (b >= 0)
                assert (a[both] == b[both]).all()
            ids = np.unique(lab[lab >= 0])
            flat = lab.ravel()
            assert (flat[ids] == ids).all()                      # the label is the index of a pixel of the component ...
            first = np.full(flat.max() + 1, -1, np.int64)
            idx = np.flatnonzero(flat >= 0)
            first[flat[idx][::-1]] = idx[::-1]                   # ... namely its raster-first one
            assert (first[ids] == ids).all()
            if pass_ == 0:
                assert (morph >= equ).all()
            else:
                er = h.stage(0, pass_, "eroded")
                assert (er <= equ).all() and (morph >= er).all()
    finally:
        h.close()


def test_prep_walks_large_batches_and_ragged_rows():
    """k_prep keeps at most 16 frames in flight and lets its CTAs walk the rest of the batch; rows whose byte length is
    not a multiple of the 2 KB TMA stage end in a partial stage.  40 distinct small frames (W = 520: one full and one
    32-byte stage per row) with blots, star-like pixels and special values: the flipped / clipped / converted planes
    and the histograms of every frame must equal NumPy's."""
    from lfd_b200 import _lib
    H, W, n = 67, 520, 40
    rng = np.random.default_rng(77)
    frames = rng.normal(0, 0.03, (n, H, W)).astype(np.float32)
    frames[rng.random((n, H, W)) < 0.002] += 9.0                           # bytes >= 2: the staged histogram path
    frames[rng.random((n, H, W)) < 0.0005] = 400.0                         # saturation: the exact conversion path
    frames[3, 5, 7] = np.nan; frames[17, 60, 519] = np.inf; frames[39, 66, 0] = np.uint32(0x5700003c).view(np.float32)
    rects = [np.array([[5 + i % 7, 20 + i % 7, 30 + i, 60 + i], [0, 3, 500, 520]], np.int32) for i in range(n)]
    pb, pd = dict(rp.DEFAULT_BRIGHT), dict(rp.DEFAULT_DIM)
    h = _lib.Handle(H, W, max_batch=n)
    try:
        h.set_params(pb, pd)
        h.submit(frames, rects, flags=_lib.KEEP_TAPS)
        h.wait()
        for f in range(n):
            img = frames[f].copy()
            for r0, r1, c0, c1 in rects[f]:
                img[r0:r1, c0:c1] = 0.0
            t = img[::-1].copy()
            with np.errstate(invalid="ignore"):
                t[t < 0] = 0
                g0 = rp.cv2.convertScaleAbs(t)
                t[t < pd["minFlux"]] = 0
                t[t > 0] += pd["addFlux"]
                g1 = rp.cv2.convertScaleAbs(t)
            assert np.array_equal(h.stage(f, 0, "gray"), g0), f
            assert np.array_equal(h.stage(f, 1, "gray"), g1), f
            assert np.array_equal(h.stage(f, 0, "hist").astype(np.int64), np.bincount(g0.ravel(), minlength=256)), f
            assert np.array_equal(h.stage(f, 1, "hist").astype(np.int64), np.bincount(g1.ravel(), minlength=256)), f
    finally:
        h.close()
