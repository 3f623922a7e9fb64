"""The Hough vote kernel evaluates cvRound(x * cos + y * sin) without the XU conversions (k_hough.cuh::hough_r):
(float)x by an exponent-biased integer add and rint() by the 1.5 * 2^23 magic add.  NumPy float32 arithmetic is IEEE
round-to-nearest-even like the device's __fadd_rn / __fmul_rn, so the identities are checked here, on the CPU, over the
ranges the host guard (hough_magic_ok) admits: 0 <= x < 2^23 and |v| < 2^22."""
import numpy as np


def _i2f_magic(x):
    return (np.uint32(0x4B000000) + x.astype(np.uint32)).view(np.float32) - np.float32(8388608.0)


def _rint_magic(v):
    return (v + np.float32(12582912.0)).view(np.int32) - np.int32(0x4B400000)


def test_int_to_float_identity():
    rng = np.random.default_rng(0)
    x = np.concatenate([np.arange(0, 70000), rng.integers(0, 1 << 23, 2_000_000), np.array([(1 << 23) - 1, (1 << 23) - 2])]).astype(np.int64)
    assert np.array_equal(_i2f_magic(x).view(np.uint32), x.astype(np.float32).view(np.uint32))


def test_rint_identity_including_ties_and_negatives():
    rng = np.random.default_rng(1)
    lim = 4194300.0
    v = np.concatenate([
        rng.uniform(-lim, lim, 3_000_000), rng.uniform(-40, 40, 1_000_000),
        np.arange(-100000, 100000) + 0.5,                   # exact ties: round half to even like cvRound (cvtss2si)
        np.arange(-5000, 5000) / 16.0, np.array([lim, -lim, 0.0, -0.0, 0.49999997, -0.49999997])]).astype(np.float32)
    assert np.array_equal(_rint_magic(v), np.rint(v).astype(np.int32))


def test_full_expression_matches_plain_float32():
    """r = rint(float(x) * c + ys) through the magic path == the plain float32 expression OpenCV evaluates."""
    rng = np.random.default_rng(2)
    for rho in (20.0, 5.0, 1.0, 0.5):
        ang = rng.uniform(0, np.pi, 4096)
        c = (np.cos(ang) / rho).astype(np.float32)
        s = (np.sin(ang) / rho).astype(np.float32)
        x = rng.integers(0, 4096, 4096)
        y = rng.integers(0, 4096, 4096).astype(np.float32)
        ys = y * s
        plain = np.rint(x.astype(np.float32) * c + ys).astype(np.int32)
        magic = _rint_magic(_i2f_magic(x) * c + ys)
        assert np.array_equal(plain, magic)
