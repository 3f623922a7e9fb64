"""ORACLE (test infrastructure, not product code): frame verdicts of the restated reference pipeline for a list of
frames, on all host cores.  Used by tests/ and by bench.py's --verify / CPU-baseline legs as the CHECKER only.

A verdict is what detecttrails.py:119-131 decides for one frame: (detected, pass index 0 bright | 1 dim | -1,
{"x1","y1","x2","y2"} | None), computed by oracle/ref_pipeline.py::process_frame (the reference's call sequence on cv2).
"""
import multiprocessing as mp
import os


_JOBS = None      # set before the fork so that the workers inherit the frames instead of unpickling 12 MB each


def _one(i):
    import cv2
    cv2.setNumThreads(1)
    from oracle import ref_pipeline as rp
    img, cat, flt, pb, pd, pr = _JOBS[i]
    try:
        return rp.process_frame(img.copy(), cat, flt, pb, pd, pr)
    except Exception as e:   # noqa: BLE001 - the reference logs any per-frame failure to errors.txt (detecttrails.py:133-139)
        return ("err", type(e).__name__, None)


def verdicts(frames, cats, filters, params_bright=None, params_dim=None, params_removestars=None, cores=None):
    """[(detected, pass, result dict | None)] in list order; ``cores`` processes (default: all)."""
    global _JOBS
    cores = cores or os.cpu_count() or 1
    _JOBS = [(f, c, flt, params_bright, params_dim, params_removestars) for f, c, flt in zip(frames, cats, filters)]
    try:
        if cores == 1 or len(_JOBS) == 1:
            return [_one(i) for i in range(len(_JOBS))]
        with mp.get_context("fork").Pool(min(cores, len(_JOBS))) as pool:
            return pool.map(_one, range(len(_JOBS)), chunksize=1)
    finally:
        _JOBS = None


def device_verdict(r, shape):
    """The same triple from one ``lfd_result`` (bright wins; dim counts only when bright found nothing)."""
    from lfd_b200.processfield import result_from_device
    try:
        for p in (0, 1):
            if r.rect_detection[p] >= 0:
                det, out = result_from_device(r, p, shape)
                if det:
                    return (True, p, out)
    except Exception as e:   # noqa: BLE001 - what the drop-in would write to errors.txt
        return ("err", type(e).__name__, None)
    return (False, -1, None)
