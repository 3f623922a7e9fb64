"""Drop-in for ``lfd.detecttrails.detecttrails``: ``DetectTrails`` and ``process_field``.

Mirrors /root/reference/lfd/detecttrails/detecttrails.py: constructor kwargs and defaults (:199-267,
including the quirk that ``params_dim=`` / ``params_removestars=`` kwargs land in ``params_bright``,
:253-256), frame-set selection ``_load`` (:290-342), loop order and ranges of ``process`` (:344-407,
``run-camcol`` strides fields by 50, ``endfield`` excluded), the results line (:115-117,127,131 - seven
of its brace groups are literal text in the reference and stay literal here), per-frame error capture
into errors.txt (:133-139) and the bz2 unpack path (:81-111).

What differs is where the work runs: frames are decoded on the host, grouped into batches, copied with
pinned cudaMemcpyAsync and pushed through liblfd_b200.so (one handle per GPU); results come back as one
small struct per frame and are written in the reference's frame order.
"""
import bz2
import os
import traceback

import numpy as _np

from . import _lib, sdssfiles as files
from .processfield import result_from_device, setup_debug
from .removestars import read_photoObj_arrays, star_rects, star_rects_batch

try:
    import fitsio
except ImportError:  # pragma: no cover - depends on the image
    from . import fitsio_lite as fitsio

# cv2 enum values the reference re-exports (detecttrails.py:14-18); cv2 itself is not needed here
RETR_EXTERNAL, RETR_LIST, RETR_CCOMP, RETR_TREE = 0, 1, 2, 3
CHAIN_APPROX_NONE, CHAIN_APPROX_SIMPLE, CHAIN_APPROX_TC89_L1, CHAIN_APPROX_TC89_KCOS = 1, 2, 3, 4

__all__ = ["DetectTrails", "process_field", "process_fields", "default_params"]


def default_params():
    """The three default dicts of detecttrails.py:202-239."""
    params_bright = {
        "lwTresh": 5, "thetaTresh": 0.15, "dilateKernel": _np.ones((4, 4), _np.uint8),
        "contoursMode": RETR_LIST, "contoursMethod": CHAIN_APPROX_NONE, "minAreaRectMinLen": 1,
        "houghMethod": 20, "nlinesInSet": 3, "lineSetTresh": 0.15, "dro": 25, "debug": False}
    params_dim = {
        "minFlux": 0.02, "addFlux": 0.5, "lwTresh": 5, "thetaTresh": 0.15,
        "erodeKernel": _np.ones((3, 3), _np.uint8), "dilateKernel": _np.ones((9, 9), _np.uint8),
        "contoursMode": RETR_LIST, "contoursMethod": CHAIN_APPROX_NONE, "minAreaRectMinLen": 1,
        "houghMethod": 20, "nlinesInSet": 3, "lineSetTresh": 0.15, "dro": 20, "debug": False}
    params_removestars = {
        "pixscale": 0.396, "defaultxy": 20, "maxxy": 60,
        "filter_caps": {"u": 22.0, "g": 22.2, "r": 22.2, "i": 21.3, "z": 20.5},
        "magcount": 3, "maxmagdiff": 3, "debug": False}
    return params_bright, params_dim, params_removestars


# Opt-in: write the seven header values instead of the reference's literal "{h['CRPIX2']}" ...
# placeholders, so that lfd.results.utils.parse_result_row (results/utils.py:185-210) can ingest the file.
# Default False = byte-identical to the reference.
FORMAT_HEADER_VALUES = False


def _load_frame(run, camcol, filter, field, raw=False):
    """detecttrails.py:73-117: resolve the path, unpack .bz2 if needed, read image + header.
    Returns (img, header_prefix_of_the_results_line[, big_endian]).  With raw=True and the built-in FITS reader the
    image is the undecoded big-endian payload (uint32 view) and big_endian is True: the device byte-swaps."""
    removefits = False
    fitspath = None
    try:
        origfitspath = files.filename("frame", run=run, camcol=camcol, field=field, filter=filter)
        if not os.path.exists(origfitspath):
            bzpath = origfitspath + ".bz2"
            if not os.path.exists(bzpath):
                errmsg = ("File {0} or its bz2 compressed version not found. Are you sure they exist?")
                raise FileNotFoundError(errmsg.format(origfitspath))
            with open(bzpath, "rb") as compressedfits:
                fitsdata = bz2.decompress(compressedfits.read())
            try:
                fitsdmp = os.environ["FITS_DUMP"]
            except KeyError:
                fitsdmp = os.path.join(os.path.split(__file__)[0], "fits_dump/")
                os.makedirs(fitsdmp, exist_ok=True)
            fitspath = os.path.join(fitsdmp, os.path.split(origfitspath)[-1])
            with open(fitspath, "wb") as decompressed:
                decompressed.write(fitsdata)
            removefits = True
        else:
            fitspath = origfitspath
        big_endian = False
        if raw and hasattr(fitsio, "read_raw_image"):
            try:
                img, h = fitsio.read_raw_image(fitspath)
                big_endian = True
            except ValueError:           # not a plain BITPIX=-32 image: decode on the host
                img = fitsio.read(fitspath)
                h = fitsio.read_header(fitspath)
        else:
            img = fitsio.read(fitspath)
            h = fitsio.read_header(fitspath)
        if not big_endian and img.dtype != _np.float32:
            img = img.astype(_np.float32)
        printit = _results_prefix(run, camcol, filter, field, h)
        return (img, printit, big_endian) if raw else (img, printit)
    finally:
        if removefits:
            os.remove(fitspath)


def _error_text(run, camcol, filter, field, exc):
    """The text detecttrails.py:133-139 appends to errors.txt for one failed frame."""
    return (f"{run} {camcol} {filter} {field}\n" +
            "".join(traceback.format_exception(type(exc), exc, exc.__traceback__, limit=3)) + str(exc) + "\n\n")


def _results_prefix(run, camcol, filter, field, h):
    """The header part of the results line (detecttrails.py:115-117); only the first fragment is an f-string in the
    reference, the seven other brace groups are literal text and stay literal unless FORMAT_HEADER_VALUES is set."""
    if FORMAT_HEADER_VALUES:
        return (f"{run} {camcol} {filter} {field} {h['TAI']} {h['CRPIX1']} "
                f"{h['CRPIX2']} {h['CRVAL1']} {h['CRVAL2']} {h['CD1_1']} "
                f"{h['CD1_2']} {h['CD2_1']} {h['CD2_2']} ")
    return (f"{run} {camcol} {filter} {field} {h['TAI']} {h['CRPIX1']} "
            "{h['CRPIX2']} {h['CRVAL1']} {h['CRVAL2']} {h['CD1_1']} "
            "{h['CD1_2']} {h['CD2_1']} {h['CD2_2']} ")


# FITS frames and photoObj tables are read by the library's host-side ingest (csrc/host_ingest.cuh) when they have the
# plain layout of the SDSS files; LFD_NATIVE_INGEST=0 keeps everything on the Python readers (same results, tested)
NATIVE_INGEST = os.environ.get("LFD_NATIVE_INGEST", "1") != "0"


class _RawHeader:
    """The nine header cards of the results line as returned by lfd_fits_load_frame, parsed on access like a
    FITSHeader (fitsio_lite._parse_value: same value formatting in the results line)."""

    def __init__(self, raw):
        self._raw = raw

    def __getitem__(self, key):
        from .fitsio_lite import _parse_value
        return _parse_value(self._raw[key])


def _load_one(frame, params_removestars, slot=None, want_rects=True):
    """Host side of one frame (runs in a loader thread): FITS image + header line prefix, photoObj catalog -> blot
    rectangles.  With ``slot`` (a row of a handle's pinned staging viewed as uint32) an uncompressed frame of that
    shape is read straight into it as raw big-endian payload.  With ``want_rects=False`` the catalog columns are
    returned in place of the rectangles (the driver resolves a whole batch at once, star_rects_batch).
    Returns ("staged", None, True, rects, printit) | ("ok", pixels, big_endian, rects, printit) | ("err", exc)."""
    run, camcol, filter, field = frame
    try:
        staged = False
        if slot is not None and hasattr(fitsio, "read_raw_image_into"):
            path = files.filename("frame", run=run, camcol=camcol, field=field, filter=filter)
            if os.path.exists(path):
                raw = _lib.fits_load_frame(path, slot) if NATIVE_INGEST else None      # native reader, GIL released
                if raw is not None:
                    h = _RawHeader(raw)
                    printit = _results_prefix(run, camcol, filter, field, h)
                    staged = True
                else:
                    try:
                        h = fitsio.read_raw_image_into(path, slot)
                        printit = _results_prefix(run, camcol, filter, field, h)
                        staged = True
                    except ValueError:
                        staged = False
        if not staged:
            img, printit, big_endian = _load_frame(run, camcol, filter, field, raw=True)
        shape = slot.shape if staged else img.shape
        opath = files.filename("photoObj", run=run, camcol=camcol, field=field)
        rects = _lib.catalog_rects(opath, filter, shape, **dict(params_removestars)) if NATIVE_INGEST else None
        if rects is None:                                    # not the plain table layout, or a value the reference raises on
            cat = read_photoObj_arrays(opath)
            rects = star_rects(cat, filter, shape, **dict(params_removestars)) if want_rects else cat
        if staged:
            return ("staged", None, True, rects, printit)
        return ("ok", img, big_endian, rects, printit)
    except Exception as e:   # noqa: BLE001 - the reference swallows everything per frame
        return ("err", e)


def _resolve_rects(loaded, chunk, shape0, params_removestars):
    """Replace the catalog columns the loaders returned by blot rectangles: one vectorised call for the frames of
    the batch that share the common frame shape, frame by frame for the others."""
    loaded = list(loaded)
    pr = dict(params_removestars)
    def have(it):                                            # rectangles already resolved by the native ingest
        return it[0] != "err" and isinstance(it[3], _np.ndarray)

    group = [j for j, it in enumerate(loaded) if not have(it) and
             (it[0] == "staged" or (it[0] == "ok" and it[1].shape == shape0))]
    if group:
        res = star_rects_batch([loaded[j][3] for j in group], [chunk[j][2] for j in group], shape0, **pr)
        for j, r in zip(group, res):
            loaded[j] = ("err", r) if isinstance(r, BaseException) else loaded[j][:3] + (r,) + loaded[j][4:]
    for j, it in enumerate(loaded):
        if it[0] == "ok" and j not in group and not have(it):
            try:
                loaded[j] = it[:3] + (star_rects(it[3], chunk[j][2], it[1].shape, **pr),) + it[4:]
            except Exception as e:   # noqa: BLE001
                loaded[j] = ("err", e)
    return loaded


_handles = {}
N_RING = 3      # handles per frame shape: one computing, one submitted (crossing PCIe), one being filled by the loaders


def _batch_handles(shape, batch, device, count=N_RING):
    key = (shape, device)
    hs = _handles.get(key)
    if hs is None or hs[0].B < batch or len(hs) < count:
        for h in hs or []:
            h.close()
        hs = [_lib.Handle(shape[0], shape[1], max_batch=batch, device=device) for _ in range(count)]
        _handles[key] = hs
    return hs


def _probe_shape(frames):
    """Shape of the first readable frame (SDSS frames of a run all have one shape); None if none can be read."""
    for (run, camcol, filter, field) in frames[:8]:
        try:
            path = files.filename("frame", run=run, camcol=camcol, field=field, filter=filter)
            if os.path.exists(path):
                h = fitsio.read_header(path)
                return (int(h["NAXIS2"]), int(h["NAXIS1"]))
        except Exception:   # noqa: BLE001
            continue
    return None


def compute_fields(frames, params_bright, params_dim, params_removestars, batch=16, device=0, loaders=None):
    """Run an ordered list of (run, camcol, filter, field) through the GPU in batches.  Returns one record per
    frame, in list order: ("line", results_line) for a detection, ("none", "") for no detection,
    ("err", errors_text) for a failure - exactly what the reference's per-frame loop would append.

    Pipeline (three handles of the common frame shape form a ring): while batch k-1 computes and batch k crosses
    PCIe, loader threads read the FITS payloads of batch k+1 straight into the third handle's pinned staging
    (raw big-endian bytes, `readinto`, no GIL; the first kernel byte-swaps) and filter the catalogs.  Frames that
    are compressed, of another shape or another BITPIX take the decoded-array path through the same ring."""
    from concurrent.futures import ThreadPoolExecutor
    debug = bool(params_bright["debug"] or params_dim["debug"])
    frames = list(frames)
    if debug:
        return _compute_fields_debug(frames, params_bright, params_dim, params_removestars)
    batch = max(int(batch), 1)
    chunks = [frames[i:i + batch] for i in range(0, len(frames), batch)]
    records = [None] * len(frames)
    outcome = {}             # global frame index -> ("ok", detected, result dict) | ("err", exc)
    loaded_all = {}          # global frame index -> loader result
    pending = []             # in-flight device batches: (handle, shape, [global indices])
    shape0 = _probe_shape(frames)
    ring = []
    if shape0 is not None:
        try:
            ring = _batch_handles(shape0, batch, device)
        except Exception:   # noqa: BLE001 - reported per frame below, when the submit fails the same way
            ring = []

    def collect(entry):
        h, shape, gidx = entry
        try:
            res = h.wait()
        except Exception as e:   # noqa: BLE001
            for g in gidx:
                outcome[g] = ("err", e)
            return
        for k, g in enumerate(gidx):
            try:
                r = res[k]
                if r.status & _lib.FRAME_OVERFLOW:
                    raise _lib.LfdError(_lib.LFD_E_CAPACITY, "per-frame work list overflow")
                det, out = (False, None)
                for p in (0, 1):
                    if r.rect_detection[p] >= 0:
                        det, out = result_from_device(r, p, shape)
                        if det:
                            break
                outcome[g] = ("ok", det, out)
            except Exception as e:   # noqa: BLE001
                outcome[g] = ("err", e)

    def release(h):
        for entry in [e for e in pending if e[0] is h]:
            collect(entry)
            pending.remove(entry)

    def submit(h, shape, gidx, rects, big_endian):
        try:
            h.set_params(params_bright, params_dim)
            h.submit(len(gidx), rects, flags=_lib.INPUT_BIGENDIAN if big_endian else 0)
            pending.append((h, shape, gidx))
        except Exception as e:   # noqa: BLE001
            for g in gidx:
                outcome[g] = ("err", e)

    # measured on the 16-core host: 12 threads with the native ingest (3.2 k frames/s; 8: 2.7 k, 16: 2.8 k), 8 with the
    # Python readers, which hold the GIL for longer (2.3 k; more threads only contend for it)
    nload = (loaders or int(os.environ.get("LFD_LOADER_THREADS", 0)) or
             max(2, min(12 if NATIVE_INGEST else 8, (os.cpu_count() or 2))))
    with ThreadPoolExecutor(max_workers=nload) as pool:
        futs = {}

        def prefetch(ci):
            if ci >= len(chunks) or ci in futs:
                return
            h = ring[ci % len(ring)] if ring else None
            if h is not None:
                release(h)                                   # its previous batch (ci - N_RING) is long done
                st = h.host_frames.view(_np.uint32)
                futs[ci] = [pool.submit(_load_one, fr, params_removestars, st[j], False) for j, fr in enumerate(chunks[ci])]
            else:
                futs[ci] = [pool.submit(_load_one, fr, params_removestars, None, False) for fr in chunks[ci]]

        prefetch(0)
        for ci, chunk in enumerate(chunks):
            prefetch(ci + 1)
            loaded = _resolve_rects([f.result() for f in futs.pop(ci)], chunk, shape0, params_removestars)
            base = ci * batch
            for j, item in enumerate(loaded):
                loaded_all[base + j] = item
            h = ring[ci % len(ring)] if ring else None
            staged = [j for j, it in enumerate(loaded) if it[0] == "staged"]
            others = {}
            for j, it in enumerate(loaded):
                if it[0] == "ok":
                    others.setdefault((it[1].shape, it[2]), []).append(j)
            if staged:
                # staged frames sit in their own slots; compact them to the front (a frame that failed or took the
                # other path leaves a hole)
                st = h.host_frames.view(_np.uint32)
                for k, j in enumerate(staged):
                    if k != j:
                        st[k] = st[j]
                submit(h, shape0, [base + j for j in staged], [loaded[j][3] for j in staged], True)
            for (shape, big_endian), idxs in others.items():
                gidx = [base + j for j in idxs]
                try:
                    if h is not None and shape == shape0 and not staged:
                        hh = h
                    else:
                        # rare: another frame shape, or a second group in this chunk -> a handle of its own, synchronously
                        hh = _lib.Handle(shape[0], shape[1], max_batch=len(idxs), device=device)
                    stg = hh.host_frames.view(_np.uint32) if big_endian else hh.host_frames
                    for slot, j in enumerate(idxs):
                        stg[slot] = loaded[j][1]
                    submit(hh, shape, gidx, [loaded[j][3] for j in idxs], big_endian)
                    if hh is not h:
                        release(hh)
                        hh.close()
                except Exception as e:   # noqa: BLE001
                    for g in gidx:
                        outcome[g] = ("err", e)
        for entry in list(pending):
            collect(entry)
    for g, (run, camcol, filter, field) in enumerate(frames):
        item = loaded_all[g]
        exc = item[1] if item[0] == "err" else (outcome[g][1] if outcome[g][0] == "err" else None)
        if exc is not None:
            if debug:
                traceback.print_exception(type(exc), exc, exc.__traceback__, limit=3)
            records[g] = ("err", _error_text(run, camcol, filter, field, exc))
        elif outcome[g][1]:
            res = outcome[g][2]
            records[g] = ("line", item[4] + f"{res['x1']} {res['y1']} {res['x2']} {res['y2']}\n")
        else:
            records[g] = ("none", "")
    return records


def _compute_fields_debug(frames, params_bright, params_dim, params_removestars):
    """debug=True: one frame at a time through the standalone stage functions, which write the reference's debug
    images into $DEBUG_PATH and print the check_theta values (detecttrails.py:119-131 literally)."""
    from .processfield import process_field_bright, process_field_dim
    from .removestars import remove_stars
    records = []
    for (run, camcol, filter, field) in frames:
        try:
            img, printit = _load_frame(run, camcol, filter, field)
            img = remove_stars(_np.ascontiguousarray(img, _np.float32), run, camcol, filter, field, **params_removestars)
            img = _np.ascontiguousarray(img[::-1])                   # cv2.flip(img, 0)
            detection, res = process_field_bright(img, **params_bright)
            if not detection:
                detection, res = process_field_dim(img, **params_dim)
            if detection:
                records.append(("line", printit + f"{res['x1']} {res['y1']} {res['x2']} {res['y2']}\n"))
            else:
                records.append(("none", ""))
        except Exception as e:   # noqa: BLE001
            traceback.print_exception(type(e), e, e.__traceback__, limit=3)
            records.append(("err", _error_text(run, camcol, filter, field, e)))
    return records


def write_records(results, errors, records):
    for kind, text in records:
        if kind == "line":
            results.write(text)
        elif kind == "err":
            errors.write(text)


def _progress_key(frame):
    run, camcol, filter, field = frame
    return f"{run} {camcol} {filter} {field}"


def read_progress(path):
    """Frames already finished by an earlier, interrupted run (one "run camcol filter field status" line each)."""
    done = set()
    if path and os.path.exists(path):
        with open(path) as f:
            for line in f:
                t = line.split()
                if len(t) >= 5:
                    done.add(" ".join(t[:4]))
    return done


def process_fields(results, errors, frames, params_bright, params_dim, params_removestars, batch=16, device=0,
                   distributed=None, compute=None, progress=None):
    """Process an ordered list of (run, camcol, filter, field); results/errors are written in list order,
    exactly the lines the reference's per-frame loop would write.

    With ``torch.distributed`` initialised (one process per GPU, e.g. under torchrun) the list is sharded by
    frame across the ranks - blocks of ``batch`` consecutive frames dealt round-robin, no data-path collective -
    and rank 0 gathers the per-frame records and writes them in the original order (lfd_b200/sharding.py).
    ``distributed=False`` forces the single-process path; ``compute`` replaces the GPU stage (tests)."""
    from . import sharding
    compute = compute or (lambda fr: compute_fields(fr, params_bright, params_dim, params_removestars, batch=batch, device=device))
    frames = list(frames)
    if distributed is None:
        distributed = sharding.is_distributed()
    # resumable runs: `progress` is a text file with one line per finished frame; frames listed
    # there are skipped, and the file is extended chunk by chunk, after the chunk's results/errors lines are flushed
    if progress:
        done = read_progress(progress)
        frames = [fr for fr in frames if _progress_key(fr) not in done]
    writer = (not distributed) or sharding.rank() == 0
    # frames per compute call: the ring drains and refills at every call boundary (one un-overlapped batch load plus
    # one un-overlapped batch of compute, ~15 ms), so calls are long; a progress file is extended once per call
    chunk = max(batch, 1) * (sharding.world_size() if distributed else 1) * 32
    for i0 in range(0, len(frames), chunk):
        part = frames[i0:i0 + chunk]
        recs = sharding.run_sharded(part, compute, block=batch) if distributed else compute(part)
        if not writer:
            continue
        write_records(results, errors, recs)
        if progress:
            results.flush(); errors.flush()
            with open(progress, "a") as pf:
                for fr, (kind, _text) in zip(part, recs):
                    pf.write(_progress_key(fr) + " " + kind + "\n")


def process_field(results, errors, run, camcol, filter, field, params_bright, params_dim, params_removestars):
    """detecttrails.py:30-143 for one frame."""
    process_fields(results, errors, [(run, camcol, filter, field)], params_bright, params_dim,
                   params_removestars, batch=1)


class DetectTrails:
    """Convenience class that processes targeted SDSS frames (detecttrails.py:146-407).

    Extra, optional kwargs that the reference does not have: ``batch`` (frames per GPU batch, default 16),
    ``device`` (CUDA device index, default: LOCAL_RANK under torchrun, else 0) and ``resume`` (path of a progress
    file, or True for ``<savepath>/progress.txt``: frames listed there are skipped and finished frames are appended,
    so an interrupted run can be restarted without duplicating lines - the reference only appends, detecttrails.py:349).
    Under torchrun (torch.distributed initialised) the frame list is sharded over the ranks and rank 0 writes."""

    def __init__(self, **kwargs):
        savepth = (kwargs["savepath"] if "savepath" in kwargs else ".")
        self.kwargs = kwargs
        self.params_bright, self.params_dim, self.params_removestars = default_params()
        self.batch = int(kwargs.get("batch", 16))
        self.device = int(kwargs.get("device", os.environ.get("LOCAL_RANK", 0)))
        resume = kwargs.get("resume", None)
        self.progress = (os.path.join(savepth, "progress.txt") if resume is True else resume) or None

        if "results" in kwargs:
            self.results = kwargs["results"]
        else:
            self.results = os.path.join(savepth, "results.txt")
        if "errors" in kwargs:
            self.errors = kwargs["errors"]
        else:
            self.errors = os.path.join(savepth, "errors.txt")

        # reference behaviour, kept on purpose: all three land in params_bright (detecttrails.py:251-256)
        if "params_bright" in kwargs:
            self.params_bright = kwargs["params_bright"]
        if "params_dim" in kwargs:
            self.params_bright = kwargs["params_dim"]
        if "params_removestars" in kwargs:
            self.params_bright = kwargs["params_removestars"]

        if "debug" in kwargs:
            self.debug = kwargs.pop("debug")
            self.params_bright["debug"] = self.debug
            self.params_dim["debug"] = self.debug
            self.params_removestars["debug"] = self.debug
        if any([self.params_removestars["debug"], self.params_bright["debug"], self.params_dim["debug"]]):
            setup_debug()

        self._load()

    def _runInfo(self):
        rl = files.runlist()
        w, = _np.where(rl["run"] == self._run)
        if len(w) == 0:
            raise ValueError("Run %s not found in runList.par" % self._run)
        return rl[w]["startfield"][0], rl[w]["endfield"][0]

    def _getRuns(self):
        rl = files.runlist()
        runs = rl["run"]
        if runs is None:
            raise ValueError("Unable to retrieve runs. Retrieved NoneType.")
        return runs

    def _load(self):
        """kwargs -> self._pick (detecttrails.py:290-342)."""
        self._run, self._camcol, self._field = int(0), int(0), int(0)
        self._filter, self._pick = str(0), str(0)
        kwargs = self.kwargs
        if "run" in kwargs:
            self._run = kwargs["run"]
            self._pick = "run"
        if "camcol" in kwargs:
            if kwargs["camcol"] not in (1, 2, 3, 4, 5, 6):
                raise ValueError("Nonexisting camcol")
            self._camcol = kwargs["camcol"]
            self._pick = "run-camcol"
        if "field" in kwargs or "frame" in kwargs:
            if self._camcol == 0:
                raise ValueError("send camcol= ")
            self._field = kwargs["field"] if "field" in kwargs else kwargs["frame"]
        if "filter" in kwargs:
            if kwargs["filter"] not in ("u", "g", "r", "i", "z"):
                raise ValueError("Nonexistting filter")
            self._filter = kwargs["filter"]
            if self._camcol != 0:
                self._pick = "camcol-filter"
            if self._run != 0:
                self._pick = "run-filter"
            if (self._camcol != 0) and (self._run != 0):
                self._pick = "run-camcol-filter"
        if "filter" not in kwargs:
            if (self._field != 0) and (self._camcol != 0):
                self._pick = "camcol-frame"
        if (self._field != 0) and (self._camcol != 0) and (self._filter != "0"):
            self._pick = "field"

    def frame_list(self):
        """The (run, camcol, filter, field) sequence ``process`` visits, in the reference's order."""
        out = []
        filters = ("u", "g", "r", "i", "z")
        camcols = (1, 2, 3, 4, 5, 6)
        if self._pick == "camcol-filter":
            for _run in self._getRuns():
                self._run = _run
                startfield, endfield = self._runInfo()
                out += [(_run, self._camcol, self._filter, f) for f in range(startfield, endfield, 1)]
            self._run = 0
        if self._pick == "run":
            startfield, endfield = self._runInfo()
            for c in camcols:
                for flt in filters:
                    out += [(self._run, c, flt, f) for f in range(startfield, endfield, 1)]
        if self._pick == "run-filter":
            startfield, endfield = self._runInfo()
            for c in camcols:
                out += [(self._run, c, self._filter, f) for f in range(startfield, endfield, 1)]
        if self._pick == "run-camcol":
            startfield, endfield = self._runInfo()
            for flt in filters:
                out += [(self._run, self._camcol, flt, f) for f in range(startfield, endfield, 50)]
        if self._pick == "run-camcol-filter":
            startfield, endfield = self._runInfo()
            out += [(self._run, self._camcol, self._filter, f) for f in range(startfield, endfield, 1)]
        if self._pick == "camcol-frame":
            out += [(self._run, self._camcol, flt, self._field) for flt in filters]
        if self._pick == "field":
            out.append((self._run, self._camcol, self._filter, self._field))
        return out

    def process(self):
        """Run the selected frames; results/errors files are opened in append mode like the original."""
        with open(self.results, "a") as results, open(self.errors, "a") as errors:
            process_fields(results, errors, self.frame_list(), self.params_bright, self.params_dim,
                           self.params_removestars, batch=self.batch, device=self.device, progress=self.progress)
