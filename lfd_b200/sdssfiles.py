"""Minimal stand-in for the two ``lfd.detecttrails.sdss.files`` entry points the hot path calls:
``files.filename('frame' | 'photoObj', ...)`` (/root/reference/lfd/detecttrails/detecttrails.py:73,
removestars.py:209; patterns from sdss/share/sdssFileTypes.par:43,75) and ``files.runlist()``
(detecttrails.py:274,284; sdss/files.py:607-671).  Path templating and the runList.par reader are
host-side metadata work and out of scope for acceleration; when the
reference's own ``sdss`` package is importable it can be used instead."""
import os

import numpy as np

_PATTERNS = {
    "frame": ("$BOSS_PHOTOOBJ/frames/$RERUN/$RUNNUM/$COL", "frame-$FILTER-$RUNSTR-$COL-$FIELDSTR.fits"),
    "photoobj": ("$BOSS_PHOTOOBJ/$RERUN/$RUNNUM/$COL", "photoObj-$RUNSTR-$COL-$FIELDSTR.fits"),
    "runlist": ("$PHOTO_REDUX", "runList.par"),
}
_FILTERS = "ugriz"
_runlist_cache = {}
_rerun_cache = {}


def runlist(reload=False):
    """Structured array with the RUNDATA columns of $PHOTO_REDUX/runList.par (cached per path)."""
    path = os.path.join(os.environ["PHOTO_REDUX"], "runList.par")
    if not reload and path in _runlist_cache:
        return _runlist_cache[path]
    rows = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) >= 8 and t[0].upper() == "RUNDATA":
                rows.append((int(t[1]), t[2].encode(), int(t[3]), int(t[4]), int(t[5]), int(t[6]), int(t[7]),
                             (t[8] if len(t) > 8 else "").encode(), (t[9] if len(t) > 9 else "").encode()))
    dt = np.dtype([("run", "i4"), ("rerun", "S8"), ("exist", "i4"), ("done", "i4"), ("calib", "i4"),
                   ("startfield", "i4"), ("endfield", "i4"), ("machine", "S32"), ("disk", "S64")])
    rl = np.array(rows, dtype=dt)
    # the reference drops the duplicate bad entry for run 5194 (sdss/files.py:668-670)
    rl = rl[(rl["run"] != 5194) | (rl["rerun"] == b"301")]
    _runlist_cache[path] = rl
    _rerun_cache.clear()
    return rl


def find_rerun(run):
    key = (os.environ.get("PHOTO_REDUX"), run)
    r = _rerun_cache.get(key)
    if r is None:
        r = _rerun_cache[key] = _find_rerun(run)
    return r


def _find_rerun(run):
    rl = runlist()
    w, = np.where(rl["run"] == run)
    if w.size == 0:
        raise ValueError("Run %s not found in runList.par" % run)
    return rl["rerun"][w[0]].decode()


def filename(ftype, run=None, camcol=None, field=None, **keys):
    try:
        d, n = _PATTERNS[ftype.lower()]
    except KeyError:
        raise ValueError("File type '%s' is unknown" % ftype)
    path = os.path.join(d, n)
    if "$RERUN" in path:
        rerun = keys.get("rerun")
        if rerun is None:
            rerun = find_rerun(run)
        path = path.replace("$RERUN", str(rerun))
    if "$FILTER" in path:
        flt = keys.get("filter")
        if flt is None:
            raise ValueError("filter keyword must be sent for file type '%s'" % ftype)
        if not isinstance(flt, str):
            flt = _FILTERS[int(flt)]
        path = path.replace("$FILTER", flt)
    if run is not None:
        path = path.replace("$RUNNUM", str(int(run))).replace("$RUNSTR", "%06d" % int(run))
    if camcol is not None:
        path = path.replace("$COL", str(int(camcol)))
    if field is not None:
        path = path.replace("$FIELDSTR", "%04d" % int(field))
    return os.path.expandvars(path)
