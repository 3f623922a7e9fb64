"""Minimal FITS reader/writer used when ``fitsio`` is not installed.

The reference reads frames with ``fitsio.read(path)`` / ``fitsio.read_header(path)``
(/root/reference/lfd/detecttrails/detecttrails.py:113-114) and photoObj tables with
``fitsio.read(path, header="True")`` (/root/reference/lfd/detecttrails/removestars.py:96).
This module covers exactly those three calls for the two file kinds on the path:

* primary HDU image, ``BITPIX=-32`` (also 8/16/32/-64), returned native-endian with shape
  ``(NAXIS2, NAXIS1)``;
* empty primary + one ``BINTABLE`` extension whose columns have ``TFORMn`` in
  ``{rJ, rE, rD, rI, rB, rK, rA}`` (``r`` an optional repeat count).

``read_raw_image`` additionally returns the *undecoded* big-endian payload so the frame
driver can upload it and byte-swap on the device.
"""
import io
import os
import re

import numpy as np

__all__ = ["read", "read_header", "read_columns", "write_image", "write_bintable", "read_raw_image", "read_raw_image_into",
           "FITSHeader"]

BLOCK = 2880
CARD = 80


class FITSHeader(dict):
    """Header as a dict keyed by upper-case keyword (enough for ``h['TAI']`` style access).

    Values read from a file are kept as the raw card text and parsed on first access: a frame header has ~40 cards
    of which the driver reads a dozen, and parsing is per-frame interpreter time the loader threads serialise on."""

    __slots__ = ("_raw",)

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._raw = set()

    def _set_raw(self, key, text):
        dict.__setitem__(self, key, text)
        self._raw.add(key)

    def __setitem__(self, key, value):
        self._raw.discard(key)
        dict.__setitem__(self, key, value)

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if key in self._raw:
            v = _parse_value(v)
            self._raw.discard(key)
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def _all(self):
        for k in list(self._raw):
            self[k]
        return self

    def items(self):
        return dict.items(self._all())

    def copy(self):
        h = FITSHeader()
        dict.update(h, self._all())
        return h

    def values(self):
        return dict.values(self._all())

    def __eq__(self, other):
        if isinstance(other, FITSHeader):
            other._all()
        return dict.__eq__(self._all(), other)

    __hash__ = None

    def __repr__(self):
        return dict.__repr__(self._all())


def _parse_value(raw):
    raw = raw.strip()
    if not raw:
        return None
    if raw[0] == "'":
        # string value: up to the closing quote ('' is an escaped quote)
        out, i = [], 1
        while i < len(raw):
            if raw[i] == "'":
                if i + 1 < len(raw) and raw[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(raw[i])
            i += 1
        return "".join(out).rstrip()
    val = raw.split("/")[0].strip()
    if val == "T":
        return True
    if val == "F":
        return False
    try:
        return int(val)
    except ValueError:
        pass
    try:
        return float(val.replace("D", "E").replace("d", "e"))
    except ValueError:
        return val


def _read_header_at(f):
    """Read header cards from the current position. Returns (FITSHeader, bytes_consumed)."""
    hdr = FITSHeader()
    nread = 0
    done = False
    while not done:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            raise EOFError("truncated FITS header")
        nread += BLOCK
        text = block.decode("ascii", errors="replace")
        for i in range(0, BLOCK, CARD):
            key = text[i:i + 8].strip()
            if key == "END":
                done = True
                break
            if not key or key == "COMMENT" or key == "HISTORY":
                continue
            if text[i + 8:i + 10] == "= ":
                hdr._set_raw(key, text[i + 10:i + CARD])
    return hdr, nread


_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _data_nbytes(h):
    naxis = h.get("NAXIS", 0)
    if naxis == 0:
        return 0
    n = 1
    for i in range(1, naxis + 1):
        n *= h["NAXIS%d" % i]
    n = abs(h["BITPIX"]) // 8 * h.get("GCOUNT", 1) * (h.get("PCOUNT", 0) + n)
    return n


def _pad(n):
    return (n + BLOCK - 1) // BLOCK * BLOCK


_TFORM_RE = re.compile(r"^(\d*)([LXBIJKAEDCMPQ])")
_TFORM_DTYPE = {"B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8", "L": "u1"}


_dtype_cache = {}


def _table_dtype(h):
    raw = dict.__getitem__              # the unparsed card text is as good a cache key as the parsed value
    key = tuple((raw(h, "TTYPE%d" % i), raw(h, "TFORM%d" % i)) for i in range(1, h["TFIELDS"] + 1))
    dt = _dtype_cache.get(key)
    if dt is None:
        dt = _dtype_cache[key] = _build_table_dtype(h)
    return dt


def _build_table_dtype(h):
    fields = []
    for i in range(1, h["TFIELDS"] + 1):
        name = h["TTYPE%d" % i]
        m = _TFORM_RE.match(h["TFORM%d" % i].strip())
        if m is None:
            raise ValueError("unsupported TFORM %r" % h["TFORM%d" % i])
        rep = int(m.group(1)) if m.group(1) else 1
        code = m.group(2)
        if code == "A":
            fields.append((name, "S%d" % rep))
        elif code in _TFORM_DTYPE:
            if rep == 1:
                fields.append((name, _TFORM_DTYPE[code]))
            else:
                fields.append((name, _TFORM_DTYPE[code], (rep,)))
        else:
            raise ValueError("unsupported TFORM code %r" % code)
    return np.dtype(fields)


def _native(arr):
    """Big-endian (on disk) -> native-endian copy, like fitsio returns."""
    if arr.dtype.fields is None:
        return arr.astype(arr.dtype.newbyteorder("="))
    return arr.astype(arr.dtype.newbyteorder("="))


def _iter_hdus_f(f):
    """(header, data offset, data bytes) of every HDU of the open file ``f``; the caller may seek between items."""
    try:
        size = os.fstat(f.fileno()).st_size
    except (AttributeError, OSError, io.UnsupportedOperation):      # an in-memory file (io.BytesIO)
        size = f.seek(0, 2)
    pos = 0
    while pos < size:
        f.seek(pos)
        try:
            h, nh = _read_header_at(f)
        except EOFError:
            return
        nbytes = _data_nbytes(h)
        pos += nh
        yield h, pos, nbytes
        pos += _pad(nbytes)


def _iter_hdus(path):
    with open(path, "rb") as f:
        yield from _iter_hdus_f(f)


def _read_hdu(path, h, pos, nbytes, f=None):
    if nbytes == 0:
        return None
    if f is None:
        with open(path, "rb") as f2:
            f2.seek(pos)
            buf = f2.read(nbytes)
    else:
        f.seek(pos)
        buf = f.read(nbytes)
    if h.get("XTENSION", "").strip() == "BINTABLE":
        dt = _table_dtype(h)
        if dt.itemsize != h["NAXIS1"]:
            raise ValueError("row size mismatch: dtype %d vs NAXIS1 %d" % (dt.itemsize, h["NAXIS1"]))
        arr = np.frombuffer(buf, dtype=dt, count=h["NAXIS2"])
        return _native(arr).view(np.recarray)
    dt = np.dtype(_BITPIX_DTYPE[h["BITPIX"]])
    shape = tuple(h["NAXIS%d" % i] for i in range(h["NAXIS"], 0, -1))
    arr = np.frombuffer(buf, dtype=dt).reshape(shape)
    arr = _native(arr)
    if "BSCALE" in h or "BZERO" in h:
        bs, bz = h.get("BSCALE", 1), h.get("BZERO", 0)
        if bs != 1 or bz != 0:
            arr = arr * bs + bz
    return arr


def read(path, ext=None, header=False):
    """``fitsio.read`` stand-in: first HDU with data unless ``ext`` is given.

    ``header`` truthy (the reference passes the *string* ``"True"``) returns ``(data, header)``.
    """
    with open(path, "rb") as f:
        found = None
        for i, (h, pos, nbytes) in enumerate(_iter_hdus_f(f)):
            if (ext is None and nbytes > 0) or (ext is not None and i == ext):
                found = (h, pos, nbytes)
                break
        if found is None:
            if ext is None:
                raise OSError("No extensions have data")
            raise IndexError("list index out of range")
        h, pos, nbytes = found
        data = _read_hdu(path, h, pos, nbytes, f)
    if header:
        return data, h
    return data


def read_columns(path, names):
    """Selected columns of the first binary-table HDU as native-endian arrays ({name: array}) - what the catalog
    filter needs (removestars.py:97-104) without converting the other columns of a photoObj table (hundreds in the
    real files).  Raises KeyError for a missing column like the recarray access would."""
    with open(path, "rb") as f:
        for h, pos, nbytes in _iter_hdus_f(f):
            if nbytes == 0:
                continue
            if h.get("XTENSION", "").strip() != "BINTABLE":
                raise ValueError("first HDU with data is not a binary table")
            dt = _table_dtype(h)
            if dt.itemsize != h["NAXIS1"]:
                raise ValueError("row size mismatch: dtype %d vs NAXIS1 %d" % (dt.itemsize, h["NAXIS1"]))
            f.seek(pos)
            arr = np.frombuffer(f.read(nbytes), dtype=dt, count=h["NAXIS2"])
            out = {}
            for k in names:
                if k not in dt.fields:
                    raise KeyError(k)
                col = arr[k]
                out[k] = col.astype(col.dtype.newbyteorder("="))
            return out
    raise OSError("No extensions have data")


def read_header(path, ext=0):
    """``fitsio.read_header`` stand-in."""
    for i, (h, pos, nbytes) in enumerate(_iter_hdus(path)):
        if i == ext:
            return h
    raise OSError("extension %d not found in %s" % (ext, path))


def _require_unscaled(h):
    """The raw payload is only the image when no scaling applies: ``read`` multiplies by BSCALE and adds BZERO, the
    raw upload path cannot, so anything but the trivial 1 / 0 takes the decoded path (ValueError -> caller falls back)."""
    if h.get("BSCALE", 1) != 1 or h.get("BZERO", 0) != 0:
        raise ValueError("raw upload path needs an unscaled image (BSCALE=1, BZERO=0)")


def read_raw_image(path):
    """Return ``(payload_bytes_view, header)`` for the primary float32 image without byte-swapping.

    The payload is the on-disk big-endian ``>f4`` data of shape (NAXIS2, NAXIS1) as a uint32 array;
    the device decodes it (see ``lfd_submit(..., LFD_INPUT_BIGENDIAN)`` in include/lfd_b200.h).
    """
    for h, pos, nbytes in _iter_hdus(path):
        if nbytes == 0:
            continue
        if h["BITPIX"] != -32 or h["NAXIS"] != 2:
            raise ValueError("raw upload path needs a 2-D BITPIX=-32 image")
        _require_unscaled(h)
        nbytes = min(nbytes, 4 * h["NAXIS1"] * h["NAXIS2"])
        with open(path, "rb") as f:
            f.seek(pos)
            buf = f.read(nbytes)
        arr = np.frombuffer(buf, dtype=np.uint32).reshape(h["NAXIS2"], h["NAXIS1"])
        return arr, h
    raise OSError("no image HDU in %s" % path)


def read_image_bytes(buf):
    """The primary image of a FITS file held in memory (``bytes``, e.g. a decompressed ``.fits.bz2``): returns
    ``(pixels, header, big_endian)`` - the undecoded big-endian payload as a uint32 array when the image is a plain
    unscaled BITPIX=-32 one (big_endian True: the device byte-swaps), else the decoded native array like ``read``."""
    f = io.BytesIO(buf)
    for h, pos, nbytes in _iter_hdus_f(f):
        if nbytes == 0:
            continue
        try:
            if h["BITPIX"] != -32 or h["NAXIS"] != 2:
                raise ValueError("not a 2-D BITPIX=-32 image")
            _require_unscaled(h)
            n = 4 * h["NAXIS1"] * h["NAXIS2"]
            if pos + n > len(buf):
                raise OSError("truncated FITS data")
            arr = np.frombuffer(buf, dtype=np.uint32, count=n // 4, offset=pos).reshape(h["NAXIS2"], h["NAXIS1"])
            return arr, h, True
        except ValueError:
            return _read_hdu(None, h, pos, nbytes, f), h, False
    raise OSError("No extensions have data")


def read_raw_image_into(path, dest):
    """Like read_raw_image, but the payload is read straight into ``dest`` (a writable C-contiguous uint32 array of
    shape (NAXIS2, NAXIS1), e.g. a slot of the library's pinned staging) with ``readinto`` - no intermediate copy and
    no GIL while the bytes move.  Returns the header; raises ValueError if the image does not fit ``dest``."""
    with open(path, "rb", buffering=0) as f:
        for h, pos, nbytes in _iter_hdus_f(f):
            if nbytes == 0:
                continue
            if h["BITPIX"] != -32 or h["NAXIS"] != 2:
                raise ValueError("raw upload path needs a 2-D BITPIX=-32 image")
            _require_unscaled(h)
            nbytes = min(nbytes, 4 * h["NAXIS1"] * h["NAXIS2"])
            if (h["NAXIS2"], h["NAXIS1"]) != tuple(dest.shape) or dest.dtype.itemsize != 4:
                raise ValueError("image shape %s does not match the staging slot %s" % ((h["NAXIS2"], h["NAXIS1"]), dest.shape))
            f.seek(pos)
            mv = memoryview(dest).cast("B")
            got = 0
            while got < nbytes:
                k = f.readinto(mv[got:nbytes])
                if not k:
                    raise OSError("short read in %s" % path)
                got += k
            return h
    raise OSError("no image HDU in %s" % path)


# ----------------------------------------------------------------------------------------------
# writers (used by the synthetic SDSS tree generator and the tests)
# ----------------------------------------------------------------------------------------------
def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, (float, np.floating)):
        s = repr(float(value)).upper()
        if "E" not in s and "." not in s and "N" not in s:
            s += "."
        v = "%20s" % s
    else:
        s = "'%-8s'" % str(value).replace("'", "''")
        v = "%-20s" % s
    card = "%-8s= %s" % (key, v)
    if comment:
        card += " / " + comment
    return card[:CARD].ljust(CARD)


def _header_bytes(cards):
    txt = "".join(cards) + "END".ljust(CARD)
    txt = txt.ljust(_pad(len(txt)))
    return txt.encode("ascii")


def write_image(path, img, header=None):
    """Write a 2-D image as a primary HDU (BITPIX from dtype), big-endian, FITS-padded."""
    img = np.asarray(img)
    bitpix = {"u1": 8, "i2": 16, "i4": 32, "f4": -32, "f8": -64}[img.dtype.str[1:]]
    cards = [_card("SIMPLE", True), _card("BITPIX", bitpix), _card("NAXIS", 2),
             _card("NAXIS1", img.shape[1]), _card("NAXIS2", img.shape[0])]
    for k, v in (header or {}).items():
        cards.append(_card(k, v))
    data = img.astype(img.dtype.newbyteorder(">")).tobytes()
    with open(path, "wb") as f:
        f.write(_header_bytes(cards))
        f.write(data)
        f.write(b"\0" * (_pad(len(data)) - len(data)))


def write_bintable(path, columns, header=None):
    """Write an empty primary HDU plus one BINTABLE extension.

    ``columns`` is an ordered mapping name -> ndarray with first axis = rows; 2-D arrays become
    vector columns (``5E`` etc).
    """
    names = list(columns)
    nrows = len(columns[names[0]])
    fields, tforms = [], []
    for n in names:
        a = np.asarray(columns[n])
        code = {"i2": "I", "i4": "J", "i8": "K", "f4": "E", "f8": "D", "u1": "B"}[a.dtype.str[1:]]
        rep = 1 if a.ndim == 1 else a.shape[1]
        tforms.append("%d%s" % (rep, code))
        be = a.dtype.newbyteorder(">")
        fields.append((n, be) if a.ndim == 1 else (n, be, (rep,)))
    dt = np.dtype(fields)
    rec = np.zeros(nrows, dtype=dt)
    for n in names:
        rec[n] = columns[n]
    primary = [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)]
    ext = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2),
           _card("NAXIS1", dt.itemsize), _card("NAXIS2", nrows), _card("PCOUNT", 0),
           _card("GCOUNT", 1), _card("TFIELDS", len(names))]
    for i, (n, tf) in enumerate(zip(names, tforms), 1):
        ext.append(_card("TTYPE%d" % i, n))
        ext.append(_card("TFORM%d" % i, tf))
    for k, v in (header or {}).items():
        ext.append(_card(k, v))
    data = rec.tobytes()
    with open(path, "wb") as f:
        f.write(_header_bytes(primary))
        f.write(_header_bytes(ext))
        f.write(data)
        f.write(b"\0" * (_pad(len(data)) - len(data)))
