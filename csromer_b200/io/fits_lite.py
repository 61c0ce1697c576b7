"""Minimal FITS image codec (primary HDU only) in NumPy -- astropy is not available in this image.

Implements exactly what the reference's ``io/io_functions.py`` needs from ``astropy.io.fits``: open a file, take
``hdu[0].header`` (a dict-like with comments) and ``hdu[0].data`` (optionally memory-mapped), and write an array with a
header.  FITS standard 4.0 subset: 2880-byte blocks, 80-character cards, ``BITPIX`` in {8, 16, 32, 64, -32, -64},
big-endian data, ``BSCALE``/``BZERO`` applied on read when present, no extensions, no CONTINUE cards.
"""
import numpy as np

BLOCK = 2880
_BITPIX = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_DTYPE_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


class Header(dict):
    """keyword -> value; ``header[key] = (value, comment)`` keeps the comment like astropy does"""

    def __init__(self, *args, **kwargs):
        self.comments = {}
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, tuple) and len(value) == 2:
            self.comments[key] = value[1]
            value = value[0]
        super().__setitem__(key, value)

    def copy(self):
        h = Header(self)
        h.comments = dict(self.comments)
        return h


def _parse_value(text):
    text = text.strip()
    if not text:
        return None
    if text.startswith("'"):
        end = text.find("'", 1)
        while end != -1 and text[end:end + 2] == "''":
            end = text.find("'", end + 2)
        return text[1:end].replace("''", "'").rstrip()
    text = text.split("/")[0].strip()
    if text in ("T", "F"):
        return text == "T"
    try:
        return int(text)
    except ValueError:
        return float(text.replace("D", "E"))


def read_header(f):
    header = Header()
    nbytes = 0
    done = False
    while not done:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            raise ValueError("truncated FITS header")
        nbytes += BLOCK
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if not key or key in ("COMMENT", "HISTORY") or card[8:10] != "= ":
                continue
            body = card[10:]
            header[key] = _parse_value(body)
            if "/" in body and not body.strip().startswith("'"):
                header.comments[key] = body.split("/", 1)[1].strip()
    return header, nbytes


def open_image(path, memmap=True):
    """-> (header, data) of the primary HDU; data shape is (NAXISn, ..., NAXIS1) like astropy"""
    with open(path, "rb") as f:
        header, offset = read_header(f)
    if not header.get("SIMPLE", False):
        raise ValueError("%s is not a simple FITS file" % path)
    naxis = int(header.get("NAXIS", 0))
    shape = tuple(int(header["NAXIS%d" % i]) for i in range(naxis, 0, -1))
    if naxis == 0:
        return header, None
    dtype = np.dtype(_BITPIX[int(header["BITPIX"])])
    if memmap:
        data = np.memmap(path, dtype=dtype, mode="r", offset=offset, shape=shape)
    else:
        data = np.fromfile(path, dtype=dtype, count=int(np.prod(shape)), offset=offset).reshape(shape)
    bscale, bzero = header.get("BSCALE", 1.0), header.get("BZERO", 0.0)
    if bscale != 1.0 or bzero != 0.0:
        data = data.astype(np.float64) * bscale + bzero
    return header, data


def _card(key, value, comment=None):
    if isinstance(value, bool):
        text = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        text = "%20d" % int(value)
    elif isinstance(value, (float, np.floating)):
        text = "%20s" % ("%.15G" % float(value))
        if "." not in text and "E" not in text and "N" not in text and "I" not in text:
            text = "%20s" % ("%.1f" % float(value))
    else:
        text = "'%-8s'" % str(value).replace("'", "''")
        text = "%-20s" % text
    card = "%-8s= %s" % (key[:8], text)
    if comment:
        card += " / " + str(comment)
    return ("%-80s" % card)[:80]


def write_image(path, data, header=None, overwrite=True):
    import os
    if os.path.exists(path) and not overwrite:
        raise OSError("%s exists" % path)
    data = np.asarray(data)
    kind = data.dtype.newbyteorder("=").str[1:]
    if kind not in _DTYPE_BITPIX:
        data = data.astype(np.float32 if data.dtype.kind == "f" else np.float64)
        kind = data.dtype.str[1:]
    structural = {"SIMPLE", "BITPIX", "NAXIS", "EXTEND", "BSCALE", "BZERO"} | {"NAXIS%d" % i for i in range(1, 1000)}
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", _DTYPE_BITPIX[kind]),
             _card("NAXIS", data.ndim)]
    for i, ln in enumerate(reversed(data.shape), start=1):
        comment = header.comments.get("NAXIS%d" % i) if isinstance(header, Header) else None
        cards.append(_card("NAXIS%d" % i, ln, comment))
    for key, value in (header or {}).items():
        if key in structural or value is None:
            continue
        cards.append(_card(key, value, header.comments.get(key) if isinstance(header, Header) else None))
    cards.append("%-80s" % "END")
    text = "".join(cards)
    text += " " * (-len(text) % BLOCK)
    with open(path, "wb") as f:
        f.write(text.encode("ascii"))
        f.write(np.ascontiguousarray(data, dtype=np.dtype(">" + kind)).tobytes())
        f.write(b"\0" * (-(data.size * data.dtype.itemsize) % BLOCK))
