"""Minimal reader for R's `save()` files (bzip2/gzip/xz + RDX2, XDR serialisation v2).  ORACLE/FIXTURE
TOOLING ONLY: used by tests/golden/make_golden.py in the build container to turn the reference's
bundled dataset (/root/reference/data/counts.rda, documented in man/counts.Rd) into small committed
fixtures.  Nothing at run time on the GPU box reads the reference tree.

Supports exactly the SEXP types a tibble of character/double/integer/logical/factor columns needs.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INT = -2147483648


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.i = 0
        self.refs: list = []

    def int(self) -> int:
        v = struct.unpack_from(">i", self.b, self.i)[0]
        self.i += 4
        return v

    def item(self):
        flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:                              # NILVALUE_SXP
            return None
        if t == 255:                              # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == 1:                                # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                                # LISTSXP (pairlist) -> list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                out.append((tag, self.item()))
                flags = self.int()
                t = flags & 0xFF
                if t == 254:
                    return out
                if t != 2:
                    raise ValueError(f"unexpected pairlist tail type {t}")
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
        if t == 9:                                # CHARSXP
            n = self.int()
            if n == -1:
                return None
            s = self.b[self.i:self.i + n].decode("utf-8", "replace")
            self.i += n
            return s
        if t in (10, 13):                         # LGLSXP, INTSXP
            n = self.int()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.i).astype(np.int32)
            self.i += 4 * n
        elif t == 14:                             # REALSXP
            n = self.int()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.i).astype(np.float64)
            self.i += 8 * n
        elif t == 16:                             # STRSXP
            n = self.int()
            v = [self.item() for _ in range(n)]
        elif t == 19:                             # VECSXP
            n = self.int()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.i}")
        attrs = dict(self.item()) if has_attr else {}
        return _Vec(v, attrs) if attrs else v


class _Vec:
    def __init__(self, value, attrs):
        self.value = value
        self.attrs = attrs


def _decompress(raw: bytes) -> bytes:
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def load_rda(path: str) -> dict:
    """Returns {object name: value}; a data.frame/tibble becomes {column: numpy array or list[str]}."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] != b"RDX2\n" or buf[5:7] != b"X\n":
        raise ValueError("not an RDX2/XDR file")
    r = _Reader(buf)
    r.i = 7
    version, _, _ = r.int(), r.int(), r.int()
    if version != 2:
        raise ValueError(f"serialisation version {version} not supported")
    top = r.item()
    out = {}
    for name, obj in top:
        out[name] = _simplify(obj)
    return out


def _simplify(obj):
    if isinstance(obj, _Vec):
        names = obj.attrs.get("names")
        cls = obj.attrs.get("class")
        if isinstance(obj.value, list) and names is not None and cls is not None and "data.frame" in cls:
            return {n: _simplify(c) for n, c in zip(names, obj.value)}
        if "levels" in obj.attrs:                 # factor -> list of strings
            lv = obj.attrs["levels"]
            return [lv[i - 1] if i != NA_INT else None for i in obj.value]
        return obj.value
    return obj
