"""Rutherford-Boeing files (assembled real / pattern, symmetric or not): the matrix input of the
reference's drivers (drivers/spldlt_test.F90 reads `matrix.rb` through SPRAL's rb_read,
spral/src/rutherford_boeing.f90).  Harness-level code: the numeric path
never sees a file, it gets ptr / row / val arrays.

Layout (rb_peek_unit :161-171, read_data_real :915-935, rb_write :716-728):
  line 1  title (a72) key (a8)
  line 2  total, ptr, row, val line counts
  line 3  type code (a3; e.g. rsa = real symmetric assembled), m, n, nnz, 0
  line 4  Fortran formats of the three data blocks (2a16, a20)
  then the column pointers, row indices and values in those fixed-width formats.
"""
from __future__ import annotations

import math
import re

import numpy as np

_FMT = re.compile(r"(\d+)\s*([IiEeDdFfGg])\s*(\d+)")


def _parse_format(fmt: str):
    """(fields per line, field width) of a Fortran format such as (10I8) or (1P,3E25.16)."""
    m = _FMT.search(fmt)
    if not m:
        raise ValueError(f"unsupported Rutherford-Boeing format {fmt!r}")
    return int(m.group(1)), int(m.group(3))


def _read_block(lines, pos, count, fmt, conv):
    per_line, width = _parse_format(fmt)
    out = []
    while len(out) < count:
        line = lines[pos].rstrip("\n")
        pos += 1
        for k in range(per_line):
            if len(out) == count:
                break
            field = line[k * width:(k + 1) * width]
            if not field.strip():
                break
            out.append(conv(field))
    return out, pos


def _to_float(s: str) -> float:
    s = s.strip().replace("D", "E").replace("d", "e")
    if re.match(r"^[+-]?[\d.]+[+-]\d+$", s):          # 1.5-300: exponent letter omitted
        s = re.sub(r"([\d.])([+-]\d+)$", r"\1E\2", s)
    return float(s)


def read(path: str):
    """Returns dict(title, key, type, m, n, ptr, row, val) -- ptr/row 1-based as in the file;
    val is None for pattern files.  Symmetric matrices hold the lower triangle."""
    with open(path) as f:
        lines = f.readlines()
    title, key = lines[0][:72].rstrip(), lines[0][72:80].strip()
    code = lines[2][:3].lower()
    if code[0] not in "rcipq" or code[1] not in "suhzr" or code[2] != "a":
        raise ValueError(f"not an assembled Rutherford-Boeing matrix (type {code!r})")
    if code[0] not in "rp":
        raise ValueError(f"only real / pattern matrices are supported (type {code!r})")
    t = lines[2][14:].split()
    m, n, nnz = int(t[0]), int(t[1]), int(t[2])
    pf, rf, vf = lines[3][:16], lines[3][16:32], lines[3][32:52]
    pos = 4
    ptr, pos = _read_block(lines, pos, n + 1, pf, int)
    row, pos = _read_block(lines, pos, nnz, rf, int)
    val = None
    if code[0] == "r":
        val, pos = _read_block(lines, pos, nnz, vf, _to_float)
        val = np.array(val, dtype=np.float64)
    return dict(title=title, key=key, type=code, m=m, n=n, ptr=np.array(ptr, dtype=np.int64),
                row=np.array(row, dtype=np.int32), val=val)


def _int_format(maxval: int):
    prec = int(math.log10(max(maxval, 1))) + 2          # create_format's precision rule (:676-684)
    per_line = 80 // prec
    return per_line, prec, f"({per_line}i{prec})"


def write(path: str, n: int, ptr, row, val=None, symmetric: bool = True, title: str = "Matrix", key: str = "0",
          val_format: str = "(3e24.16)"):
    """Writes what SPRAL's rb_write writes (same header layout and format rules)."""
    ptr = np.asarray(ptr, dtype=np.int64)
    row = np.asarray(row, dtype=np.int32)
    nnz = int(ptr[n] - 1)
    pp, pw, pf = _int_format(int(ptr.max()))
    rp, rw, rf = _int_format(int(row[:nnz].max()) if nnz else 1)
    vp, vw = _parse_format(val_format)
    digits = int(val_format.split(".")[1].rstrip(")"))
    plines = n // pp + 1
    rlines = (nnz - 1) // rp + 1 if nnz else 1
    vlines = ((nnz - 1) // vp + 1 if nnz else 1) if val is not None else 0
    code = ("r" if val is not None else "p") + ("s" if symmetric else "u") + "a"
    with open(path, "w") as f:
        f.write(f"{title[:72]:<72}{key[:8]:<8}\n")
        f.write(f"{plines + rlines + vlines:14d} {plines:13d} {rlines:13d} {vlines:13d}\n")
        f.write(f"{code:<3}{'':11}{n:14d} {n:13d} {nnz:13d} {0:13d}\n")
        f.write(f"{pf:<16}{rf:<16}{val_format:<20}\n")

        def block(values, per_line, fmt):
            for i in range(0, len(values), per_line):
                f.write("".join(fmt(v) for v in values[i:i + per_line]) + "\n")

        block(ptr[: n + 1].tolist(), pp, lambda v: f"{v:{pw}d}")
        block(row[:nnz].tolist(), rp, lambda v: f"{v:{rw}d}")
        if val is not None:
            block(np.asarray(val, dtype=np.float64)[:nnz].tolist(), vp, lambda v: f"{v:{vw}.{digits}e}")
