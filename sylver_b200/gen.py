"""Synthetic inputs for the BASELINE.json configs (harness-side, CPU, numpy).

Matrices are symmetric, lower triangle in CSC with 1-based ``ptr`` (int64) and
``row`` (int32) exactly as the SyLVER C API expects
(/root/reference/src/interfaces/C/sylver_ciface.F90:374-467).  Definitions follow
SURVEY.md section 8(d):

* config 1/5: 3D 7-point Laplacian (diag 6, off-diagonals -1)
* config 3:   3D 27-point Laplacian (diag 26, all 26 neighbours -1)
* config 4:   Stokes-like KKT saddle point [[A, B^T], [B, 0]]
* config 2:   dense random symmetric indefinite front (glibc rand(), seed 1;
              /root/reference/tests/common.hxx:752-775)

The pivot order is a deterministic geometric nested dissection (METIS is an
un-vendored, un-pinned dependency of the reference, so the order is an *input*
shared by the oracle and the B200 engine).
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np


# ----------------------------------------------------------------------------
# grid Laplacians
# ----------------------------------------------------------------------------
def _grid_lower(k: int, offsets, diag: float, off: float):
    """Lower-triangular CSC of a stencil on a k^3 grid (x fastest)."""
    n = k * k * k
    idx = np.arange(n, dtype=np.int64)
    x = idx % k
    y = (idx // k) % k
    z = idx // (k * k)
    cols = [idx]
    rows = [idx]
    vals = [np.full(n, diag)]
    for dx, dy, dz in offsets:
        ok = ((x + dx >= 0) & (x + dx < k) & (y + dy >= 0) & (y + dy < k) &
              (z + dz >= 0) & (z + dz < k))
        c = idx[ok]
        r = c + dx + k * (dy + k * dz)
        assert np.all(r > c)
        cols.append(c)
        rows.append(r)
        vals.append(np.full(c.size, off))
    cols = np.concatenate(cols)
    rows = np.concatenate(rows)
    vals = np.concatenate(vals)
    o = np.lexsort((rows, cols))
    cols, rows, vals = cols[o], rows[o], vals[o]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, cols + 1, 1)
    ptr = np.cumsum(ptr) + 1
    return n, ptr.astype(np.int64), (rows + 1).astype(np.int32), vals.astype(np.float64)


def laplacian_7pt(k: int):
    """3D 7-point Laplacian on a k^3 grid: 6 on the diagonal, -1 to +-x,+-y,+-z."""
    return _grid_lower(k, [(1, 0, 0), (0, 1, 0), (0, 0, 1)], 6.0, -1.0)


def laplacian_27pt(k: int):
    """3D 27-point Laplacian: 26 on the diagonal, -1 to all 26 neighbours."""
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
            if (dz, dy, dx) > (0, 0, 0)]          # the 13 neighbours with a larger index
    return _grid_lower(k, offs, 26.0, -1.0)


# ----------------------------------------------------------------------------
# geometric nested dissection on a k^3 grid
# ----------------------------------------------------------------------------
def nested_dissection_order(k: int, dofs_per_cell: int = 1, leaf: int = 4):
    """order[i] = 1-based pivot position of variable i.

    Recursive coordinate bisection: the longest side of the box is cut by a
    one-cell-thick plane; numbering is [first half, second half, separator].
    Boxes with every side <= ``leaf`` are numbered in natural order.  With
    ``dofs_per_cell`` > 1 the dofs of a cell stay adjacent (variable index =
    cell*dofs + d), d increasing.
    """
    n = k * k * k
    pos_of_cell = np.empty(n, dtype=np.int64)

    def cells(b):
        x0, x1, y0, y1, z0, z1 = b
        zz, yy, xx = np.meshgrid(np.arange(z0, z1), np.arange(y0, y1),
                                 np.arange(x0, x1), indexing="ij")
        return (xx + k * (yy + k * zz)).ravel()

    stack = [((0, k, 0, k, 0, k), 0)]
    while stack:
        b, start = stack.pop()
        x0, x1, y0, y1, z0, z1 = b
        sx, sy, sz = x1 - x0, y1 - y0, z1 - z0
        vol = sx * sy * sz
        if vol == 0:
            continue
        if max(sx, sy, sz) <= leaf:
            c = cells(b)
            pos_of_cell[c] = start + np.arange(c.size)
            continue
        # cut the longest side (ties: z, then y, then x -> planes of x-fastest cells)
        if sz >= sy and sz >= sx:
            mid = z0 + sz // 2
            left = (x0, x1, y0, y1, z0, mid)
            right = (x0, x1, y0, y1, mid + 1, z1)
            sep = (x0, x1, y0, y1, mid, mid + 1)
        elif sy >= sx:
            mid = y0 + sy // 2
            left = (x0, x1, y0, mid, z0, z1)
            right = (x0, x1, mid + 1, y1, z0, z1)
            sep = (x0, x1, mid, mid + 1, z0, z1)
        else:
            mid = x0 + sx // 2
            left = (x0, mid, y0, y1, z0, z1)
            right = (mid + 1, x1, y0, y1, z0, z1)
            sep = (mid, mid + 1, y0, y1, z0, z1)
        nl = (left[1] - left[0]) * (left[3] - left[2]) * (left[5] - left[4])
        nr = (right[1] - right[0]) * (right[3] - right[2]) * (right[5] - right[4])
        stack.append((left, start))
        stack.append((right, start + nl))
        c = cells(sep)
        pos_of_cell[c] = start + nl + nr + np.arange(c.size)
    if dofs_per_cell == 1:
        return (pos_of_cell + 1).astype(np.int32)
    d = dofs_per_cell
    order = (pos_of_cell[:, None] * d + np.arange(d)[None, :] + 1).ravel()
    return order.astype(np.int32)


# ----------------------------------------------------------------------------
# Stokes-like KKT system (config 4)
# ----------------------------------------------------------------------------
def stokes_kkt(k: int):
    """[[A, B^T], [B, 0]] on a k^3 grid, variables interleaved per cell.

    Variable index = 4*cell + d, d = 0,1,2 velocity components (each a 7-point
    Laplacian block, diag 6 / off -1), d = 3 pressure.  B is the forward
    difference divergence: row (pressure of cell c) has -1 on u_d(c) and +1 on
    u_d(c + e_d) when that neighbour exists.  The (2,2) block is exactly zero.
    Returns lower-triangle CSC, 1-based.
    """
    nc = k * k * k
    n = 4 * nc
    cell = np.arange(nc, dtype=np.int64)
    x = cell % k
    y = (cell // k) % k
    z = cell // (k * k)
    step = (1, k, k * k)
    coord = (x, y, z)
    R, C, V = [], [], []
    for d in range(3):
        v = 4 * cell + d
        R.append(v); C.append(v); V.append(np.full(nc, 6.0))
        for e in range(3):
            ok = coord[e] < k - 1
            c = cell[ok]
            R.append(4 * (c + step[e]) + d); C.append(4 * c + d)
            V.append(np.full(c.size, -1.0))
        # divergence row of cell c: -u_d(c) + u_d(c+e_d)
        p = 4 * cell + 3
        R.append(p); C.append(v); V.append(np.full(nc, -1.0))          # p(c) x u_d(c), p > v
        ok = coord[d] < k - 1
        c = cell[ok]
        # entry (p(c), u_d(c+e_d)) = +1 ; lower triangle needs row >= col
        rr = 4 * c + 3
        cc = 4 * (c + step[d]) + d
        lo = np.minimum(rr, cc)
        hi = np.maximum(rr, cc)
        R.append(hi); C.append(lo); V.append(np.full(c.size, 1.0))
    # explicit zero diagonal for the pressure block keeps every column non-empty
    p = 4 * cell + 3
    R.append(p); C.append(p); V.append(np.zeros(nc))
    R = np.concatenate(R); C = np.concatenate(C); V = np.concatenate(V)
    o = np.lexsort((R, C))
    R, C, V = R[o], C[o], V[o]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, C + 1, 1)
    ptr = np.cumsum(ptr) + 1
    return n, ptr.astype(np.int64), (R + 1).astype(np.int32), V.astype(np.float64)


def scale_down_some(n, ptr, row, val, frac: float, lo: float, hi: float):
    """Symmetric scaling S A S with s_i = 10^(lo + (hi-lo) h2(i)) for the fraction `frac` of the
    variables picked by a multiplicative hash of the index (seed-free), 1 elsewhere."""
    i = np.arange(n, dtype=np.uint64)
    h1 = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    h2 = ((i * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFF)).astype(np.float64) / 65536.0
    sc = np.where(h1 < np.uint64(int(0xFFFFFFFF * frac)), 10.0 ** (lo + (hi - lo) * h2), 1.0)
    col = np.repeat(np.arange(n), np.diff(ptr))
    return val * sc[col] * sc[row - 1]


KKTD_FRAC, KKTD_LO, KKTD_HI = 1.0 / 64, -6.0, -3.0


def stokes_kkt_delays(k: int):
    """BASELINE config 4 "with delayed pivots": the Stokes KKT system with a fraction of its
    variables scaled down symmetrically by 10^-6 .. 10^-3 (the idea of the reference harness's
    cause_delays, tests/common.hxx:162-179, applied to a sparse matrix): pivots that look
    acceptable inside a front fail the a-posteriori threshold test and are delayed up the
    tree.  Seed-free: the choice and the exponents are multiplicative hashes of the index."""
    n, ptr, row, val = stokes_kkt(k)
    return n, ptr, row, scale_down_some(n, ptr, row, val, KKTD_FRAC, KKTD_LO, KKTD_HI)


# ----------------------------------------------------------------------------
# dense fronts (config 2) -- reference generators tests/common.hxx:752-789
# ----------------------------------------------------------------------------
RAND_MAX = 2147483647


class GlibcRand:
    """Bit-exact restatement of glibc rand() (random_r TYPE_3, degree 31, sep 3)
    after srand(seed); the reference's generators call rand() unseeded, i.e.
    with the default seed 1 (tests/common.hxx:752-758)."""

    def __init__(self, seed: int = 1):
        r = [0] * 34
        r[0] = seed if seed != 0 else 1
        for i in range(1, 31):
            # 16807 * r mod (2^31 - 1) with glibc's signed Schrage arithmetic
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            if w < 0:
                w += 2147483647
            r[i] = w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self._r = r
        self._discard = 310

    def draw(self, count: int) -> np.ndarray:
        r = self._r
        total = self._discard + count
        mask = 0xFFFFFFFF
        ext = r + [0] * total
        base = len(r)
        for i in range(base, base + total):
            ext[i] = (ext[i - 31] + ext[i - 3]) & mask
        out = np.array(ext[base + self._discard:], dtype=np.int64) >> 1
        self._r = ext[-34:]
        self._discard = 0
        return out


def _check_against_libc(stream_head: np.ndarray, seed: int) -> None:
    try:
        libc = ctypes.CDLL(ctypes.util.find_library("c") or "libc.so.6")
        libc.srand(seed)
        chk = np.array([libc.rand() for _ in range(stream_head.size)], dtype=np.int64)
    except (OSError, AttributeError):
        return
    assert np.array_equal(stream_head, chk), "glibc rand() restatement mismatch"


def dense_sym_indef(m: int, seed: int = 1, rng: GlibcRand | None = None) -> np.ndarray:
    """gen_sym_indef (tests/common.hxx:752-775): a[j*lda+i] = 1 - (2*rand())/RAND_MAX
    for j outer, i >= j inner, then symmetrised.  Returns an (m, m) Fortran array."""
    own = rng is None
    rng = rng or GlibcRand(seed)
    cnt = m * (m + 1) // 2
    stream = rng.draw(cnt)
    if own:
        _check_against_libc(stream[:32], seed)
    vals = 1.0 - (2.0 * stream.astype(np.float64)) / float(RAND_MAX)
    a = np.zeros((m, m), order="F")
    jj, ii = np.triu_indices(m)          # (j, i) with j outer, i >= j inner
    a[ii, jj] = vals
    a[jj, ii] = vals
    return a


def dense_posdef(m: int, seed: int = 1) -> np.ndarray:
    """Diagonally dominant SPD front as documented for gen_posdef
    (tests/common.hxx:772-789): a_ii = |a_ii| + 0.1 + sum_{j != i} |a_ij|.
    (The reference loop additionally doubles the running diagonal twice; that
    quirk is not reproduced -- the posdef harness has no pinned outputs.)"""
    a = dense_sym_indef(m, seed=seed)
    off = np.abs(a).sum(axis=1) - np.abs(np.diag(a))
    a[np.arange(m), np.arange(m)] = np.abs(np.diag(a)) + 0.1 + off
    return a


def cause_delays(a: np.ndarray, rng: GlibcRand) -> np.ndarray:
    """cause_delays (tests/common.hxx:162-179): n/8 random rows/cols x1000 plus
    n/8 random single entries x1000, drawing from the same rand() stream."""
    n = a.shape[0]
    a = a.copy(order="F")
    nsing = max(1, n // 8)
    draws = rng.draw(3 * nsing)
    f = (np.float32(n) * draws.astype(np.float32)) / np.float32(RAND_MAX)
    pick = f.astype(np.int64)
    for i in range(nsing):
        idx, r, c = int(pick[3 * i]), int(pick[3 * i + 1]), int(pick[3 * i + 2])
        idx = min(idx, n - 1); r = min(r, n - 1); c = min(c, n - 1)
        a[idx, :] *= 1000.0
        a[:, idx] *= 1000.0
        a[idx, idx] /= 1000.0            # the diagonal is hit once by the reference loops
        if r != c:
            a[r, c] *= 1000.0
            a[c, r] *= 1000.0
        else:
            a[r, r] *= 1000.0
    return a


# ----------------------------------------------------------------------------
# helpers shared by tests and bench
# ----------------------------------------------------------------------------
def sym_matvec(n, ptr, row, val, x):
    """y = A x for a lower-triangle CSC symmetric matrix (1-based)."""
    import scipy.sparse as sp
    L = sp.csc_matrix((val, row.astype(np.int64) - 1, ptr - 1), shape=(n, n))
    D = L.diagonal()
    return L @ x + L.T @ x - D * x if x.ndim == 1 else L @ x + L.T @ x - D[:, None] * x


def backward_error(n, ptr, row, val, x, b):
    """||Ax-b||_inf / (||A||_inf ||x||_inf + ||b||_inf), true symmetric row sums
    (/root/reference/tests/common.hxx:1002-1049)."""
    import scipy.sparse as sp
    L = sp.csc_matrix((np.abs(val), row.astype(np.int64) - 1, ptr - 1), shape=(n, n))
    rowsum = np.asarray(L.sum(axis=1)).ravel() + np.asarray(L.sum(axis=0)).ravel() - L.diagonal()
    anorm = rowsum.max()
    r = sym_matvec(n, ptr, row, val, x) - b
    return float(np.abs(r).max() / (anorm * np.abs(x).max() + np.abs(b).max()))
