"""TEST INFRASTRUCTURE (oracle): SPRAL's pseudo-random generator and random symmetric test
matrices, restated so that the reference's own scaling tests can be replayed on exactly their
inputs (/root/reference/spral/src/random.f90, src/random_matrix.f90:26-245,
tests/scaling.f90:866-907 gen_random_sym).  Only tests/ may import this."""
from __future__ import annotations

import numpy as np

_A, _C, _M = 1103515245, 12345, 2 ** 31


class RandomState:
    def __init__(self, seed: int = 486502):                # random.f90:21
        self.x = seed

    def _step(self):
        self.x = (_A * self.x + _C) % _M                   # int(mod(a*state%x+c, m))

    def real(self, positive: bool = False) -> float:       # random_real :45-62
        self._step()
        if positive:
            return float(self.x) / float(_M)
        return 1.0 - 2.0 * float(self.x) / float(_M)

    def integer(self, n: int) -> int:                      # random_integer64 :64-77
        if n <= 0:
            return n
        self._step()
        return int(self.x * (float(n) / float(_M))) + 1

    def integer_in_range(self, lo: int, hi: int) -> int:   # random_matrix.f90:302-309
        return lo + self.integer(hi - lo + 1) - 1

    def sym_wt_integer(self, n: int) -> int:               # random_matrix.f90:283-297
        r1 = self.integer(n)
        r2 = self.integer(n)
        while r2 < r1:
            r1 = self.integer(n)
            r2 = self.integer(n)
        return r1


def random_matrix_generate_sym(state: RandomState, n: int, nnz: int):
    """random_matrix_generate64 for a symmetric matrix with nonsingular=.true., sort=.true.
    (random_matrix.f90:70-245): lower triangle CSC, 1-based, diagonal forced, rows sorted."""
    m = n
    cnt = [0] * (n + 1)
    for i in range(1, n + 1):                              # forced diagonal (rperm = cperm = identity)
        cnt[i] += 1
    for _ in range(nnz - min(m, n)):
        j = state.sym_wt_integer(n)
        while cnt[j] >= (m - j + 1):
            j = state.sym_wt_integer(n)
        cnt[j] += 1
    ptr = [0, 1]
    rows = [0]
    rused = [False] * (m + 1)
    for i in range(1, n + 1):
        start = len(rows)
        rows.append(i)                                     # k = rperm(cperm(i)) = i
        rused[i] = True
        for _ in range(cnt[i] - 1):
            k = state.integer_in_range(i, m)
            while rused[k]:
                k = state.integer_in_range(i, m)
            rows.append(k)
            rused[k] = True
        for jj in range(start, len(rows)):
            rused[rows[jj]] = False
        ptr.append(ptr[-1] + cnt[i])
        rows[start:] = sorted(rows[start:])                # sort=.true. (dbl_tr_sort)
    nz = ptr[n + 1] - 1
    val = [state.real() for _ in range(nz)]                # values drawn after the sort
    return (np.array(ptr[1:], dtype=np.int64), np.array(rows[1:], dtype=np.int32), np.array(val, dtype=np.float64))


def gen_random_sym(state: RandomState, n: int, nza: int):
    """tests/scaling.f90:866-907 without zr."""
    ptr, row, val = random_matrix_generate_sym(state, n, nza)
    if n > 3:
        l = state.integer(n // 2)
        for k in range(1, n + 1, max(1, l)):               # some zeros on the diagonal
            if ptr[k] > ptr[k - 1] + 1:
                val[ptr[k - 1] - 1] = 0.0
        for k in range(1, n + 1):                          # some large off-diagonals
            val[ptr[k] - 2] = val[ptr[k] - 2] * 1000
    return ptr, row, val
