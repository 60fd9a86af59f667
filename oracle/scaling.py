"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's norm-equilibration scaling.

Only tests/ may import this.  It restates SPRAL's inf_norm_equilib_sym
(/root/reference/spral/src/scaling.f90:480-521, the routine behind options%scaling >= 4,
/root/reference/src/spldlt_factorize_mod.F90:804-831): Algorithm 1 of Knight, Ruiz, Ucar,
"A Symmetry Preserving Algorithm for Matrix Scaling" with equilib_options' defaults
(max_iterations = 10, tol = 1e-8 as a default REAL; scaling.f90:29-32).

Parity: the Fortran cannot be compiled here (no Fortran compiler), so this restatement is
pinned by definition only -- PARITY UNPINNED against a run of the reference; the product's C++
(sylver_b200/csrc/api.cpp::equilib_scale_sym) is compared with it bit for bit
(tests/test_scaling.py), and both are checked against the property the algorithm converges to
(every row of |S A S| has infinity norm 1 within tol once the iteration has converged).
"""
from __future__ import annotations

import numpy as np


def inf_norm_equilib_sym(n: int, ptr: np.ndarray, row: np.ndarray, val: np.ndarray,
                         max_iterations: int = 10, tol: float = float(np.float32(1e-8))):
    """Lower triangle CSC, 1-based ptr/row (scaling.f90:480-521).  Returns (scaling, iterations)."""
    ptr = np.asarray(ptr, dtype=np.int64)
    r = np.asarray(row, dtype=np.int64)[: ptr[n] - 1] - 1
    v = np.asarray(val, dtype=np.float64)[: ptr[n] - 1]
    c = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr[: n + 1]))
    scaling = np.ones(n)                                   # scaling(1:n) = 1.0        (:497)
    itr = 1
    while itr <= max_iterations:                           # do itr = 1, max_iterations (:498)
        a = np.abs(scaling[r] * v * scaling[c])            # abs(scaling(r)*val(j)*scaling(c)) (:505)
        maxentry = np.zeros(n)
        np.maximum.at(maxentry, r, a)                      # maxentry(r) = max(maxentry(r), v) (:506)
        np.maximum.at(maxentry, c, a)                      # maxentry(c) = max(maxentry(c), v) (:507)
        nz = maxentry > 0                                  # beware empty cols          (:511)
        scaling[nz] = scaling[nz] / np.sqrt(maxentry[nz])
        if n == 0 or np.abs(1 - maxentry).max() < tol:     # convergence test           (:514)
            break
        itr += 1
    return scaling, itr - 1                                # inform%iterations = itr-1  (:516)


# ---------------------------------------------------------------------------------------------
# options%scaling == 2: auction_scale_sym (scaling.f90:269-309) -> auction_match (:1504-1609) ->
# half_to_full (spral/src/matrix_util.f90:3167-3302) -> auction_match_core (:1351-1489) ->
# match_postproc, square case (:1631-1638).  Pure-Python loops over 1-based lists (element 0 is
# a dummy) so that every statement can be read against the Fortran; math.log / math.exp are the
# C library's (numpy's vectorised versions may differ in the last bit).  Small cases only.
# ---------------------------------------------------------------------------------------------
import math

_HUGE = float(np.finfo(np.float64).max)


def _half_to_full(n, row, ptr, a):
    iw = [0] * (n + 1)
    oldtau = ptr[n + 1] - 1
    ndiag = 0
    for j in range(1, n + 1):                              # matrix_util.f90:3218-3231
        i1, i2 = ptr[j], ptr[j + 1] - 1
        iw[j] += (i2 - i1) + 1
        for ii in range(i1, i2 + 1):
            i = row[ii]
            if i != j:
                iw[i] += 1
            else:
                ndiag += 1
    newtau = 2 * oldtau - ndiag
    ipkp1 = oldtau + 1
    ckp1 = newtau + 1
    for j in range(n, 0, -1):                              # :3239-3271
        i1 = ptr[j]
        i2 = ipkp1
        lenk = i2 - i1
        jstart = ckp1
        ipkp1 = i1
        i2 -= 1
        for ii in range(i2, i1 - 1, -1):
            jstart -= 1
            a[jstart] = a[ii]
            row[jstart] = row[ii]
        ptr[j] = jstart
        ckp1 -= iw[j]
        iw[j] = lenk
    for j in range(n, 0, -1):                              # :3276-3299
        i1, i2 = ptr[j], ptr[j] + iw[j] - 1
        for ii in range(i1, i2 + 1):
            i = row[ii]
            if i == j:
                continue
            ptr[i] -= 1
            ipos = ptr[i]
            a[ipos] = a[ii]
            row[ipos] = j
    ptr[n + 1] = newtau + 1


def _auction_match_core(m, n, ptr, row, val, dualv):
    f32 = np.float32
    max_iterations = 30000                                 # auction_options defaults, scaling.f90:33-38
    max_unchanged = (10, 100, 100)
    min_proportion = (f32(0.90), f32(0.0), f32(0.0))
    eps_initial = f32(0.01)
    unmatchable = 0
    owner = [0] * (m + 1)
    nxt = list(range(n + 1))
    minmn = min(m, n)
    unmatched = minmn
    match = [0] * (n + 1)
    dualu = [0.0] * (m + 1)
    prev, nunchanged = -1, 0
    tail = n
    eps = float(eps_initial)
    itr = 1
    while itr <= max_iterations:                           # :1418
        if unmatched == 0:
            break
        if unmatched != prev:
            nunchanged = 0
        prev = unmatched
        nunchanged += 1
        prop = f32(minmn - unmatched) / f32(minmn)         # real(minmn-unmatched)/minmn  :1426-1431
        if nunchanged >= max_unchanged[0] and prop >= min_proportion[0]:
            break
        if nunchanged >= max_unchanged[1] and prop >= min_proportion[1]:
            break
        if nunchanged >= max_unchanged[2] and prop >= min_proportion[2]:
            break
        eps = min(1.0, eps + 1.0 / (n + 1))                # :1433
        insert = 0
        for cptr in range(1, tail + 1):                    # :1437
            col = nxt[cptr]
            if match[col] != 0:
                continue
            if ptr[col] == ptr[col + 1]:
                continue
            j = ptr[col]
            bestr = row[j]
            bestu = val[j] - dualu[bestr]
            bestv = -_HUGE
            for j in range(ptr[col] + 1, ptr[col + 1]):    # :1449-1458
                u = val[j] - dualu[row[j]]
                if u > bestu:
                    bestv = bestu
                    bestr = row[j]
                    bestu = u
                elif u > bestv:
                    bestv = u
            if bestv == -_HUGE:
                bestv = 0.0
            if bestu > 0:                                  # :1461-1476
                dualu[bestr] = dualu[bestr] + bestu - bestv + eps
                dualv[col] = bestv - eps
                match[col] = bestr
                unmatched -= 1
                k = owner[bestr]
                owner[bestr] = col
                if k != 0:
                    match[k] = 0
                    unmatched += 1
                    insert += 1
                    nxt[insert] = k
            else:                                          # :1477-1482
                match[col] = -1
                unmatched -= 1
                unmatchable += 1
        tail = insert
        itr += 1
    iterations = itr - 1
    match = [0 if v == -1 else v for v in match]
    return match, dualu, iterations, unmatchable


def auction_scale_sym(n: int, ptr, row, val):
    """Lower triangle CSC, 1-based ptr/row.  Returns (scaling, match, inform dict) with
    match[i] = column (1-based) matched to row i+1, 0 = unmatched."""
    ptr = [0] + [int(x) for x in ptr[: n + 1]]
    m = n
    ne = 2 * (ptr[n + 1] - 1)
    ptr2 = [0] * (n + 2)
    row2 = [0] * (ne + 1)
    val2 = [0.0] * (ne + 1)
    cmax = [0.0] * (n + 1)
    k = 1
    for i in range(1, n + 1):                              # scaling.f90:1543-1555
        ptr2[i] = k
        for j in range(ptr[i], ptr[i + 1]):
            if val[j - 1] == 0.0:
                continue
            row2[k] = int(row[j - 1])
            val2[k] = abs(float(val[j - 1]))
            k += 1
        for j in range(ptr2[i], k):
            val2[j] = math.log(val2[j])
    ptr2[n + 1] = k
    _half_to_full(n, row2, ptr2, val2)
    for i in range(1, n + 1):                              # :1573-1582
        if ptr2[i + 1] <= ptr2[i]:
            cmax[i] = 0.0
            continue
        colmax = max(val2[ptr2[i]:ptr2[i + 1]])
        cmax[i] = colmax
        for j in range(ptr2[i], ptr2[i + 1]):
            val2[j] = colmax - val2[j]
    maxentry = max(val2[1:ptr2[n + 1]]) if ptr2[n + 1] > 1 else -_HUGE
    maxentry = 2 * maxentry + 1                            # :1586
    for j in range(1, ptr2[n + 1]):
        val2[j] = maxentry - val2[j]
    cscaling = [0.0] + [-cmax[i] for i in range(1, n + 1)]
    cmatch, rscaling, iterations, unmatchable = _auction_match_core(m, n, ptr2, row2, val2, cscaling)
    matched = sum(1 for i in range(1, n + 1) if cmatch[i] != 0)
    rscaling = [0.0] + [-rscaling[i] + maxentry for i in range(1, m + 1)]      # :1598-1599
    cscaling = [0.0] + [-cscaling[i] - cmax[i] for i in range(1, n + 1)]
    match = [0] * m
    for i in range(1, n + 1):
        if cmatch[i] != 0:
            match[cmatch[i] - 1] = i
    if n > 0:                                              # match_postproc, square: :1631-1638
        rsum = 0.0
        for i in range(1, m + 1):
            rsum += rscaling[i]
        csum = 0.0
        for i in range(1, n + 1):
            csum += cscaling[i]
        adjust = (rsum / m - csum / n) / 2
        rscaling = [0.0] + [rscaling[i] - adjust for i in range(1, m + 1)]
        cscaling = [0.0] + [cscaling[i] + adjust for i in range(1, n + 1)]
    scaling = np.array([math.exp((rscaling[i] + cscaling[i]) / 2) for i in range(1, n + 1)])
    return scaling, np.array(match, dtype=np.int32), dict(flag=0, matched=matched, iterations=iterations,
                                                            unmatchable=unmatchable)
