"""TEST INFRASTRUCTURE (oracle): restatement of SPRAL's matching-based ordering
(/root/reference/spral/src/match_order.f90:135-629 match_order_metis -> mo_scale -> mo_match ->
mo_split; expand_matrix /root/reference/spral/src/ssids/anal.f90:87-138), the routine behind
options%ordering = 2 (/root/reference/src/spldlt_analyse_mod.F90:788-817).

Only tests/ may import this.  The Fortran cannot be compiled here -> PARITY UNPINNED against a
run of the reference; METIS itself is a third-party library: the caller passes it in as a
function (tests use the product's thin wrapper around the CUDA toolkit's METIS 5), so that
everything AROUND the METIS call is what is restated and compared.  Pure-Python, 1-based lists.
"""
from __future__ import annotations

import math

import numpy as np

from oracle.scaling import _HUGE, _hungarian_match


def expand_matrix(n, ptr, row, val):                       # anal.f90:87-138
    aptr = [0] * (n + 2)
    nz = ptr[n + 1] - 1
    arow = [0] * (2 * nz + 1)
    aval = [0.0] * (2 * nz + 1)
    for j in range(1, n + 1):
        for kk in range(ptr[j], ptr[j + 1]):
            i = row[kk]
            aptr[i] += 1
            if j == i:
                continue
            aptr[j] += 1
    for j in range(2, n + 1):
        aptr[j] = aptr[j - 1] + aptr[j]
    aptr[n + 1] = aptr[n] + 1
    for j in range(1, n + 1):
        for kk in range(ptr[j], ptr[j + 1]):
            i = row[kk]
            atemp = val[kk]
            ipos = aptr[i]
            arow[ipos] = j
            aval[ipos] = atemp
            aptr[i] = ipos - 1
            if j == i:
                continue
            jpos = aptr[j]
            arow[jpos] = i
            aval[jpos] = atemp
            aptr[j] = jpos - 1
    for j in range(1, n + 1):
        aptr[j] += 1
    return aptr, arow, aval


def mo_match(n, ptr2, row2, val2):                         # match_order.f90:495-629
    cmax = [0.0] * (n + 1)
    for i in range(1, n + 1):
        seg = val2[ptr2[i]:ptr2[i + 1]]
        colmax = max(0.0, max(seg)) if seg else 0.0
        if colmax != 0.0:
            colmax = math.log(colmax)
        cmax[i] = colmax
    for i in range(1, n + 1):
        for j in range(ptr2[i], ptr2[i + 1]):
            val2[j] = cmax[i] - math.log(val2[j])
    cperm, rank, dualu, dualv = _hungarian_match(n, n, ptr2, row2, val2)
    scale = [0.0] * (n + 1)
    if rank == n:
        for i in range(1, n + 1):
            scale[i] = (dualu[i] + dualv[i] - cmax[i]) / 2
        return 0, scale, list(cperm)
    old_to_new = [0] * (n + 1)
    new_to_old = [0] * (n + 1)
    k = 0
    for i in range(1, n + 1):
        if cperm[i] < 0:
            old_to_new[i] = -1
        else:
            k += 1
            old_to_new[i] = k
            new_to_old[k] = i
    nne = 0
    k = 0
    ptr2[1] = 1
    j2 = 1
    for i in range(1, n + 1):
        j1 = j2
        j2 = ptr2[i + 1]
        if cperm[i] < 0:
            continue
        k += 1
        for jl in range(j1, j2):
            jj = row2[jl]
            if cperm[jj] < 0:
                continue
            nne += 1
            row2[nne] = old_to_new[jj]
            val2[nne] = val2[jl]
        ptr2[k + 1] = nne + 1
    nn = k
    cperm2, rank, dualu, dualv = _hungarian_match(nn, nn, ptr2, row2, val2)
    for i in range(1, n + 1):
        j = old_to_new[i]
        scale[i] = -_HUGE if j < 0 else (dualu[j] + dualv[j] - cmax[i]) / 2
    perm = [-1] * (n + 1)
    for i in range(1, nn + 1):
        perm[new_to_old[i]] = new_to_old[cperm2[i]]
    return 1, scale, perm


def match_order_metis(n: int, ptr, row, val, metis):
    """Lower triangle CSC, 1-based ptr/row.  `metis(ncomp, ptr3, row3)` must return the 1-based
    positions METIS_NodeND gives on that lower-triangle pattern (numpy arrays in and out).
    Returns (flag, order, scale, pairs, (ptr3, row3))."""
    ptr = [0] + [int(x) for x in ptr[: n + 1]]
    row = [0] + [int(x) for x in row]
    val = [0.0] + [float(x) for x in val]
    aptr, arow, aval = expand_matrix(n, ptr, row, val)
    ne = aptr[n + 1] - 1
    ptr2 = [0] * (n + 2)
    row2 = [0] * (ne + 1)
    val2 = [0.0] * (ne + 1)
    k = 1
    for i in range(1, n + 1):                              # :171-181
        ptr2[i] = k
        for j in range(aptr[i], aptr[i + 1]):
            if aval[j] == 0.0:
                continue
            row2[k] = arow[j]
            val2[k] = abs(aval[j])
            k += 1
    ptr2[n + 1] = k
    flag, scale, cperm = mo_match(n, list(ptr2), list(row2), list(val2))
    # mo_split :220-396
    iwork = [0] * (n + 1)
    for i in range(1, n + 1):
        if iwork[i] != 0:
            continue
        j = i
        while True:
            if cperm[j] == -1:
                iwork[j] = -2
                break
            elif cperm[j] == i:
                iwork[j] = -1
                break
            jj = cperm[j]
            iwork[j] = jj
            iwork[jj] = j
            j = cperm[jj]
            if j == i:
                break
    cperm = list(iwork)
    old_to_new = [0] * (n + 1)
    new_to_old = [0] * (n + 1)
    k = 1
    for i in range(1, n + 1):
        j = cperm[i]
        if j < i and j > 0:
            continue
        old_to_new[i] = k
        new_to_old[k] = i
        if j > 0:
            old_to_new[j] = k
        k += 1
    ncomp_matched = k - 1
    ptr3 = [0] * (n + 2)
    row3 = [0] * (ne + 1)
    iwork = [0] * (n + 1)
    ptr3[1] = 1
    ncomp = 1
    jj = 1
    for i in range(1, n + 1):
        j = cperm[i]
        if j < i and j > 0:
            continue
        for col in ((i, j) if j > 0 else (i,)):
            for kl in range(ptr2[col], ptr2[col + 1]):
                krow = old_to_new[row2[kl]]
                if iwork[krow] == i:
                    continue
                if krow > ncomp_matched:
                    continue
                row3[jj] = krow
                jj += 1
                iwork[krow] = i
        ptr3[ncomp + 1] = jj
        ncomp += 1
    ncomp -= 1
    ptr3[1] = 1
    jj = 1
    j1 = 1
    for i in range(1, ncomp + 1):
        j2 = ptr3[i + 1]
        for kl in range(j1, j2):
            krow = row3[kl]
            if krow < i:
                continue
            row3[jj] = krow
            jj += 1
        ptr3[i + 1] = jj
        j1 = j2
    p3 = np.array(ptr3[1: ncomp + 2], dtype=np.int64)
    r3 = np.array(row3[1: ptr3[ncomp + 1]], dtype=np.int32)
    corder = metis(ncomp, p3, r3)
    iwork = [0] * (n + 1)
    for i in range(1, ncomp + 1):
        iwork[int(corder[i - 1])] = i
    order = [0] * (n + 1)
    k = 1
    for i in range(1, ncomp + 1):
        j = new_to_old[iwork[i]]
        order[j] = k
        k += 1
        if cperm[j] > 0:
            j = cperm[j]
            order[j] = k
            k += 1
    scale_out = np.array([math.exp(scale[i]) for i in range(1, n + 1)])
    return (flag, np.array(order[1:], dtype=np.int32), scale_out, np.array(cperm[1:], dtype=np.int32), (p3, r3))
