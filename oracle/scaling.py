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


# ---------------------------------------------------------------------------------------------
# options%scaling == 1: hungarian_scale_sym (scaling.f90:134-170) -> hungarian_wrapper (:596-801)
# -> hungarian_match (:938-1194), hungarian_init_heurisitic (:810-929), heap (:1206-1325).
# Same conventions as the auction restatement above (1-based lists, C library log/exp).
# ---------------------------------------------------------------------------------------------
_RINF = _HUGE


def _heap_update(idx, Q, val, L):                          # :1206-1239
    pos = L[idx]
    if pos <= 1:
        Q[pos] = idx
        return
    v = val[idx]
    while pos > 1:
        parent_pos = pos // 2
        parent_idx = Q[parent_pos]
        if v >= val[parent_idx]:
            break
        Q[pos] = parent_idx
        L[parent_idx] = pos
        pos = parent_pos
    Q[pos] = idx
    L[idx] = pos


def _heap_delete(pos0, qlen, Q, D, L):                     # :1268-1325, returns the new qlen
    if qlen == pos0:
        return qlen - 1
    idx = Q[qlen]
    v = D[idx]
    qlen -= 1
    pos = pos0
    if pos > 1:
        while True:
            parent = pos // 2
            qk = Q[parent]
            if v >= D[qk]:
                break
            Q[pos] = qk
            L[qk] = pos
            pos = parent
            if pos <= 1:
                break
    Q[pos] = idx
    L[idx] = pos
    if pos != pos0:
        return qlen
    while True:
        child = 2 * pos
        if child > qlen:
            break
        dk = D[Q[child]]
        if child < qlen:
            dr = D[Q[child + 1]]
            if dk > dr:
                child += 1
                dk = dr
        if v <= dk:
            break
        qk = Q[child]
        Q[pos] = qk
        L[qk] = pos
        pos = child
    Q[pos] = idx
    L[idx] = pos
    return qlen


def _hungarian_init_heuristic(m, n, ptr, row, val, iperm, jperm, dualu, d, l, search_from):
    num = 0
    for i in range(1, m + 1):                              # :838-839
        dualu[i] = _RINF
        l[i] = 0
    for j in range(1, n + 1):                              # :840-848
        for k in range(ptr[j], ptr[j + 1]):
            i = row[k]
            if val[k] > dualu[i]:
                continue
            dualu[i] = val[k]
            iperm[i] = j
            l[i] = k
    for i in range(1, m + 1):                              # :852-863
        j = iperm[i]
        if j == 0:
            continue
        iperm[i] = 0
        if jperm[j] != 0:
            continue
        if (ptr[j + 1] - ptr[j] > m // 10) and (m > 50):
            continue
        num += 1
        iperm[i] = j
        jperm[j] = l[i]
    if num == min(m, n):
        return num
    for j in range(1, n + 1):                              # :870-871
        d[j] = 0.0
        search_from[j] = ptr[j]
    for j in range(1, n + 1):                              # improve_assign :872-927
        if jperm[j] != 0:
            continue
        if ptr[j] > ptr[j + 1] - 1:
            continue
        i0 = row[ptr[j]]
        vj = val[ptr[j]] - dualu[i0]
        k0 = ptr[j]
        for k in range(ptr[j] + 1, ptr[j + 1]):
            i = row[k]
            di = val[k] - dualu[i]
            if di > vj:
                continue
            if di == vj and di != _RINF:
                if iperm[i] != 0 or iperm[i0] == 0:
                    continue
            vj = di
            i0 = i
            k0 = k
        d[j] = vj
        if iperm[i0] == 0:
            num += 1
            jperm[j] = k0
            iperm[i0] = j
            search_from[j] = k0 + 1
            continue
        done = False
        for k in range(k0, ptr[j + 1]):
            i = row[k]
            if (val[k] - dualu[i]) > vj:
                continue
            jj = iperm[i]
            for kk in range(search_from[jj], ptr[jj + 1]):
                ii = row[kk]
                if iperm[ii] > 0:
                    continue
                if (val[kk] - dualu[ii]) <= d[jj]:
                    jperm[jj] = kk
                    iperm[ii] = jj
                    search_from[jj] = kk + 1
                    num += 1
                    jperm[j] = k
                    iperm[i] = j
                    search_from[j] = k + 1
                    done = True
                    break
            if done:
                break
            search_from[jj] = ptr[jj + 1]
    return num


def _hungarian_match(m, n, ptr, row, val):
    """Returns (iperm, num, dualu, dualv), all 1-based lists."""
    jperm = [0] * (n + 1)
    out = [0] * (n + 1)
    pr = [0] * (n + 1)
    q = [0] * (m + 2)
    longwork = [0] * (m + 1)
    l = [0] * (m + 1)
    d = [0.0] * (max(m, n) + 1)
    iperm = [0] * (m + 1)
    dualu = [0.0] * (m + 1)
    dualv = [0.0] * (n + 1)
    num = _hungarian_init_heuristic(m, n, ptr, row, val, iperm, jperm, dualu, d, longwork, out)
    if num != min(m, n):
        for i in range(1, m + 1):                          # :983-984
            d[i] = _RINF
            l[i] = 0
        isp, jsp = -1, -1
        for jord in range(1, n + 1):                       # :986
            if jperm[jord] != 0:
                continue
            dmin = _RINF
            qlen = 0
            low = m + 1
            up = m + 1
            csp = _RINF
            j = jord
            pr[j] = -1
            for klong in range(ptr[j], ptr[j + 1]):        # :1004-1018
                i = row[klong]
                dnew = val[klong] - dualu[i]
                if dnew >= csp:
                    continue
                if iperm[i] == 0:
                    csp = dnew
                    isp = klong
                    jsp = j
                else:
                    if dnew < dmin:
                        dmin = dnew
                    d[i] = dnew
                    qlen += 1
                    longwork[qlen] = klong
            q0 = qlen
            qlen = 0
            for kk in range(1, q0 + 1):                    # :1022-1043
                klong = longwork[kk]
                i = row[klong]
                if csp <= d[i]:
                    d[i] = _RINF
                    continue
                if d[i] <= dmin:
                    low -= 1
                    q[low] = i
                    l[i] = low
                else:
                    qlen += 1
                    l[i] = qlen
                    _heap_update(i, q, d, l)
                jj = iperm[i]
                out[jj] = klong
                pr[jj] = j
            for _jdum in range(1, num + 1):                # :1045
                if low == up:
                    if qlen == 0:
                        break
                    i = q[1]
                    if d[i] >= csp:
                        break
                    dmin = d[i]
                    while qlen > 0:
                        i = q[1]
                        if d[i] > dmin:
                            break
                        i = q[1]
                        qlen = _heap_delete(1, qlen, q, d, l)      # heap_pop
                        low -= 1
                        q[low] = i
                        l[i] = low
                q0 = q[up - 1]
                dq0 = d[q0]
                if dq0 >= csp:
                    break
                up -= 1
                j = iperm[q0]
                vj = dq0 - val[jperm[j]] + dualu[q0]
                for klong in range(ptr[j], ptr[j + 1]):    # :1072-1110
                    i = row[klong]
                    if l[i] >= up:
                        continue
                    dnew = vj + val[klong] - dualu[i]
                    if dnew >= csp:
                        continue
                    if iperm[i] == 0:
                        csp = dnew
                        isp = klong
                        jsp = j
                    else:
                        di = d[i]
                        if di <= dnew:
                            continue
                        if l[i] >= low:
                            continue
                        d[i] = dnew
                        if dnew <= dmin:
                            lpos = l[i]
                            if lpos != 0:
                                qlen = _heap_delete(lpos, qlen, q, d, l)
                            low -= 1
                            q[low] = i
                            l[i] = low
                        else:
                            if l[i] == 0:
                                qlen += 1
                                l[i] = qlen
                            _heap_update(i, q, d, l)
                        jj = iperm[i]
                        out[jj] = klong
                        pr[jj] = j
            if csp != _RINF:                               # :1114-1135
                num += 1
                i = row[isp]
                iperm[i] = jsp
                jperm[jsp] = isp
                j = jsp
                for _jdum in range(1, num + 1):
                    jj = pr[j]
                    if jj == -1:
                        break
                    klong = out[j]
                    i = row[klong]
                    iperm[i] = jj
                    jperm[jj] = klong
                    j = jj
                for kk in range(up, m + 1):
                    i = q[kk]
                    dualu[i] = dualu[i] + d[i] - csp
            for kk in range(low, m + 1):                   # 190 :1136-1145
                i = q[kk]
                d[i] = _RINF
                l[i] = 0
            for kk in range(1, qlen + 1):
                i = q[kk]
                d[i] = _RINF
                l[i] = 0
    for j in range(1, n + 1):                              # 1000 :1152-1159
        klong = jperm[j]
        dualv[j] = val[klong] - dualu[row[klong]] if klong != 0 else 0.0
    for i in range(1, m + 1):
        if iperm[i] == 0:
            dualu[i] = 0.0
    if num == min(m, n):
        return iperm, num, dualu, dualv
    jperm = [0] * (n + 1)                                  # :1169-1193
    k = 0
    for i in range(1, m + 1):
        if iperm[i] == 0:
            k += 1
            out[k] = i
        else:
            jperm[iperm[i]] = i
    k = 0
    for j in range(1, n + 1):
        if jperm[j] != 0:
            continue
        k += 1
        iperm[out[k]] = -j
    return iperm, num, dualu, dualv


def hungarian_scale_sym(n: int, ptr, row, val, scale_if_singular: bool = False):
    """Lower triangle CSC, 1-based ptr/row.  Returns (scaling, match, inform dict)."""
    ptr = [0] + [int(x) for x in ptr[: n + 1]]
    m = n
    ne = 2 * (ptr[n + 1] - 1)
    ptr2 = [0] * (n + 2)
    row2 = [0] * (ne + 1)
    val2 = [0.0] * (ne + 1)
    cmax = [0.0] * (n + 1)
    k = 1
    for i in range(1, n + 1):                              # :637-649
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
    for i in range(1, n + 1):                              # :653-657
        colmax = max(val2[ptr2[i]:ptr2[i + 1]]) if ptr2[i + 1] > ptr2[i] else -_HUGE
        cmax[i] = colmax
        for j in range(ptr2[i], ptr2[i + 1]):
            val2[j] = colmax - val2[j]
    match, matched, dualu, dualv = _hungarian_match(m, n, ptr2, row2, val2)
    flag = 0
    if matched != min(m, n):                               # :666-677
        flag = 1 if scale_if_singular else -2
    rscaling = [0.0] * (m + 1)
    cscaling = [0.0] * (n + 1)
    if matched == n:                                       # :679-686
        for i in range(1, m + 1):
            rscaling[i] = dualu[i]
        for i in range(1, n + 1):
            cscaling[i] = dualv[i] - cmax[i]
        if n > 0:                                          # match_postproc, square
            rsum = 0.0
            for i in range(1, m + 1):
                rsum += rscaling[i]
            csum = 0.0
            for i in range(1, n + 1):
                csum += cscaling[i]
            adjust = (rsum / n - csum / n) / 2
            for i in range(1, m + 1):
                rscaling[i] = rscaling[i] - adjust
            for i in range(1, n + 1):
                cscaling[i] = cscaling[i] + adjust
    else:                                                  # :688-800
        old_to_new = [0] * (n + 1)
        new_to_old = [0] * (n + 1)
        j = matched + 1
        k = 0
        for i in range(1, m + 1):
            if match[i] < 0:
                old_to_new[i] = -j
                j += 1
            else:
                k += 1
                old_to_new[i] = k
                new_to_old[k] = i
        nent = 0
        k = 0
        ptr2[1] = 1
        j2 = 1
        for i in range(1, n + 1):
            j1 = j2
            j2 = ptr2[i + 1]
            if match[i] < 0:
                continue
            k += 1
            for jl in range(j1, j2):
                jj = row2[jl]
                if match[jj] < 0:
                    continue
                nent += 1
                row2[nent] = old_to_new[jj]
                val2[nent] = val2[jl]
            ptr2[k + 1] = nent + 1
        nn = k
        cperm, matched, dualu, dualv = _hungarian_match(nn, nn, ptr2, row2, val2)
        for i in range(1, n + 1):
            jn = old_to_new[i]
            rscaling[i] = -_HUGE if jn < 0 else (dualu[jn] + dualv[jn] - cmax[i]) / 2
        match = [-1] * (n + 1)
        for i in range(1, nn + 1):
            match[new_to_old[i]] = cperm[i]
        for i in range(1, n + 1):
            if match[i] == -1:
                match[i] = old_to_new[i]
        cscale = list(rscaling)
        for i in range(1, n + 1):
            for jl in range(ptr[i], ptr[i + 1]):
                kr = int(row[jl - 1])
                a = abs(float(val[jl - 1]))
                lg = math.log(a) if a > 0 else -math.inf
                if cscale[i] == -_HUGE and cscale[kr] != -_HUGE:
                    rscaling[i] = max(rscaling[i], lg + rscaling[kr])
                if cscale[kr] == -_HUGE and cscale[i] != -_HUGE:
                    rscaling[kr] = max(rscaling[kr], lg + rscaling[i])
        for i in range(1, n + 1):
            if cscale[i] != -_HUGE:
                continue
            rscaling[i] = 0.0 if rscaling[i] == -_HUGE else -rscaling[i]
        cscaling = list(rscaling)
    scaling = np.array([math.exp((rscaling[i] + cscaling[i]) / 2) for i in range(1, n + 1)])
    return scaling, np.array(match[1:], dtype=np.int32), dict(flag=flag, matched=matched)
