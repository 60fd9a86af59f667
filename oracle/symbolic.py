"""oracle/symbolic.py -- TEST INFRASTRUCTURE ONLY (checker, never the product path).

Pure-Python restatement of the reference's symbolic phase, used to check the
product's C++ analysis (sylver_b200/csrc/analyse.cpp) bit for bit on small
inputs.  PARITY UNPINNED: the reference implements this phase in Fortran
(no Fortran compiler exists in this environment and the reference ships no golden
vectors for it), so this oracle is pinned only by (a) definitional brute-force
checks in tests/test_symbolic.py (dense symbolic Cholesky) and (b) the fact that
the real SSIDS numeric engine (oracle/_ref) accepts its output and solves to
1e-16 backward error.

Every routine follows the cited Fortran routine statement by statement, with
1-based arrays (index 0 unused).

  expand_pattern   spral/src/ssids/anal.f90:37-81
  find_etree       spral/src/core_analyse.f90:173-224
  find_postorder   spral/src/core_analyse.f90:233-353
  find_col_counts  spral/src/core_analyse.f90:387-523
  find_supernodes  spral/src/core_analyse.f90:536-705 (+ do_merge 806-819, merge_nodes 824-853,
                   sort_by_val 712-800)
  apply_perm       spral/src/core_analyse.f90:1069-1100
  find_row_lists   spral/src/core_analyse.f90:911-1003
  dbl_tr_sort      spral/src/core_analyse.f90:1007-1065
  calc_stats       spral/src/core_analyse.f90:862-905
  build_map        src/spldlt_analyse_mod.F90:130-232
  analyse_core     src/spldlt_analyse_mod.F90:307-580 (order/invp fix-up only)
"""
from __future__ import annotations

import numpy as np

HUGE = 2 ** 63 - 1


def expand_pattern(n, ptr, row):
    """ptr,row 1-based lists (index 0 unused). Returns aptr, arow (1-based)."""
    nz = ptr[n + 1] - 1
    aptr = [0] * (n + 2)
    arow = [0] * (2 * nz + 1)
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
            arow[aptr[i]] = j
            aptr[i] -= 1
            if j == i:
                continue
            arow[aptr[j]] = i
            aptr[j] -= 1
    for j in range(1, n + 1):
        aptr[j] += 1
    return aptr, arow


def find_etree(n, ptr, row, perm, invp):
    parent = [0] * (n + 2)
    vforest = [n + 1] * (n + 2)
    for piv in range(1, n + 1):
        rowidx = invp[piv]
        for i in range(ptr[rowidx], ptr[rowidx + 1]):
            j = perm[row[i]]
            if j >= piv:
                continue
            k = j
            while vforest[k] < piv:
                l = vforest[k]
                vforest[k] = piv
                k = l
            if vforest[k] == piv:
                continue
            parent[k] = piv
            vforest[k] = piv
        parent[piv] = n + 1
    return parent


def find_postorder(n, ptr, perm, invp, parent):
    realn = n
    chead = [-1] * (n + 2)
    cnext = [-1] * (n + 2)
    for i in range(n, 0, -1):
        j = parent[i]
        cnext[i] = chead[j]
        chead[j] = i
    mp = [0] * (n + 2)
    stack = [n + 1]
    idn = n + 1
    while stack:
        node = stack.pop()
        mp[node] = idn
        idn -= 1
        if node == n + 1:
            i = chead[node]
            while i != -1:
                if ptr[invp[i] + 1] - ptr[invp[i]] != 0:
                    stack.append(i)
                i = cnext[i]
            i = chead[node]
            while i != -1:
                if ptr[invp[i] + 1] - ptr[invp[i]] == 0:
                    realn -= 1
                    stack.append(i)
                i = cnext[i]
        else:
            i = chead[node]
            while i != -1:
                stack.append(i)
                i = cnext[i]
    old = invp[:]
    for i in range(1, n + 1):
        invp[mp[i]] = old[i]
    for i in range(1, n + 1):
        perm[invp[i]] = i
    tmp = [0] * (n + 2)
    for i in range(1, n + 1):
        tmp[i] = mp[parent[i]]
    for i in range(1, n + 1):
        parent[mp[i]] = tmp[i]
    return realn


def _find(vforest, u):
    prev = -1
    cur = u
    while vforest[cur] != 0:
        prev = cur
        cur = vforest[cur]
        if vforest[cur] != 0:
            vforest[prev] = vforest[cur]
    return cur


def find_col_counts(n, ptr, row, perm, invp, parent):
    cc = [0] * (n + 2)
    first = list(range(n + 2))
    for i in range(1, n + 1):
        par = parent[i]
        first[par] = min(first[i], first[par])
        cc[i] = 1 if first[i] == i else 0
    cc[n + 1] = n + 1
    vforest = [0] * (n + 2)
    last_p = [0] * (n + 2)
    last_nbr = [0] * (n + 2)
    for piv in range(1, n + 1):
        col = invp[piv]
        for ii in range(ptr[col], ptr[col + 1]):
            u = perm[row[ii]]
            if u <= piv:
                continue
            if first[piv] > last_nbr[u]:
                cc[piv] += 1
                pp = last_p[u]
                if pp != 0:
                    lca = _find(vforest, pp)
                    cc[lca] -= 1
                last_p[u] = piv
            last_nbr[u] = piv
        par = parent[piv]
        cc[par] = cc[par] + cc[piv] - 1
        vforest[piv] = par
    return cc


def _sort_by_val(idx, val):
    """Stable descending sort (both Fortran variants are stable with >=)."""
    n = len(idx)
    if n >= 16:
        mid = (n - 1) // 2 + 1
        a = _sort_by_val(idx[:mid], val)
        b = _sort_by_val(idx[mid:], val)
        out = []
        j = k = 0
        while j < len(a) and k < len(b):
            if val[a[j]] >= val[b[k]]:
                out.append(a[j]); j += 1
            else:
                out.append(b[k]); k += 1
        out.extend(a[j:]); out.extend(b[k:])
        return out
    idx = idx[:]
    # insertion sort from the back, exactly as the Fortran loop
    kor = n          # 1-based in Fortran; here operate on 0-based list
    for _ in range(2, n + 1):
        ice_idx = idx[kor - 2]
        ice_val = val[ice_idx]
        k = kor - 1
        while k <= n - 1:
            ik_idx = idx[k]
            if ice_val >= val[ik_idx]:
                break
            idx[k - 1] = ik_idx
            k += 1
        idx[k - 1] = ice_idx
        kor -= 1
    return idx


def find_supernodes(n, realn, parent, cc, nemin):
    nelim = [1] * (n + 2)
    nvert = [1] * (n + 2)
    vhead = [-1] * (n + 2)
    vnext = [-1] * (n + 2)
    ezero = [0] * (n + 2)
    ezero[n + 1] = HUGE
    nelim[n + 1] = n + 1 + nemin
    mark = [False] * (n + 2)
    chead = [-1] * (n + 2)
    cnext = [-1] * (n + 2)
    for i in range(realn, 0, -1):
        j = parent[i]
        cnext[i] = chead[j]
        chead[j] = i
    for par in range(1, n + 2):
        child = []
        node = chead[par]
        while node != -1:
            child.append(node)
            node = cnext[node]
        child = _sort_by_val(child, cc)
        for node in child:
            merge = False
            if ezero[par] != HUGE:
                merge = ((cc[par] == cc[node] - 1) and (nelim[par] == 1)) or \
                        ((nelim[par] < nemin) and (nelim[node] < nemin))
            if merge:
                vnext[node] = vhead[par]
                vhead[par] = node
                ezero[par] = ezero[par] + ezero[node] + (cc[par] - 1 + nelim[par] - cc[node] + 1) * nelim[par]
                nelim[par] += nelim[node]
                nvert[par] += nvert[node]
                mark[node] = False
            else:
                mark[node] = True
    sperm = [0] * (n + 2)
    sptr = [0] * (n + 3)
    npar = [0] * (n + 3)
    scc = [0] * (n + 2)
    mp = [0] * (n + 2)
    v = 1
    nnodes = 0
    for node in range(1, realn + 1):
        if not mark[node]:
            continue
        nnodes += 1
        sptr[nnodes] = v
        npar[nnodes] = parent[node]
        scc[nnodes] = cc[node] + nelim[node] - 1
        v += nvert[node]
        k = v
        stack = [node]
        while stack:
            i = stack.pop()
            k -= 1
            sperm[i] = k
            mp[i] = nnodes
            if vnext[i] != -1:
                stack.append(vnext[i])
            if vhead[i] != -1:
                stack.append(vhead[i])
    sptr[nnodes + 1] = v
    mp[n + 1] = nnodes + 1
    npar[nnodes + 1] = n + 1
    for i in range(realn + 1, n + 1):
        sperm[i] = i
    sparent = [0] * (nnodes + 2)
    for node in range(1, nnodes + 1):
        sparent[node] = mp[npar[node]]
    return sperm, nnodes, sptr, sparent, scc


def apply_perm(n, perm, order, invp, cc):
    tmp = cc[:]
    for i in range(1, n + 1):
        cc[perm[i]] = tmp[i]
    tmp = invp[:]
    for i in range(1, n + 1):
        invp[perm[i]] = tmp[i]
    for i in range(1, n + 1):
        order[invp[i]] = i


def find_row_lists(n, ptr, row, perm, invp, nnodes, sptr, sparent, scc):
    seen = [0] * (n + 2)
    chead = [-1] * (nnodes + 2)
    cnext = [-1] * (nnodes + 2)
    for node in range(nnodes, 0, -1):
        i = sparent[node]
        cnext[node] = chead[i]
        chead[i] = node
    rptr = [0] * (nnodes + 2)
    rlist = [0] * (sum(scc[1:nnodes + 1]) + 1)
    rptr[1] = 1
    for node in range(1, nnodes + 1):
        rptr[node + 1] = rptr[node] + scc[node]
        idx = rptr[node]
        for piv in range(sptr[node], sptr[node + 1]):
            seen[piv] = node
            rlist[idx] = piv
            idx += 1
        child = chead[node]
        while child != -1:
            for i in range(rptr[child], rptr[child + 1]):
                j = rlist[i]
                if j < sptr[node] or seen[j] == node:
                    continue
                seen[j] = node
                rlist[idx] = j
                idx += 1
            child = cnext[child]
        for piv in range(sptr[node], sptr[node + 1]):
            col = invp[piv]
            for i in range(ptr[col], ptr[col + 1]):
                j = perm[row[i]]
                if j < piv or seen[j] == node:
                    continue
                seen[j] = node
                rlist[idx] = j
                idx += 1
    return rptr, rlist


def dbl_tr_sort(n, nnodes, rptr, rlist):
    cnt = [0] * (n + 3)
    for node in range(1, nnodes + 1):
        for ii in range(rptr[node], rptr[node + 1]):
            cnt[rlist[ii] + 2] += 1
    cnt[1] = cnt[2] = 1
    for i in range(1, n + 1):
        cnt[i + 2] = cnt[i + 1] + cnt[i + 2]
    col = [0] * (cnt[n + 2])
    for node in range(1, nnodes + 1):
        for ii in range(rptr[node], rptr[node + 1]):
            j = rlist[ii]
            col[cnt[j + 1]] = node
            cnt[j + 1] += 1
    nptr = rptr[:]
    for i in range(1, n + 1):
        for jj in range(cnt[i], cnt[i + 1]):
            node = col[jj]
            rlist[nptr[node]] = i
            nptr[node] += 1


def build_map(n, ptr, row, perm, invp, nnodes, sptr, rptr, rlist):
    nz = ptr[n + 1] - 1
    ptr2 = [0] * (n + 4)
    row2 = [0] * (nz + 1)
    origin = [0] * (nz + 1)
    for i in range(1, n + 1):
        for jj in range(ptr[i], ptr[i + 1]):
            k = row[jj]
            if k == i:
                continue
            ptr2[k + 2] += 1
    ptr2[1] = ptr2[2] = 1
    for i in range(1, n + 1):
        ptr2[i + 2] += ptr2[i + 1]
    for i in range(1, n + 1):
        for jj in range(ptr[i], ptr[i + 1]):
            k = row[jj]
            if k == i:
                continue
            row2[ptr2[k + 1]] = i
            origin[ptr2[k + 1]] = jj
            ptr2[k + 1] += 1
    nptr = [0] * (nnodes + 2)
    nlist = []
    mp = [0] * (n + 2)
    pp = 1
    for node in range(1, nnodes + 1):
        blkm = rptr[node + 1] - rptr[node]
        nptr[node] = pp
        for jj in range(rptr[node], rptr[node + 1]):
            mp[rlist[jj]] = jj - rptr[node] + 1
        for j in range(sptr[node], sptr[node + 1]):
            col = invp[j]
            for i in range(ptr2[col], ptr2[col + 1]):
                k = abs(perm[row2[i]])
                if k < j:
                    continue
                nlist.append(origin[i])
                nlist.append((j - sptr[node]) * blkm + mp[k])
                pp += 1
        for j in range(sptr[node], sptr[node + 1]):
            col = invp[j]
            for ii in range(ptr[col], ptr[col + 1]):
                k = abs(perm[row[ii]])
                if k < j:
                    continue
                nlist.append(ii)
                nlist.append((j - sptr[node]) * blkm + mp[k])
                pp += 1
    nptr[nnodes + 1] = pp
    return nptr, nlist


def analyse(n, ptr0, row0, order0, nemin=32):
    """ptr0/row0/order0: numpy arrays with 1-based values (C layout, index 0 = first).
    Returns a dict with the same keys/arrays as sylver_b200.Solver.symbolic()."""
    ptr = [0] + [int(x) for x in ptr0]
    row = [0] + [int(x) for x in row0]
    perm = [0] + [abs(int(x)) for x in order0] + [0]
    invp = [0] * (n + 2)
    for i in range(1, n + 1):
        invp[perm[i]] = i
    aptr, arow = expand_pattern(n, ptr, row)
    parent = find_etree(n, aptr, arow, perm, invp)
    realn = find_postorder(n, aptr, perm, invp, parent)
    cc = find_col_counts(n, aptr, arow, perm, invp, parent)
    sperm, nnodes, sptr, sparent, scc = find_supernodes(n, realn, parent, cc, nemin)
    apply_perm(n, sperm, perm, invp, cc)
    rptr, rlist = find_row_lists(n, aptr, arow, perm, invp, nnodes, sptr, sparent, scc)
    nfact = nflops = 0
    for node in range(1, nnodes + 1):
        ne = sptr[node + 1] - sptr[node]
        m = scc[node] - ne
        nfact += (ne * (ne + 1)) // 2 + ne * m
        for j in range(1, ne + 1):
            nflops += (m + j) ** 2
    dbl_tr_sort(n, nnodes, rptr, rlist)
    for i in range(1, n + 1):
        invp[perm[i]] = i
    for j in range(sptr[nnodes + 1], n + 1):
        perm[invp[j]] = 0
    nptr, nlist = build_map(n, ptr, row, perm, invp, nnodes, sptr, rptr, rlist)
    return dict(
        n=n, nnodes=nnodes,
        sptr=np.array(sptr[1:nnodes + 2], dtype=np.int32),
        sparent=np.array(sparent[1:nnodes + 1], dtype=np.int32),
        rptr=np.array(rptr[1:nnodes + 2], dtype=np.int64),
        rlist=np.array(rlist[1:rptr[nnodes + 1]], dtype=np.int32),
        nptr=np.array(nptr[1:nnodes + 2], dtype=np.int64),
        nlist=np.array(nlist, dtype=np.int64),
        order=np.array(perm[1:n + 1], dtype=np.int32),
        invp=np.array(invp[1:n + 1], dtype=np.int32),
        num_factor=nfact, num_flops=nflops,
        parent_etree=parent, cc=cc,
    )


def cmap(sym):
    """Per-edge assembly maps by definition: position of each child contribution
    row in the parent's row list (what src/assemble.hxx:152-302 builds through the
    per-front `map` vector).  Returns (cptr, cmap) 0-based."""
    nn = sym["nnodes"]
    sptr, sparent, rptr, rlist = sym["sptr"], sym["sparent"], sym["rptr"], sym["rlist"]
    cptr = [0]
    out = []
    for c in range(nn):
        ncol = sptr[c + 1] - sptr[c]
        rows = rlist[rptr[c] - 1 + ncol: rptr[c + 1] - 1]
        p = sparent[c] - 1
        if p >= nn:
            out.extend([-1] * len(rows))
        else:
            prow = {int(r): i for i, r in enumerate(rlist[rptr[p] - 1: rptr[p + 1] - 1])}
            out.extend(prow[int(r)] for r in rows)
        cptr.append(len(out))
    return np.array(cptr, dtype=np.int64), np.array(out, dtype=np.int32)
