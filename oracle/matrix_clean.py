"""TEST INFRASTRUCTURE (oracle): restatement of SPRAL's clean_cscl_oop for a real symmetric
indefinite matrix with a conversion map, as the reference's analyse(check=.true.) calls it
(/root/reference/src/spldlt_analyse_mod.F90:707-739 ->
 /root/reference/spral/src/matrix_util.f90:1024-1398, heap sort :2729-2777 + pushdown64,
 apply_conversion_map :2559-2606).

Only tests/ may import this.  The Fortran cannot be compiled here -> PARITY UNPINNED against a
run of the reference; pinned by (i) the flags the reference's own tests expect for its fixtures
(/root/reference/spral/tests/ssids/ssids.f90:140-175,870-1002, restated in
tests/test_matrix_clean.py) and (ii) the definition: the cleaned matrix is the input with
out-of-range entries dropped and duplicates summed.  Pure-Python loops over 1-based lists.
"""
from __future__ import annotations

import numpy as np


def _pushdown(root, last, array, mp, base):
    """array/mp are whole lists; the section being sorted starts at base + 1 (1-based root/last)."""
    root_idx = array[base + root]
    root_map = mp[base + root]
    insert = root
    test = 2 * insert
    while test <= last:
        if test != last:
            if array[base + test + 1] > array[base + test]:
                test += 1
        if array[base + test] <= root_idx:
            break
        array[base + insert] = array[base + test]
        mp[base + insert] = mp[base + test]
        insert = test
        test = 2 * insert
    array[base + insert] = root_idx
    mp[base + insert] = root_map


def _heap_sort(array, mp, base, n):                        # sort64, matrix_util.f90:2729-2777
    if n <= 1:
        return
    for root in range(n // 2, 0, -1):
        _pushdown(root, n, array, mp, base)
    for i in range(n, 1, -1):
        array[base + 1], array[base + i] = array[base + i], array[base + 1]
        mp[base + 1], mp[base + i] = mp[base + i], mp[base + 1]
        _pushdown(1, i - 1, array, mp, base)


def clean_cscl_oop_sym_indef(n: int, ptr, row):
    """Returns dict(flag, noor, ndup, ptr, row, map) -- ptr/row/map as 1-based numpy arrays
    (map: ne source indices, then (dest, src) pairs) -- or dict(flag<0)."""
    if n < 0:
        return dict(flag=-3)
    ptr_in = [0] + [int(x) for x in ptr[: n + 1]]
    row_in = [0] + [int(x) for x in row]
    if ptr_in[1] < 1:
        return dict(flag=-5)                               # ERROR_PTR_1 :1057
    m = n
    nin = max(ptr_in[n + 1] - 1, 0)
    ptr_out = [0] * (n + 2)
    row_out = [0] * (nin + 1)
    mp = [0] * (2 * nin + 2)
    dups = []                                              # linked list, head insertion
    idup = ioor = idiag = 0
    k = 1
    for col in range(1, n + 1):                            # :1092
        ptr_out[col] = k
        if ptr_in[col + 1] < ptr_in[col]:
            return dict(flag=-6)                           # ERROR_PTR_MONO
        minidx = col                                       # abs(matrix_type) >= SYM_PSDEF
        for i in range(ptr_in[col], ptr_in[col + 1]):
            j = row_in[i]
            if j < minidx or j > m:
                ioor += 1
                continue
            row_out[k] = j
            mp[k] = i                                      # multiplier = 1 (CSC)
            k += 1
        cnt = k - ptr_out[col]
        if cnt == 0 and ptr_in[col + 1] - ptr_in[col] != 0:
            return dict(flag=-10)                          # ERROR_ALL_OOR
        if cnt != 0:
            _heap_sort(row_out, mp, ptr_out[col] - 1, cnt)
            last = k - 1
            k = ptr_out[col] + 1
            if row_out[ptr_out[col]] == col:
                idiag += 1
            for i in range(ptr_out[col] + 1, last + 1):
                if row_out[i] == row_out[i - 1]:
                    idup += 1
                    dups.insert(0, (mp[i], k - 1))         # (src, dest), new head
                    continue
                if row_out[i] == col:
                    idiag += 1
                row_out[k] = row_out[i]
                mp[k] = mp[i]
                k += 1
    ptr_out[n + 1] = k
    lmap = k - 1
    for src, dest in dups:                                 # :1336-1344 (idup counted again)
        idup += 1
        mp[lmap + 1] = dest
        mp[lmap + 2] = src
        lmap += 2
    flag = 0
    if ioor > 0 or idup > 0 or idiag < n:                  # :1372-1387
        if ioor > 0:
            flag = 1
        if idup > 0:
            flag = 2
        if idup > 0 and ioor > 0:
            flag = 3
        if idiag < n and ioor > 0:
            flag = 5
        elif idiag < n and idup > 0:
            flag = 5
        elif idiag < n:
            flag = 4
    ne = k - 1
    return dict(flag=flag, noor=ioor, ndup=idup, ptr=np.array(ptr_out[1:], dtype=np.int64),
                row=np.array(row_out[1: ne + 1], dtype=np.int32), map=np.array(mp[1: lmap + 1], dtype=np.int64))


def apply_conversion_map(cm: dict, val) -> np.ndarray:     # matrix_util.f90:2577-2587
    ne = len(cm["row"])
    mp = cm["map"]
    out = [float(val[int(mp[i]) - 1]) for i in range(ne)]
    for i in range(ne, len(mp) - 1, 2):
        j, k = int(mp[i]), int(mp[i + 1])
        out[j - 1] = out[j - 1] + float(val[k - 1])
    return np.array(out)
