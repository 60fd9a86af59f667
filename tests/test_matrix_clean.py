"""spldlt_analyse(check=True): matrix cleaning (host code, no GPU needed).  The product's C++
(csrc/clean.cpp) against the restatement of SPRAL's clean_cscl_oop (oracle/matrix_clean.py), bit
for bit; both against the definition (out-of-range entries dropped, duplicates summed); and the
flags the reference's own tests expect for its fixtures (spral/tests/ssids/ssids.f90:140-175,
870-1002 -- SyLVER's analyse makes the same calls, src/spldlt_analyse_mod.F90:707-739)."""
import numpy as np
import pytest

import sylver_b200 as sb
from oracle import matrix_clean as oclean

# SYLVER_WARNING_* / SYLVER_ERROR_* (src/sylver_datatypes_mod.F90:13-45)
IDX_OOR, DUP_IDX, DUP_AND_OOR, MISSING_DIAG, MISS_DIAG_OORDUP = 1, 2, 3, 4, 5
ERR_A_N_OOR, ERR_A_PTR, ERR_A_ALL_OOR = -2, -3, -4


def simple_mat_lower():
    """4x4 symmetric matrix, lower triangle, diagonal present (the shape of the reference's
    simple_mat_lower fixture: a dense first column plus a few entries)."""
    ptr = np.array([1, 5, 7, 9, 10], dtype=np.int64)
    row = np.array([1, 2, 3, 4, 2, 3, 3, 4, 4], dtype=np.int32)
    val = np.array([10.0, 2.0, 1.0, 3.0, 10.0, 1.0, 10.0, 2.0, 10.0])
    return 4, ptr, row, val


def dense_lower(n, ptr, row, val):
    a = np.zeros((n, n))
    for j in range(n):
        for k in range(ptr[j] - 1, ptr[j + 1] - 1):
            i = row[k] - 1
            if j <= i < n:
                a[i, j] += val[k]
    return a


def dirty_random(n, rng):
    """Random matrix with duplicates, entries above the diagonal, rows outside 1..n, unsorted
    columns, some missing diagonals and some empty columns."""
    ptr = [1]
    rows, vals = [], []
    for j in range(n):
        r = []
        if rng.random() < 0.85:
            r.append(j + 1)
        nx = int(rng.integers(0, 6))
        if nx:
            r += list(rng.integers(-1, n + 3, nx))          # may be < j+1 (upper), 0, -1, > n
        if r and rng.random() < 0.5:
            r += list(rng.choice(r, size=int(rng.integers(1, 3))))      # duplicates
        valid = [x for x in r if j + 1 <= x <= n]
        if r and not valid:
            r.append(j + 1)                                 # never a column with ONLY out-of-range entries
        rng.shuffle(r)
        rows += r
        vals += list(rng.uniform(-1, 1, len(r)))
        ptr.append(ptr[-1] + len(r))
    return np.array(ptr, dtype=np.int64), np.array(rows, dtype=np.int32), np.array(vals)


@pytest.mark.parametrize("seed", range(20))
def test_clean_matches_restatement_and_definition(lib, seed):
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(1, 60))
    ptr, row, val = dirty_random(n, rng)
    c = sb.clean_matrix(n, ptr, row)
    o = oclean.clean_cscl_oop_sym_indef(n, ptr, row)
    assert c["flag"] == o["flag"] >= 0
    assert (c["noor"], c["ndup"]) == (o["noor"], o["ndup"])
    for key in ("ptr", "row", "map"):
        assert np.array_equal(c[key], o[key]), key
    # definition: strictly increasing rows in the lower triangle, same matrix after summing
    cp, cr = c["ptr"], c["row"]
    for j in range(n):
        seg = cr[cp[j] - 1: cp[j + 1] - 1]
        assert (np.diff(seg) > 0).all() and (seg >= j + 1).all() and (seg <= n).all()
    v2 = oclean.apply_conversion_map(c, val)
    assert np.array_equal(sb.apply_conversion_map(c, val), v2)        # duplicates summed in the same order
    assert np.allclose(dense_lower(n, cp, cr, v2), dense_lower(n, ptr, row, val), rtol=1e-14, atol=1e-15)
    # counts: out-of-range as defined; the reference counts every duplicate twice when it builds a map
    col = np.repeat(np.arange(1, n + 1), np.diff(ptr))
    oor = int(((row < col) | (row > n)).sum())
    assert c["noor"] == oor
    assert c["ndup"] == 2 * (len(row) - oor - len(cr))


def analyse_flag(n, ptr, row, check=True):
    s = sb.Solver()
    inf = s.analyse(n, ptr, row, np.arange(1, max(n, 0) + 1, dtype=np.int32), check=check)
    out = (inf.flag, inf.matrix_outrange, inf.matrix_dup)
    s.free()
    return out


def test_reference_error_fixtures(lib):
    """spral/tests/ssids/ssids.f90:140-175."""
    n, ptr, row, val = simple_mat_lower()
    assert analyse_flag(-1, ptr, row)[0] == ERR_A_N_OOR
    p = ptr.copy(); p[0] = 0
    assert analyse_flag(n, p, row)[0] == ERR_A_PTR                    # ptr with zero component
    p = ptr.copy(); p[1], p[2] = p[2], p[1]
    assert analyse_flag(n, p, row)[0] == ERR_A_PTR                    # non-monotonic ptr
    r = row.copy(); r[: ptr[1] - 1] = 0
    assert analyse_flag(n, ptr, r)[0] == ERR_A_ALL_OOR                # all of column 1 out of range


def test_reference_warning_fixtures(lib):
    """spral/tests/ssids/ssids.f90:870-1002: one extra entry appended to the last column."""
    n, ptr, row, val = simple_mat_lower()

    def appended(extra_rows):
        p = ptr.copy(); p[-1] += len(extra_rows)
        return p, np.concatenate([row, np.array(extra_rows, dtype=np.int32)])

    assert analyse_flag(n, *appended([-1])) == (IDX_OOR, 1, 0)         # out of range above
    assert analyse_flag(n, *appended([n + 1])) == (IDX_OOR, 1, 0)      # out of range below
    assert analyse_flag(n, *appended([n])) == (DUP_IDX, 0, 2)          # duplicate (counted twice, as the reference)
    assert analyse_flag(n, *appended([n + 1, n])) == (DUP_AND_OOR, 1, 2)
    # missing diagonal entry (indef): the reference's literal fixture (:953-958)
    p = np.array([1, 4, 5, 6, 7], dtype=np.int64)
    r = np.array([1, 2, 4, 2, 4, 4], dtype=np.int32)
    assert analyse_flag(4, p, r) == (MISSING_DIAG, 0, 0)
    # missing diagonal and out of range (:978-983)
    p = np.array([1, 4, 5, 6, 8], dtype=np.int64)
    r = np.array([1, 2, 4, 2, 4, 4, -1], dtype=np.int32)
    assert analyse_flag(4, p, r) == (MISS_DIAG_OORDUP, 1, 0)
    # a clean matrix raises nothing, with or without checking
    assert analyse_flag(n, ptr, row) == (0, 0, 0)
    assert analyse_flag(n, ptr, row, check=False) == (0, 0, 0)


def test_checked_analysis_equals_analysis_of_the_clean_matrix(lib):
    """The symbolic output of analyse(check=True) on a dirty matrix is bit for bit that of
    analyse(check=False) on its cleaned structure."""
    rng = np.random.default_rng(77)
    n = 40
    ptr, row, val = dirty_random(n, rng)
    c = sb.clean_matrix(n, ptr, row)
    order = rng.permutation(n).astype(np.int32) + 1
    s1, s2 = sb.Solver(), sb.Solver()
    i1 = s1.analyse(n, ptr, row, order, check=True)
    i2 = s2.analyse(n, c["ptr"], c["row"], order, check=False)
    assert i1.flag in (c["flag"], 6) and i2.flag in (0, 6)
    y1, y2 = s1.symbolic(), s2.symbolic()
    for key in ("sptr", "sparent", "rptr", "rlist", "nptr", "nlist", "order", "invp"):
        assert np.array_equal(y1[key], y2[key]), key
    assert (i1.num_factor, i1.num_flops) == (i2.num_factor, i2.num_flops)
    s1.free(); s2.free()


def test_committed_golden_vectors(lib):
    """tests/golden/preprocess.json: cleaned structure, conversion map and mapped values of
    fixed dirty matrices, as the oracle restatement produced them."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess.json")))
    assert len(g["clean"]) >= 2
    for rec in g["clean"]:
        ptr = np.array(rec["ptr_in"], dtype=np.int64)
        row = np.array(rec["row_in"], dtype=np.int32)
        val = np.array([float.fromhex(x) for x in rec["val_in"]])
        c = sb.clean_matrix(rec["n"], ptr, row)
        assert (c["flag"], c["noor"], c["ndup"]) == (rec["flag"], rec["noor"], rec["ndup"])
        assert c["ptr"].tolist() == rec["ptr"] and c["row"].tolist() == rec["row"] and c["map"].tolist() == rec["map"]
        assert [float(x).hex() for x in sb.apply_conversion_map(c, val)] == rec["val"]
