"""Scalings computed at factorize, host code, no GPU needed:
options.scaling >= 4 (norm equilibration) and == 2 (auction matching).  The product's C++
(csrc/scaling.cpp) is compared bit for bit with the restatements of SPRAL's Fortran in
oracle/scaling.py, and both are held to the properties the reference's own tests check
(spral/tests/scaling.f90: test_auction_sym_random :46-184, test_equilib_sym_random :400-457)."""
import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen
from oracle import scaling as oscal


def badly_scaled(kind, k, seed):
    n, ptr, row, val = (gen.laplacian_7pt if kind == "lap7" else gen.laplacian_27pt)(k) if kind != "kkt" \
        else gen.stokes_kkt(k)
    rng = np.random.default_rng(seed)
    d = 10.0 ** rng.uniform(-4, 4, n)
    col = np.repeat(np.arange(n), np.diff(ptr))
    return n, ptr, row, val * d[row - 1] * d[col]


def row_inf_norms(n, ptr, row, val, s):
    col = np.repeat(np.arange(n), np.diff(ptr))
    a = np.abs(s[row - 1] * val * s[col])
    mx = np.zeros(n)
    np.maximum.at(mx, row - 1, a)
    np.maximum.at(mx, col, a)
    return mx


@pytest.mark.parametrize("kind,k,seed", [("lap7", 8, 1), ("lap27", 6, 2), ("kkt", 5, 3), ("lap7", 14, 4)])
def test_equilib_matches_restatement_bitwise(lib, kind, k, seed):
    n, ptr, row, val = badly_scaled(kind, k, seed)
    s, it = sb.equilib_scale(n, ptr, row, val)
    so, ito = oscal.inf_norm_equilib_sym(n, ptr, row, val)
    assert it == ito
    assert np.array_equal(s, so)                      # same operations in the same order
    mx = row_inf_norms(n, ptr, row, val, s)
    if it < 10:                                       # converged: |S A S| rows have inf-norm 1
        assert np.abs(1 - mx).max() < 1e-6
    assert mx.max() < 10 and mx.min() > 0.1           # and is well equilibrated after 10 sweeps anyway


def test_equilib_edge_cases(lib):
    # empty matrix, a matrix with an empty column (scaling stays 1 there), a diagonal matrix
    s, it = sb.equilib_scale(0, np.array([1], dtype=np.int64), np.zeros(0, np.int32), np.zeros(0))
    assert len(s) == 0
    ptr = np.array([1, 3, 3, 4], dtype=np.int64)      # column 2 empty
    row = np.array([1, 3, 3], dtype=np.int32)
    val = np.array([4.0, -2.0, 9.0])
    s, it = sb.equilib_scale(3, ptr, row, val)
    so, ito = oscal.inf_norm_equilib_sym(3, ptr, row, val)
    assert np.array_equal(s, so) and it == ito and s[1] == 1.0
    ptr = np.arange(1, 6, dtype=np.int64)
    row = np.arange(1, 5, dtype=np.int32)
    val = np.array([1e-8, 4.0, 1e6, 0.25])
    s, it = sb.equilib_scale(4, ptr, row, val)
    assert np.allclose(s * s * val, 1.0, rtol=1e-12) and it <= 2
    so, ito = oscal.inf_norm_equilib_sym(4, ptr, row, val)
    assert np.array_equal(s, so) and it == ito


def test_scaling_options_are_rejected_or_accepted(lib):
    """3 (the scaling saved by a matching-based ordering at analyse) cannot exist here, orderings
    being inputs: the reference's own error for that case, without a GPU being touched."""
    n, ptr, row, val = gen.laplacian_7pt(4)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, gen.nested_dissection_order(4)).flag == 0
    s.options.scaling = 3
    assert s.factorize(val, posdef=True).flag == -15       # SYLVER_ERROR_NO_SAVED_SCALING
    s.free()


def random_sym(n, nza, rng, wide=False):
    """Random sparse symmetric matrix, lower triangle CSC 1-based, every column non-empty, about
    nza entries: uniform in (-1, 1) like the reference test's gen_random_sym
    (spral/src/random_matrix.f90), or with magnitudes over twelve orders (wide)."""
    cols = []
    ptr = [1]
    rows_all, vals_all = [], []
    extra = max(nza - n, 0)
    per_col = rng.multinomial(extra, np.ones(n) / n) if n else []
    for j in range(n):
        cand = np.arange(j + 1, n)
        k = min(int(per_col[j]), len(cand))
        r = np.sort(rng.choice(cand, size=k, replace=False)) if k else np.zeros(0, dtype=np.int64)
        r = np.concatenate([[j], r])
        rows_all.append(r + 1)
        if wide:
            vals_all.append(rng.choice([-1.0, 1.0], len(r)) * 10.0 ** rng.uniform(-6, 6, len(r)))
        else:
            v = rng.uniform(-1.0, 1.0, len(r))
            vals_all.append(np.where(v == 0.0, 0.5, v))
        ptr.append(ptr[-1] + len(r))
    return (np.array(ptr, dtype=np.int64), np.concatenate(rows_all).astype(np.int32) if n else np.zeros(0, np.int32),
            np.concatenate(vals_all) if n else np.zeros(0))


def check_auction_properties(n, ptr, row, val, scaling, match, scaled_entries=True):
    # spral/tests/scaling.f90:113-152: matching is valid, injective, covers >= 90 %
    assert ((match >= 0) & (match <= n)).all()
    ent = set()
    for j in range(n):
        for k in range(ptr[j] - 1, ptr[j + 1] - 1):
            ent.add((int(row[k]), j + 1))
            ent.add((j + 1, int(row[k])))
    nz = match[match != 0]
    assert len(np.unique(nz)) == len(nz)
    for i in range(n):
        if match[i]:
            assert (i + 1, int(match[i])) in ent
    assert len(nz) >= 0.9 * n
    if not scaled_entries:
        return
    # :157-179: every scaled entry < 2, every row has one >= 0.75
    mx = row_inf_norms(n, ptr, row, val, scaling)
    col = np.repeat(np.arange(n), np.diff(ptr))
    assert (np.abs(scaling[row - 1] * val * scaling[col]) < 2.0).all()
    assert (mx >= 0.75).all()


@pytest.mark.parametrize("seed", range(12))
def test_auction_matches_restatement_and_reference_properties(lib, seed):
    rng = np.random.default_rng(100 + seed)
    n = seed + 1 if seed < 5 else int(rng.integers(20, 160))         # very small problems first, as the reference
    nza = n + int(rng.integers(0, max(n * n // 2 - n, 0) + 1)) // (1 if n < 30 else 8)
    ptr, row, val = random_sym(n, nza, rng)
    s, m, inf = sb.auction_scale(n, ptr, row, val)
    so, mo, info = oscal.auction_scale_sym(n, ptr, row, val)
    assert inf == info
    assert np.array_equal(m, mo)
    assert np.array_equal(s, so)
    check_auction_properties(n, ptr, row, val, s, m)
    # entries spread over twelve orders of magnitude: same code path, bitwise agreement; the
    # early termination (90 % matched) leaves a few rows further from 1 than the reference's
    # test family, so only the matching is checked
    ptr, row, val = random_sym(n, nza, rng, wide=True)
    s, m, inf = sb.auction_scale(n, ptr, row, val)
    so, mo, info = oscal.auction_scale_sym(n, ptr, row, val)
    assert inf == info and np.array_equal(m, mo) and np.array_equal(s, so)
    assert len(np.unique(m[m != 0])) == int((m != 0).sum())


@pytest.mark.parametrize("kind,k,seed", [("lap7", 7, 1), ("lap27", 5, 2), ("kkt", 5, 3)])
def test_auction_on_the_benchmark_families(lib, kind, k, seed):
    n, ptr, row, val = badly_scaled(kind, k, seed)
    s, m, inf = sb.auction_scale(n, ptr, row, val)
    so, mo, info = oscal.auction_scale_sym(n, ptr, row, val)
    assert inf == info and np.array_equal(m, mo) and np.array_equal(s, so)
    # rows and columns scaled by 10^U(-4,4): the few rows the early termination leaves
    # unmatched end further from 1 than on the reference's test family
    check_auction_properties(n, ptr, row, val, s, m, scaled_entries=False)
    mx = row_inf_norms(n, ptr, row, val, s)
    assert np.median(mx) > 0.9 and mx.max() < 2.0
    # an explicit zero is dropped before the matching (scaling.f90:1546), not treated as an entry
    val0 = val.copy()
    off = int(np.nonzero(row != np.repeat(np.arange(n), np.diff(ptr)) + 1)[0][0])
    val0[off] = 0.0
    s0, m0, _ = sb.auction_scale(n, ptr, row, val0)
    so0, mo0, _ = oscal.auction_scale_sym(n, ptr, row, val0)
    assert np.array_equal(s0, so0) and np.array_equal(m0, mo0)


@pytest.mark.parametrize("seed", range(6))
def test_equilib_reference_property_on_random_matrices(lib, seed):
    # spral/tests/scaling.f90:434-452: the infinity norm of every scaled row is within 0.05 of 1
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(5, 300))
    ptr, row, val = random_sym(n, 4 * n, rng, wide=bool(seed % 2))
    s, it = sb.equilib_scale(n, ptr, row, val)
    so, ito = oscal.inf_norm_equilib_sym(n, ptr, row, val)
    assert np.array_equal(s, so) and it == ito
    assert (1.0 - row_inf_norms(n, ptr, row, val, s) <= 0.05).all()


ERR_TOL = 5e-14          # spral/tests/scaling.f90:14


def check_hungarian_properties(n, ptr, row, val, scaling, match):
    # spral/tests/scaling.f90:613-640: a perfect matching on entries of the matrix
    assert ((match >= 1) & (match <= n)).all()
    assert np.array_equal(np.sort(match), np.arange(1, n + 1))
    ent = set()
    for j in range(n):
        for k in range(ptr[j] - 1, ptr[j + 1] - 1):
            ent.add((int(row[k]), j + 1))
            ent.add((j + 1, int(row[k])))
    assert all((i + 1, int(match[i])) in ent for i in range(n))
    # :645-662: every scaled entry <= 1, every row attains 1 (the optimality of the matching)
    col = np.repeat(np.arange(n), np.diff(ptr))
    assert (np.abs(scaling[row - 1] * val * scaling[col]) < 1.0 + ERR_TOL).all()
    assert (row_inf_norms(n, ptr, row, val, scaling) >= 1.0 - ERR_TOL).all()


@pytest.mark.parametrize("seed", range(14))
def test_hungarian_matches_restatement_and_reference_properties(lib, seed):
    rng = np.random.default_rng(300 + seed)
    n = seed + 1 if seed < 6 else int(rng.integers(20, 200))         # very small problems first, as the reference
    nza = n + int(rng.integers(0, max(n * n // 10 - n, 0) + 1))
    ptr, row, val = random_sym(n, nza, rng, wide=bool(seed % 2))
    s, m, inf = sb.hungarian_scale(n, ptr, row, val)
    so, mo, info = oscal.hungarian_scale_sym(n, ptr, row, val)
    assert inf == info and inf["flag"] == 0 and inf["matched"] == n
    assert np.array_equal(m, mo)
    assert np.array_equal(s, so)
    check_hungarian_properties(n, ptr, row, val, s, m)


@pytest.mark.parametrize("kind,k,seed", [("lap7", 8, 1), ("lap27", 6, 2), ("kkt", 6, 3)])
def test_hungarian_on_the_benchmark_families(lib, kind, k, seed):
    n, ptr, row, val = badly_scaled(kind, k, seed)
    s, m, inf = sb.hungarian_scale(n, ptr, row, val)
    so, mo, info = oscal.hungarian_scale_sym(n, ptr, row, val)
    assert inf == info and np.array_equal(m, mo) and np.array_equal(s, so)
    check_hungarian_properties(n, ptr, row, val, s, m)


def structurally_singular(n, rng, nempty):
    """Random symmetric matrix whose last `nempty` rows/columns are empty, and whose first
    diagonal entries are missing (off-diagonal matching needed)."""
    ptr, row, val = random_sym(n - nempty, 3 * (n - nempty), rng)
    ptr = np.concatenate([ptr, np.full(nempty, ptr[-1], dtype=np.int64)])
    return ptr, row, val


@pytest.mark.parametrize("seed", range(6))
def test_hungarian_structurally_singular(lib, seed):
    """The rank-deficient branch of hungarian_wrapper (scaling.f90:688-800): matching on the
    full-rank part, Duff-Pralet scaling of the rest; flag 1 with scale_if_singular (what
    options.action = true selects), -2 without."""
    rng = np.random.default_rng(400 + seed)
    n = int(rng.integers(8, 80))
    nempty = int(rng.integers(1, 4))
    ptr, row, val = structurally_singular(n, rng, nempty)
    s, m, inf = sb.hungarian_scale(n, ptr, row, val, scale_if_singular=True)
    so, mo, info = oscal.hungarian_scale_sym(n, ptr, row, val, scale_if_singular=True)
    assert inf == info and inf["flag"] == 1 and inf["matched"] == n - nempty
    assert np.array_equal(m, mo) and np.array_equal(s, so)
    assert (m[n - nempty:] < 0).all() and (m[: n - nempty] >= 1).all()
    assert np.array_equal(s[n - nempty:], np.ones(nempty))          # isolated rows: scaling 1
    # the matched part satisfies the same optimality property
    k = n - nempty
    check_hungarian_properties(k, ptr[: k + 1], row, val, s[:k], m[:k])
    s2, m2, inf2 = sb.hungarian_scale(n, ptr, row, val, scale_if_singular=False)
    assert inf2["flag"] == -2
    assert oscal.hungarian_scale_sym(n, ptr, row, val, scale_if_singular=False)[2]["flag"] == -2


def test_hungarian_rank_deficient_with_entries(lib):
    """Structurally singular although no row is empty: three rows that only see one column.
    Arrow-like pattern: rows 2,3,4 have entries only in column 1 (and no diagonal)."""
    #     [ 2  1  1  1 ]
    # A = [ 1  0  0  0 ]    structural rank 2
    #     [ 1  0  0  0 ]
    #     [ 1  0  0  0 ]
    ptr = np.array([1, 5, 5, 5, 5], dtype=np.int64)
    row = np.array([1, 2, 3, 4], dtype=np.int32)
    val = np.array([2.0, 1.0, 4.0, 0.5])
    s, m, inf = sb.hungarian_scale(4, ptr, row, val, scale_if_singular=True)
    so, mo, info = oscal.hungarian_scale_sym(4, ptr, row, val, scale_if_singular=True)
    assert inf == info and inf["flag"] == 1
    assert np.array_equal(m, mo) and np.array_equal(s, so)
    assert np.isfinite(s).all() and (s > 0).all()
    assert (m < 0).sum() == 2


def test_committed_golden_vectors(lib):
    """tests/golden/preprocess.json (written by tests/golden/make_golden.py from the oracle
    restatements): the product reproduces the committed scalings and matchings bit for bit."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "preprocess.json")))
    assert len(g["scaling"]) >= 2
    for rec in g["scaling"]:
        n, ptr, row, val = badly_scaled(rec["kind"], rec["k"], rec["seed"])
        s, it = sb.equilib_scale(n, ptr, row, val)
        assert [float(x).hex() for x in s] == rec["equilib"]["scaling"] and it == rec["equilib"]["iterations"]
        s, m, inf = sb.auction_scale(n, ptr, row, val)
        assert [float(x).hex() for x in s] == rec["auction"]["scaling"]
        assert m.tolist() == rec["auction"]["match"] and inf == rec["auction"]["inform"]
        s, m, inf = sb.hungarian_scale(n, ptr, row, val)
        assert [float(x).hex() for x in s] == rec["hungarian"]["scaling"]
        assert m.tolist() == rec["hungarian"]["match"] and inf == rec["hungarian"]["inform"]
