"""The reference lineage's own scaling tests, replayed on exactly their inputs.

spral/tests/scaling.f90 draws 100 random symmetric matrices per routine from SPRAL's
generator with its default seed (test_auction_sym_random :46-184, test_equilib_sym_random
:370-457, test_hungarian_sym_random :576-704).  oracle/spral_random.py restates that generator
(spral/src/random.f90, random_matrix.f90, gen_random_sym), so the product's C++ scalings are run
on the same 100 matrices and held to the same checks with the same tolerances; on the small
ones the Python restatements of the Fortran must also agree bit for bit.  No GPU needed."""
import numpy as np
import pytest

import sylver_b200 as sb
from oracle import scaling as oscal, spral_random as sr

MAXN, MAXNZ, NPROB = 1000, 1000000, 100
ERR_TOL = 5e-14


def problems(divisor):
    """The (n, ptr, row, val) sequence of one test routine: nza = n + random(n^2/divisor - n)."""
    state = sr.RandomState()                   # a fresh default-initialised random_state per routine
    out = []
    for prblm in range(1, NPROB + 1):
        n = state.integer(MAXN)
        if prblm < 21:
            n = prblm                          # check very small problems
        i = max(0, n * n // divisor - n)
        nza = n + state.integer(i)
        if nza > MAXNZ or n > MAXN:
            continue
        out.append((n,) + sr.gen_random_sym(state, n, nza))
    return out


@pytest.fixture(scope="module")
def half_family():
    return problems(2)                          # auction and equilibration tests


@pytest.fixture(scope="module")
def tenth_family():
    return problems(10)                         # Hungarian test


def scaled_abs(n, ptr, row, val, s):
    col = np.repeat(np.arange(n), np.diff(ptr))
    v = np.abs(s[col] * val * s[row - 1])
    rmax = np.zeros(n)
    np.maximum.at(rmax, row - 1, v)
    np.maximum.at(rmax, col, v)
    return v, rmax


def entry_set(n, ptr, row):
    col = np.repeat(np.arange(1, n + 1), np.diff(ptr))
    return set(zip(row.tolist(), col.tolist())) | set(zip(col.tolist(), row.tolist()))


def test_generator_first_problems(half_family):
    """Anchor of the restated generator: sizes are the deterministic stream of the LCG."""
    assert len(half_family) == NPROB
    assert [p[0] for p in half_family[:20]] == list(range(1, 21))
    n, ptr, row, val = half_family[20]
    assert (n, int(ptr[n] - 1)) == (573, 26506)
    for n, ptr, row, val in half_family[:40]:
        col = np.repeat(np.arange(1, n + 1), np.diff(ptr))
        assert (row >= col).all() and (row <= n).all()                      # lower triangle
        assert (row[ptr[:-1] - 1] == np.arange(1, n + 1)).all()             # diagonal first in every column
        assert all((np.diff(row[ptr[j] - 1: ptr[j + 1] - 1]) > 0).all() for j in range(n))
        assert (np.abs(val) <= 1000.0).all()


def test_auction_sym_random(lib, half_family):
    for n, ptr, row, val in half_family:
        s, match, inf = sb.auction_scale(n, ptr, row, val)
        assert inf["flag"] >= 0
        assert ((match >= 0) & (match <= n)).all()                           # :113-122
        nz = match[match != 0]
        ent = entry_set(n, ptr, row)
        assert all((i + 1, int(match[i])) in ent for i in range(n) if match[i])   # :126-137
        assert len(nz) >= 0.9 * n                                            # :140
        assert len(np.unique(nz)) == len(nz)                                 # :146
        v, rmax = scaled_abs(n, ptr, row, val, s)
        assert (v < 2.0).all()                                               # :161
        assert (rmax >= 0.75).all()                                          # :172
        if ptr[n] - 1 < 3000:
            so, mo, info = oscal.auction_scale_sym(n, ptr, row, val)
            assert inf == info and np.array_equal(match, mo) and np.array_equal(s, so)


def test_equilib_sym_random(lib, half_family):
    for n, ptr, row, val in half_family:
        s, it = sb.equilib_scale(n, ptr, row, val)
        v, rinf = scaled_abs(n, ptr, row, val, s)
        assert (1.0 - rinf <= 0.05).all()                                    # :446
        if ptr[n] - 1 < 20000:
            so, ito = oscal.inf_norm_equilib_sym(n, ptr, row, val)
            assert it == ito and np.array_equal(s, so)


def test_hungarian_sym_random(lib, tenth_family):
    assert len(tenth_family) == NPROB
    for n, ptr, row, val in tenth_family:
        s, match, inf = sb.hungarian_scale(n, ptr, row, val)
        assert inf["flag"] >= 0
        assert ((match >= 1) & (match <= n)).all()                           # :613-619
        ent = entry_set(n, ptr, row)
        assert all((i + 1, int(match[i])) in ent for i in range(n))          # :620-633
        assert np.array_equal(np.sort(match), np.arange(1, n + 1))           # :636
        v, rmax = scaled_abs(n, ptr, row, val, s)
        assert (v < 1.0 + ERR_TOL).all()                                     # :651
        assert (rmax >= 1.0 - ERR_TOL).all()                                 # :663
        if ptr[n] - 1 < 3000:
            so, mo, info = oscal.hungarian_scale_sym(n, ptr, row, val)
            assert inf == info and np.array_equal(match, mo) and np.array_equal(s, so)
