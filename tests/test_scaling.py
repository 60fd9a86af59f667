"""options.scaling >= 4 (norm equilibration): the product's C++ against the
numpy restatement of SPRAL's inf_norm_equilib_sym (oracle/scaling.py), bit for bit, and against
the property the iteration converges to.  No GPU needed."""
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
    """1..3 (MC64, auction, saved matching scaling) stay outside this path: flag -98 without a
    GPU being touched; >= 4 needs the structure arrays."""
    n, ptr, row, val = gen.laplacian_7pt(4)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, gen.nested_dissection_order(4)).flag == 0
    for sc in (1, 2, 3):
        s.options.scaling = sc
        assert s.factorize(val, posdef=True).flag == -98
    s.free()
