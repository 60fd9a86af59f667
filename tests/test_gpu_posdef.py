"""GPU parity tests, positive definite path (call through the C ABI)."""
import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

pytestmark = pytest.mark.gpu


def _bwderr_tree(k, stencil, oracle_ref, posdef=True):
    n, ptr, row, val = (gen.laplacian_7pt if stencil == 7 else gen.laplacian_27pt)(k)
    order = gen.nested_dissection_order(k)
    s = sb.Solver()
    inf = s.analyse(n, ptr, row, order)
    assert inf.flag == 0
    inf = s.factorize(val, posdef=posdef)
    assert inf.flag == 0, inf.flag
    x0 = np.ones(n)
    b = gen.sym_matvec(n, ptr, row, val, x0)
    x = s.solve(b)
    be = gen.backward_error(n, ptr, row, val, x, b)
    sym = s.symbolic()
    ot = oracle_ref.OracleTree(sym)
    ot.factor(val, posdef)
    xo = ot.solve_original(b)
    beo = gen.backward_error(n, ptr, row, val, xo, b)
    return be, beo, inf, ot.stats, s


def test_dmma_microbench(lib):
    sb.require_gpu()
    peak = lib.sylver_b200_bench_dmma(0, 0, 0, 3)
    fma = lib.sylver_b200_bench_dmma(2, 0, 0, 3)
    tile = lib.sylver_b200_bench_dmma(1, 8192, 256, 3)
    copy = lib.sylver_b200_bench_copy(1 << 30, 5)
    print(f"\nDMMA issue peak {peak:.2f} TF/s, DFMA peak {fma:.2f} TF/s, tile kernel {tile:.2f} TF/s, copy {copy:.0f} GB/s")
    assert peak > 1.0 and tile > 0.5


@pytest.mark.parametrize("m,n", [(64, 64), (200, 72), (500, 500), (1000, 300), (1301, 517), (2048, 1024)])
def test_dense_front_posdef(lib, oracle_ref, m, n):
    sb.require_gpu()
    a = gen.dense_posdef(m)
    lda = m + 3
    buf = np.zeros((lda, n), order="F")
    buf[:m, :] = np.tril(a)[:, :n]
    k = m - n
    contrib = np.zeros((max(k, 1), max(k, 1)), order="F")
    ret = lib.sylver_b200_factor_front_posdef(m, n, buf.ctypes.data, lda, contrib.ctypes.data, 128, None)
    assert ret == n
    Lo, Co, info = oracle_ref.factor_front_posdef(a, n)
    assert info == -1
    L = np.tril(buf[:m, :])
    Lo = np.tril(Lo)
    scale = np.abs(Lo).max()
    # tolerance: both are backward-stable Cholesky factors of the same SPD matrix
    assert np.abs(L - Lo).max() <= 1e-11 * scale
    if k > 0:
        # the engine's contribution is -L21 L21^T (A22 arrives by assembly)
        C = np.tril(contrib[:k, :k])
        L21 = Lo[n:, :]
        Cref = -np.tril(L21 @ L21.T)
        assert np.abs(C - Cref).max() <= 1e-11 * max(1.0, np.abs(Cref).max())


@pytest.mark.parametrize("k,stencil", [(8, 7), (20, 7), (12, 27), (30, 7), (24, 27)])
def test_laplacian_posdef_tree(lib, oracle_ref, k, stencil):
    sb.require_gpu()
    be, beo, inf, ostats, s = _bwderr_tree(k, stencil, oracle_ref)
    print(f"\nk={k} stencil={stencil} bwderr={be:.2e} oracle={beo:.2e} timings={s.timings()}")
    # north_star tolerance: <= 1e-14 and within 10x of the reference
    assert be <= 1e-14
    assert be <= 10 * max(beo, 1e-16)
    assert inf.num_neg == 0


def test_not_posdef_flag(lib):
    sb.require_gpu()
    n, ptr, row, val = gen.laplacian_7pt(6)
    val = val.copy()
    val[ptr[100] - 1] = -5.0   # a negative diagonal
    s = sb.Solver()
    s.analyse(n, ptr, row, gen.nested_dissection_order(6))
    inf = s.factorize(val, posdef=True)
    assert inf.flag == -6


def _illcond_spd(m, cond, kind, seed=7):
    rng = np.random.default_rng(seed)
    if kind == "graded":
        # D B D with a well-conditioned SPD B and a graded diagonal: cond(A) ~ cond
        b = rng.standard_normal((m, m))
        b = b @ b.T / m + np.eye(m)
        dd = cond ** (-0.5 * np.arange(m) / (m - 1))
        return dd[:, None] * b * dd[None, :]
    # Q diag(lambda) Q^T, eigenvalues log-spaced between 1 and 1/cond, no grading to exploit
    q, _ = np.linalg.qr(rng.standard_normal((m, m)))
    lam = cond ** (-np.arange(m) / (m - 1))
    a = (q * lam) @ q.T
    return 0.5 * (a + a.T)


@pytest.mark.parametrize("kind,cond", [("graded", 1e12), ("spectrum", 1e8), ("spectrum", 1e11)])
def test_dense_front_posdef_ill_conditioned(lib, oracle_ref, kind, cond):
    """Stability guard of the panel solve (solve_block = dtrsm in the reference,
    src/kernels/factor.hxx:95-138): on an ill-conditioned SPD front the factor must reproduce A
    as well as the reference's backward-stable LAPACK/BLAS path does (10x rule) -- a panel solve
    through an explicitly formed inverse alone loses a factor cond(L11)."""
    sb.require_gpu()
    m = 1024
    a = _illcond_spd(m, cond, kind)
    buf = np.asfortranarray(np.tril(a))
    ret = lib.sylver_b200_factor_front_posdef(m, m, buf.ctypes.data, m, None, 128, None)
    assert ret == m
    Lo, _, info = oracle_ref.factor_front_posdef(a, m)
    assert info == -1
    L, Lo = np.tril(buf), np.tril(Lo)
    # componentwise-scaled residual |A - L L^T|_ij / (|L||L|^T)_ij, the quantity Cholesky's
    # backward error bound is stated in
    def resid(F):
        r = np.abs(a - F @ F.T)
        s = np.abs(F) @ np.abs(F).T
        return float((r / s).max())
    r_gpu, r_ref = resid(L), resid(Lo)
    print(f"\nill-conditioned SPD ({kind}, cond {cond:.0e}): residual gpu {r_gpu:.2e} reference {r_ref:.2e}")
    assert r_gpu <= 10 * r_ref
