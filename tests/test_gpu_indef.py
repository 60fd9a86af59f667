"""GPU parity tests, indefinite APTP path (through the C ABI).

Parity criteria of the north star: identical inertia (num_neg) vs the reference engine,
scaled backward error <= 1e-14 and within 10x of the reference.  The reference side is
oracle/_ref when it travelled with the snapshot, and always the committed golden vectors
(tests/golden/numeric.json, generated from oracle/_ref by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "numeric.json")))


def rebuild_D(d, ne):
    D = np.zeros((ne, ne))
    i = 0
    while i < ne:
        if i + 1 == ne or np.isfinite(d[2 * i + 2]):
            D[i, i] = 1.0 / d[2 * i] if d[2 * i] != 0 else 0.0
            i += 1
        else:
            blk = np.array([[d[2 * i], d[2 * i + 1]], [d[2 * i + 1], d[2 * i + 3]]])
            D[i:i + 2, i:i + 2] = np.linalg.inv(blk)
            i += 2
    return D


def inertia_from_d(d, ne):
    neg = two = 0
    i = 0
    while i < ne:
        if i + 1 == ne or np.isfinite(d[2 * i + 2]):
            neg += d[2 * i] < 0
            i += 1
        else:
            a11, a21, a22 = d[2 * i], d[2 * i + 1], d[2 * i + 3]
            det = a11 * a22 - a21 * a21
            two += 1
            neg += 1 if det < 0 else (2 if a11 + a22 < 0 else 0)
            i += 2
    return int(neg), two


@pytest.mark.parametrize("rec", GOLDEN["dense"], ids=lambda r: f"{r['m']}x{r['n']}{'d' if r['delays'] else ''}")
def test_dense_front_indef(lib, rec):
    """Dense single front (reference harness tests/testing_factor_node_indef.hxx:44-460):
    P A P^T = L D L^T on the eliminated columns, |l_ij| <= 1/u, inertia vs reference."""
    sb.require_gpu()
    m, n = rec["m"], rec["n"]
    rng = gen.GlibcRand(1)
    a = gen.dense_sym_indef(m, rng=rng)
    if rec["delays"]:
        a = gen.cause_delays(a, rng)
    r = sb.factor_front_indef(a, n)
    ne = r["nelim"]
    assert 0 <= ne <= n, ne
    st = r["stats"]
    assert st.num_delay == n - ne
    if not rec["delays"]:
        assert ne == n
    perm = r["perm"] - 1
    assert sorted(perm) == list(range(n))
    if ne == 0:
        return
    L = np.tril(r["L"][:, :ne], -1) + np.eye(m, ne)
    D = rebuild_D(r["d"], ne)
    rows = np.concatenate([perm, np.arange(n, m)])
    ap = a[np.ix_(rows, rows)]
    recon = L @ D @ L.T
    # eliminated columns reproduce A; scaled like the reference's backward error
    err = np.abs(ap[:, :ne] - recon[:, :ne]).max() / np.abs(a).max()
    assert err <= 5e-14 * max(1.0, np.abs(L).max()) * m ** 0.5, err
    assert np.abs(np.tril(r["L"][:, :ne], -1)).max() <= 100.0 * (1 + 1e-12)     # |l| <= 1/u
    # Schur complement of the uneliminated part: failed columns (still in the panel) and contribution
    S = ap[ne:, ne:] - recon[ne:, ne:]
    if ne < n:
        Sf = np.tril(r["L"][ne:, ne:n])
        assert np.abs(Sf - np.tril(S)[:, :n - ne]).max() <= 1e-10 * max(1.0, np.abs(S).max())
    if m > n:
        C = np.tril(r["contrib"])
        want = np.tril(-recon[n:, n:])
        assert np.abs(C - want).max() <= 1e-10 * max(1.0, np.abs(want).max())
    neg, two = inertia_from_d(r["d"], ne)
    assert (neg, two) == (st.num_neg, st.num_two)
    if ne == m:
        # full factorization: inertia is an invariant (Sylvester) -> identical to the reference
        assert st.num_neg == rec["num_neg"]
        # solve and check the reference harness bound u*bwderr <= 5e-14
        b = a @ np.ones(m)
        y = np.linalg.solve(L, b[rows])
        z = np.linalg.solve(D, y)
        x = np.empty(m)
        x[rows] = np.linalg.solve(L.T, z)
        bw = np.abs(a @ x - b).max() / (np.abs(a).sum(axis=1).max() * np.abs(x).max() + np.abs(b).max())
        assert 0.01 * bw <= 5e-14, bw


def _tree(rec):
    if rec["kind"] == "kkt":
        n, ptr, row, val = gen.stokes_kkt(rec["k"])
        order = gen.nested_dissection_order(rec["k"], dofs_per_cell=4)
    else:
        n, ptr, row, val = (gen.laplacian_7pt if rec["kind"] == "lap7" else gen.laplacian_27pt)(rec["k"])
        order = gen.nested_dissection_order(rec["k"])
    return n, ptr, row, val, order


@pytest.mark.parametrize("rec", [r for r in GOLDEN["trees"] if not r["posdef"]],
                         ids=lambda r: f"{r['kind']}_{r['k']}")
def test_indef_tree_parity(lib, rec):
    sb.require_gpu()
    n, ptr, row, val, order = _tree(rec)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, order).flag == 0
    inf = s.factorize(val, posdef=False)
    assert inf.flag == rec["flag"], inf.flag
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = s.solve(b)
    be = gen.backward_error(n, ptr, row, val, x, b)
    print(f"\n{rec['kind']}_{rec['k']}: bwderr={be:.2e} (ref {rec['bwderr']:.2e}) num_neg={inf.num_neg} "
          f"num_two={inf.num_two} num_delay={inf.num_delay} (ref {rec['num_delay']}) {s.timings()}")
    assert inf.num_neg == rec["num_neg"]                     # inertia identical
    assert inf.matrix_rank == n - rec["num_zero"]
    assert be <= 1e-14 and be <= 10 * max(rec["bwderr"], 1e-16)
    # partial solves compose to the full solve (job 1, 2, 3 / 1, 4)
    y = s.solve(b, job=1)
    y = s.solve(y, job=2)
    y = s.solve(y, job=3)
    assert np.abs(y - x).max() <= 1e-12 * np.abs(x).max()
    s.free()


def _kkt_with_delays(k, seed):
    """Stokes KKT whose velocity block is scaled down row/column-wise at random: pivots
    that look fine inside a front become unacceptable -> delayed columns across levels."""
    n, ptr, row, val = gen.stokes_kkt(k)
    rng = np.random.default_rng(seed)
    sc = np.ones(n)
    pick = rng.choice(n, size=n // 6, replace=False)
    sc[pick] = 10.0 ** rng.uniform(-6, -3, size=pick.size)
    col = np.repeat(np.arange(n), np.diff(ptr))
    val = val * sc[col] * sc[row - 1]
    return n, ptr, row, val, gen.nested_dissection_order(k, dofs_per_cell=4)


@pytest.mark.parametrize("k,seed", [(4, 1), (8, 2), (12, 3)])
def test_indef_tree_with_delays(lib, oracle_ref, k, seed):
    """Delayed pivots must travel up the tree (assemble_delays) and the result must match
    the reference engine run side by side on the same input."""
    sb.require_gpu()
    n, ptr, row, val, order = _kkt_with_delays(k, seed)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, order).flag == 0
    inf = s.factorize(val, posdef=False)
    assert inf.flag >= 0
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = s.solve(b)
    be = gen.backward_error(n, ptr, row, val, x, b)
    ot = oracle_ref.OracleTree(s.symbolic())
    ot.factor(val, False)
    xo = ot.solve_original(b)
    beo = gen.backward_error(n, ptr, row, val, xo, b)
    print(f"\nkkt_delays k={k}: bwderr={be:.2e} oracle={beo:.2e} num_delay={inf.num_delay} "
          f"(oracle {ot.stats.num_delay}) num_neg={inf.num_neg} (oracle {ot.stats.num_neg}) two={inf.num_two}")
    assert inf.num_neg == ot.stats.num_neg
    assert be <= 1e-14 and be <= 10 * max(beo, 1e-16)
    assert inf.num_delay > 0 or ot.stats.num_delay == 0
    s.free()


def test_refactor_indef_same_tree(lib):
    """A second spldlt_factorize on the same fkeep (new values) reuses the arenas."""
    sb.require_gpu()
    n, ptr, row, val = gen.stokes_kkt(6)
    order = gen.nested_dissection_order(6, dofs_per_cell=4)
    s = sb.Solver()
    s.analyse(n, ptr, row, order)
    i1 = s.factorize(val, posdef=False)
    neg1 = i1.num_neg
    i2 = s.factorize(2.0 * val, posdef=False)
    assert i2.flag == 0 and i2.num_neg == neg1
    b = gen.sym_matvec(n, ptr, row, 2.0 * val, np.ones(n))
    x = s.solve(b)
    assert gen.backward_error(n, ptr, row, 2.0 * val, x, b) <= 1e-14


def test_singular_matrix_warning(lib):
    """simple_sing_mat semantics (tests/sylver_test_mod.F90): action=true -> warning 7 and
    matrix_rank < n; action=false -> error -5."""
    sb.require_gpu()
    # 3x3 with an exactly dependent row: [[1,1,0],[1,1,0],[0,0,2]]
    ptr = np.array([1, 3, 4, 5], dtype=np.int64)
    row = np.array([1, 2, 2, 3], dtype=np.int32)
    val = np.array([1.0, 1.0, 1.0, 2.0])
    s = sb.Solver()
    s.analyse(3, ptr, row, np.arange(1, 4, dtype=np.int32))
    inf = s.factorize(val, posdef=False)
    assert inf.flag == 7 and inf.matrix_rank == 2
    s2 = sb.Solver()
    s2.options.action = False
    s2.analyse(3, ptr, row, np.arange(1, 4, dtype=np.int32))
    assert s2.factorize(val, posdef=False).flag == -5
