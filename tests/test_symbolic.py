"""Symbolic parity (bit-exact): the product's C++ analysis (sylver_b200/csrc/analyse.cpp)
against oracle/symbolic.py, the statement-by-statement restatement of SPRAL's
basic_analyse (spral/src/core_analyse.f90:38-150) and SyLVER's build_map
(src/spldlt_analyse_mod.F90:130-232); the oracle itself is pinned by brute-force
definitional checks (dense symbolic Cholesky) and by the committed golden fixtures
in tests/golden/ (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen
from oracle import symbolic as osym

KEYS = ("sptr", "sparent", "rptr", "rlist", "nptr", "nlist", "order", "invp")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def random_sym_pattern(n, density, seed):
    rng = np.random.default_rng(seed)
    cols, rows = [], []
    for j in range(n):
        r = [j] + [i for i in range(j + 1, n) if rng.random() < density]
        cols += [j] * len(r)
        rows += r
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, np.array(cols) + 1, 1)
    ptr = np.cumsum(ptr) + 1
    row = (np.array(rows) + 1).astype(np.int32)
    order = (rng.permutation(n) + 1).astype(np.int32)
    return n, ptr, row, order


def cases():
    out = {}
    for k in (3, 5, 6):
        n, ptr, row, _ = gen.laplacian_7pt(k)
        out[f"lap7_{k}"] = (n, ptr, row, gen.nested_dissection_order(k))
    n, ptr, row, _ = gen.laplacian_27pt(5)
    out["lap27_5"] = (n, ptr, row, gen.nested_dissection_order(5))
    n, ptr, row, _ = gen.stokes_kkt(3)
    out["kkt_3"] = (n, ptr, row, gen.nested_dissection_order(3, dofs_per_cell=4))
    out["rand_60"] = random_sym_pattern(60, 0.08, 1)
    out["rand_120"] = random_sym_pattern(120, 0.03, 2)
    out["rand_dense_40"] = random_sym_pattern(40, 0.5, 3)
    # reference fixtures: 4x4 simple_mat (tests/sylver_test_mod.F90:120-157) and the 3x3
    # tridiagonal of the C example (examples/C/spldlt_simple_example_c.c:27-40), natural order
    out["simple_mat"] = (4, np.array([1, 4, 5, 7, 8], dtype=np.int64),
                         np.array([1, 2, 4, 2, 3, 4, 4], dtype=np.int32), np.arange(1, 5, dtype=np.int32))
    out["tridiag3"] = (3, np.array([1, 3, 5, 6], dtype=np.int64),
                       np.array([1, 2, 2, 3, 3], dtype=np.int32), np.arange(1, 4, dtype=np.int32))
    # diagonal matrix: n independent roots
    out["diag_7"] = (7, np.arange(1, 9, dtype=np.int64), np.arange(1, 8, dtype=np.int32),
                     np.arange(1, 8, dtype=np.int32))
    return out


CASES = cases()


def product_symbolic(n, ptr, row, order, nemin):
    s = sb.Solver()
    s.options.nemin = nemin
    inf = s.analyse(n, ptr, row, order)
    assert inf.flag == 0, inf.flag
    sym = s.symbolic()
    cm = s.cmap()
    info = dict(num_factor=inf.num_factor, num_flops=inf.num_flops, num_sup=inf.num_sup,
                maxfront=inf.maxfront, order_out=s.order.copy())
    s.free()
    return sym, cm, info


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("nemin", [1, 8, 32])
def test_cpp_analysis_matches_reference_restatement(lib, name, nemin):
    n, ptr, row, order = CASES[name]
    sym, (cptr, cmap), info = product_symbolic(n, ptr, row, order, nemin)
    ref = osym.analyse(n, ptr, row, order, nemin=nemin)
    assert sym["nnodes"] == ref["nnodes"]
    for k in KEYS:
        assert np.array_equal(sym[k], ref[k]), f"{name}: {k} differs"
    assert info["num_factor"] == ref["num_factor"]
    assert info["num_flops"] == ref["num_flops"]
    # order is overwritten with the final elimination order (spldlt_analyse_mod.F90:878)
    assert np.array_equal(info["order_out"], ref["order"])
    rcptr, rcmap = osym.cmap(ref)
    assert np.array_equal(cptr, rcptr)
    assert np.array_equal(cmap, rcmap)


def dense_symbolic_cholesky(n, ptr, row, pos):
    """Definition: structure of L for the permuted matrix, by explicit elimination."""
    adj = [set() for _ in range(n)]
    for j in range(n):
        for p in range(ptr[j] - 1, ptr[j + 1] - 1):
            i = row[p] - 1
            a, b = pos[i] - 1, pos[j] - 1
            if a != b:
                adj[min(a, b)].add(max(a, b))
    struct = []
    for j in range(n):
        s = sorted(adj[j])
        struct.append(s)
        if s:
            par = s[0]
            adj[par] |= set(s[1:])
    return struct


@pytest.mark.parametrize("name", ["lap7_3", "rand_60", "rand_dense_40", "kkt_3", "simple_mat"])
def test_oracle_restatement_against_definition(name):
    """Pins oracle/symbolic.py: with nemin=1 only fundamental supernodes merge, so every
    supernode's row list must equal the elimination structure of its first column."""
    n, ptr, row, order = CASES[name]
    ref = osym.analyse(n, ptr, row, order, nemin=1)
    struct = dense_symbolic_cholesky(n, ptr, row, ref["order"])
    sptr, rptr, rlist = ref["sptr"], ref["rptr"], ref["rlist"]
    nfact = 0
    for nd in range(ref["nnodes"]):
        c0 = sptr[nd] - 1
        rows = rlist[rptr[nd] - 1:rptr[nd + 1] - 1] - 1
        want = [c0] + struct[c0]
        assert list(rows) == want, f"node {nd}"
        ncol = sptr[nd + 1] - sptr[nd]
        # columns of a supernode are nested
        for j in range(1, ncol):
            assert struct[c0 + j] == want[j + 1:]
        nfact += sum(len(want) - j for j in range(ncol))
    assert nfact == ref["num_factor"]
    # etree parent of the last column of each supernode is the first column of its parent
    for nd in range(ref["nnodes"]):
        ncol = sptr[nd + 1] - sptr[nd]
        last = struct[sptr[nd] - 1 + ncol - 1]
        p = ref["sparent"][nd] - 1
        if last:
            assert sptr[p] - 1 <= last[0] < sptr[p + 1] - 1
        else:
            assert p == ref["nnodes"]


@pytest.mark.parametrize("name", ["lap7_5", "rand_120", "kkt_3"])
def test_build_map_scatters_every_entry_once(name):
    """nlist (src,dest) pairs: each stored entry of A lands exactly once, at the
    (row, col) of the front that owns its column (src/spldlt_analyse_mod.F90:130-232)."""
    n, ptr, row, order = CASES[name]
    ref = osym.analyse(n, ptr, row, order, nemin=8)
    pos = ref["order"]
    nl = ref["nlist"].reshape(-1, 2)
    assert sorted(nl[:, 0]) == list(range(1, int(ptr[-1])))
    colof = np.repeat(np.arange(n), np.diff(ptr))
    for nd in range(ref["nnodes"]):
        rows = ref["rlist"][ref["rptr"][nd] - 1:ref["rptr"][nd + 1] - 1]
        nrow = len(rows)
        for e in range(ref["nptr"][nd] - 1, ref["nptr"][nd + 1] - 1):
            src, dest = nl[e]
            c, r = divmod(dest - 1, nrow)
            i, j = row[src - 1] - 1, colof[src - 1]
            a, b = sorted((pos[i], pos[j]))
            assert rows[c] == a and rows[r] == b


def test_golden_symbolic_fixtures(lib):
    """Committed vectors (tests/golden/symbolic.json, generated by make_golden.py from the
    restatement): guards both implementations against silent drift."""
    g = json.load(open(os.path.join(GOLDEN, "symbolic.json")))
    for name, rec in g.items():
        n, ptr, row, order = CASES[name]
        sym, _, info = product_symbolic(n, ptr, row, order, rec["nemin"])
        for k in KEYS:
            assert sym[k].tolist() == rec[k], f"{name}: {k}"
        assert info["num_flops"] == rec["num_flops"] and info["num_factor"] == rec["num_factor"]


def test_analyse_argument_errors(lib):
    """Flags of /root/reference/tests/sylver_test_mod.F90 test_errors that belong to this path."""
    n, ptr, row, order = CASES["simple_mat"]
    s = sb.Solver()
    assert s.analyse(-1, ptr, row, order).flag == -2          # ERROR_A_N_OOR
    s.free()
    s = sb.Solver()
    bad = order.copy(); bad[1] = bad[0]                        # duplicate position
    assert s.analyse(n, ptr, row, bad).flag == -8             # ERROR_ORDER
    s.free()
    s = sb.Solver()
    bad = order.copy(); bad[0] = n + 3
    assert s.analyse(n, ptr, row, bad).flag == -8
    s.free()
    s = sb.Solver()
    inf = s.analyse(0, np.array([1], dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert inf.flag == 0                                       # n = 0 is legal (test_special)
    s.free()
    # factorize before analyse -> ERROR_CALL_SEQUENCE
    s = sb.Solver()
    assert s.factorize(np.ones(7), posdef=True).flag == -1
    s.free()
