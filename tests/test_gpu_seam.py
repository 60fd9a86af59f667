"""GPU tests through the INTERNAL seam the reference's Fortran layer calls
(src/SymbolicTree.cxx:109-137, src/NumericTree.cxx:26-136, src/NumericTreePosdef.cxx:19-83):
spldlt_create_symbolic_tree -> spldlt_create_numeric_tree[_posdef]_dbl ->
spldlt_tree_solve_{fwd,diag,bwd,diag_bwd}_dbl, driven exactly as
src/spldlt_factorize_mod.F90:473-567 (factor_core) and :901-1060 (solve) drive them: 1-based
index arrays borrowed from the caller, `scaling` and `x` already permuted to elimination order.
The result must equal the public API's (same engine underneath) and the reference oracle's."""
import ctypes as C

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

pytestmark = pytest.mark.gpu


def _case(kind, k):
    if kind == "kkt":
        n, ptr, row, val = gen.stokes_kkt(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
    elif kind == "lap27":
        n, ptr, row, val = gen.laplacian_27pt(k)
        order = gen.nested_dissection_order(k)
    else:
        n, ptr, row, val = gen.laplacian_7pt(k)
        order = gen.nested_dissection_order(k)
    return n, ptr, row, val, order


class Seam:
    """What factor_core / the solve routine of spldlt_factorize_mod.F90 do, in ctypes."""

    def __init__(self, L, sym):
        self.L = L
        self.sym = sym          # keeps the borrowed arrays alive (the seam does not copy them)
        nn = sym["nnodes"]
        self.tree = L.spldlt_create_symbolic_tree(
            None, sym["n"], nn, sb._ptr(sym["sptr"]), sb._ptr(sym["sparent"]), sb._ptr(sym["rptr"]),
            sb._ptr(sym["rlist"]), sb._ptr(sym["nptr"]), sb._ptr(sym["nlist"]), 0, None, None, None, None)
        assert self.tree
        self.num = None
        self.posdef = False

    def factor(self, posdef, val, scaling_perm=None, options=None):
        self.free_numeric()
        opt = options or sb.default_options_c()
        st = sb.InformC()
        aval = np.ascontiguousarray(val, dtype=np.float64)
        if posdef:
            self.num = self.L.spldlt_create_numeric_tree_posdef_dbl(None, self.tree, sb._ptr(aval), sb._ptr(scaling_perm),
                                                                    None, C.byref(opt), C.byref(st))
        else:
            self.num = self.L.spldlt_create_numeric_tree_dbl(False, None, self.tree, sb._ptr(aval), sb._ptr(scaling_perm),
                                                             None, C.byref(opt), C.byref(st))
        self.posdef = posdef
        assert self.num
        return st

    def solve(self, b, scaling_perm=None, split_diag=False):
        """x = A^-1 b: permute to elimination order (invp), scale, fwd, [diag, bwd | diag_bwd],
        scale, permute back (spldlt_factorize_mod.F90:958-1052)."""
        invp = self.sym["invp"] - 1
        x = np.ascontiguousarray(b[invp], dtype=np.float64)
        n = x.size
        if scaling_perm is not None:
            x *= scaling_perm
        L = self.L
        if self.posdef:
            assert L.spldlt_tree_solve_fwd_posdef_dbl(self.num, 1, sb._ptr(x), n) == 0
            assert L.spldlt_tree_solve_bwd_posdef_dbl(self.num, 1, sb._ptr(x), n) == 0
        else:
            assert L.spldlt_tree_solve_fwd_dbl(False, self.num, 1, sb._ptr(x), n) == 0
            if split_diag:
                assert L.spldlt_tree_solve_diag_dbl(False, self.num, 1, sb._ptr(x), n) == 0
                assert L.spldlt_tree_solve_bwd_dbl(False, self.num, 1, sb._ptr(x), n) == 0
            else:
                assert L.spldlt_tree_solve_diag_bwd_dbl(False, self.num, 1, sb._ptr(x), n) == 0
        if scaling_perm is not None:
            x *= scaling_perm
        out = np.empty(n)
        out[invp] = x
        return out

    def free_numeric(self):
        if self.num:
            if self.posdef:
                self.L.spldlt_destroy_numeric_tree_posdef_dbl(self.num)
            else:
                self.L.spldlt_destroy_numeric_tree_dbl(False, self.num)
            self.num = None

    def close(self):
        self.free_numeric()
        self.L.spldlt_destroy_symbolic_tree(self.tree)


@pytest.mark.parametrize("kind,k,posdef", [("lap7", 12, True), ("lap27", 10, True), ("lap7", 12, False), ("kkt", 8, False)])
def test_seam_matches_public_api_and_oracle(lib, oracle_ref, kind, k, posdef):
    sb.require_gpu()
    n, ptr, row, val, order = _case(kind, k)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, order).flag == 0
    inf = s.factorize(val, posdef=posdef)
    assert inf.flag == 0
    b = gen.sym_matvec(n, ptr, row, val, np.arange(1, n + 1) / n)
    x_api = s.solve(b)
    sym = s.symbolic()
    seam = Seam(lib, sym)
    st = seam.factor(posdef, val)
    assert st.flag == 0
    assert (st.num_neg, st.num_two, st.num_delay) == (inf.num_neg, inf.num_two, inf.num_delay)
    x = seam.solve(b)
    assert np.abs(x - x_api).max() <= 1e-13 * np.abs(x_api).max()
    if not posdef:
        x2 = seam.solve(b, split_diag=True)
        assert np.abs(x2 - x).max() <= 1e-13 * np.abs(x).max()
    be = gen.backward_error(n, ptr, row, val, x, b)
    ot = oracle_ref.OracleTree(sym)
    ot.factor(val, posdef)
    xo = ot.solve_original(b)
    beo = gen.backward_error(n, ptr, row, val, xo, b)
    assert ot.stats.num_neg == st.num_neg
    assert be <= 1e-14 and be <= 10 * max(beo, 1e-16), (be, beo)
    seam.close()
    s.free()


def test_seam_scaling_is_in_elimination_order(lib):
    """`scaling` arrives already permuted (factor_core passes fkeep%scaling, built as
    scaling(i) = scale(invp(i)), spldlt_factorize_mod.F90:744-749)."""
    sb.require_gpu()
    n, ptr, row, val, order = _case("kkt", 6)
    rng = np.random.default_rng(5)
    scale = 10.0 ** rng.uniform(-2, 2, n)
    s = sb.Solver()
    s.analyse(n, ptr, row, order)
    sym = s.symbolic()
    sperm = np.ascontiguousarray(scale[sym["invp"] - 1])
    seam = Seam(lib, sym)
    st = seam.factor(False, val, scaling_perm=sperm)
    assert st.flag == 0
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = seam.solve(b, scaling_perm=sperm)
    assert gen.backward_error(n, ptr, row, val, x, b) <= 1e-14
    # the public API with the same user scaling gives the same inertia and solution
    inf = s.factorize(val, posdef=False, scale=scale.copy())
    assert inf.num_neg == st.num_neg
    xa = s.solve(b)
    assert np.abs(xa - x).max() <= 1e-12 * np.abs(x).max()
    seam.close()
    s.free()


def test_seam_ignores_subtree_partition(lib):
    """nsubtrees > 0 (the reference's default prune_tree = .true. call shape): the partition is
    accepted and ignored -- the engine factorizes every node itself, child_contrib is never
    touched and the tree solves cover all nodes (include/sylver_b200.h, seam notes)."""
    sb.require_gpu()
    n, ptr, row, val, order = _case("lap7", 8)
    s = sb.Solver()
    s.analyse(n, ptr, row, order)
    sym = s.symbolic()
    nn = sym["nnodes"]
    # a plausible partition: the first leaf-side third of the nodes as one "subtree"
    sub = np.array([max(nn // 3, 1)], dtype=np.int32)
    small = np.zeros(nn, dtype=np.int32)
    small[: nn // 3] = 1
    dest = np.zeros(1, dtype=np.int32)
    loc = np.full(nn, -1, dtype=np.int32)
    seam = Seam(lib, sym)
    lib.spldlt_destroy_symbolic_tree(seam.tree)
    seam.tree = lib.spldlt_create_symbolic_tree(None, sym["n"], nn, sb._ptr(sym["sptr"]), sb._ptr(sym["sparent"]),
                                                sb._ptr(sym["rptr"]), sb._ptr(sym["rlist"]), sb._ptr(sym["nptr"]),
                                                sb._ptr(sym["nlist"]), 1, sb._ptr(sub), sb._ptr(small), sb._ptr(dest),
                                                sb._ptr(loc))
    assert seam.tree
    contrib = (C.c_void_p * 1)(None)
    opt, st = sb.default_options_c(), sb.InformC()
    seam.num = lib.spldlt_create_numeric_tree_posdef_dbl(None, seam.tree, sb._ptr(val), None, contrib, C.byref(opt), C.byref(st))
    seam.posdef = True
    assert seam.num and st.flag == 0 and contrib[0] is None
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = seam.solve(b)
    assert gen.backward_error(n, ptr, row, val, x, b) <= 1e-14
    seam.close()
    s.free()


def test_reanalyse_on_same_handles(lib):
    """analyse -> factorize -> analyse (another matrix, same akeep) -> factorize (same fkeep):
    the reference C interface reuses both handles (sylver_ciface.F90:436-443,601-608); the
    numeric tree of the first analysis must be rebuilt, not refactored."""
    sb.require_gpu()
    s = sb.Solver()
    n, ptr, row, val, order = _case("lap7", 10)
    s.analyse(n, ptr, row, order)
    assert s.factorize(val, posdef=True).flag == 0
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    assert gen.backward_error(n, ptr, row, val, s.solve(b), b) <= 1e-14
    n2, ptr2, row2, val2, order2 = _case("lap27", 7)
    s.analyse(n2, ptr2, row2, order2)
    # solving with the stale factors is a call-sequence error, not a crash
    s.n = n2
    s.solve(np.ones(n2))
    assert s.inform.flag == -1
    assert s.factorize(val2, posdef=True).flag == 0
    b2 = gen.sym_matvec(n2, ptr2, row2, val2, np.ones(n2))
    assert gen.backward_error(n2, ptr2, row2, val2, s.solve(b2), b2) <= 1e-14
    # and back to an indefinite problem on the same handles
    n3, ptr3, row3, val3, order3 = _case("kkt", 5)
    s.analyse(n3, ptr3, row3, order3)
    inf = s.factorize(val3, posdef=False)
    assert inf.flag == 0 and inf.num_neg == 5 ** 3
    s.free()


def test_refactor_reads_options_again(lib):
    """Second spldlt_factorize on the same fkeep with different options: the reference reads
    options on every call (spldlt_factorize_mod.F90:855-864 for action)."""
    sb.require_gpu()
    ptr = np.array([1, 3, 4, 5], dtype=np.int64)
    row = np.array([1, 2, 2, 3], dtype=np.int32)
    val = np.array([1.0, 1.0, 1.0, 2.0])          # singular: rows 1 and 2 equal
    s = sb.Solver()
    s.analyse(3, ptr, row, np.arange(1, 4, dtype=np.int32))
    assert s.factorize(val, posdef=False).flag == 7          # action = true: warning
    s.options.action = False
    assert s.factorize(val, posdef=False).flag == -5         # same fkeep, action = false: error
    s.options.action = True
    assert s.factorize(val, posdef=False).flag == 7
    s.free()
