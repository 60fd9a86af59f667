"""Multi-rank factorization and solve on ONE device: the ranks run as threads
of this process over the library's in-process fabric (sylver_b200_comm_init_local), which
serves the same send/recv/all-reduce/broadcast calls NCCL serves between GPUs.  What is
checked is the multi-rank SCHEDULE -- tree partition, contribution-block hand-over, delayed
pivots crossing ranks, replicated-x solves -- against the single-rank result on the same
input; tests/test_gpu_multi.py runs the NCCL transport when the box has >= 2 GPUs."""
import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

from test_gpu_indef import _kkt_with_delays

pytestmark = pytest.mark.gpu


def _problem(kind, k):
    if kind == "kkt":
        n, ptr, row, val = gen.stokes_kkt(k)
        return n, ptr, row, val, gen.nested_dissection_order(k, dofs_per_cell=4)
    if kind == "kktd":
        return _kkt_with_delays(k, 5)
    n, ptr, row, val = (gen.laplacian_7pt if kind == "lap7" else gen.laplacian_27pt)(k)
    return n, ptr, row, val, gen.nested_dissection_order(k)


def _run(kind, k, posdef, world):
    n, ptr, row, val, order = _problem(kind, k)
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))

    def rank_body(rank, w):
        s = sb.Solver()
        assert s.analyse(n, ptr, row, order).flag == 0
        inf = s.factorize(val, posdef=posdef)
        rec = dict(flag=inf.flag, num_neg=inf.num_neg, num_two=inf.num_two, num_delay=inf.num_delay,
                   matrix_rank=inf.matrix_rank, maxfront=inf.maxfront)
        x = s.solve(b)
        y = s.solve(b, job=1)
        if not posdef:
            y = s.solve(y, job=2)
        y = s.solve(y, job=3)
        # a second factorization on the same fkeep (arenas reused) must reproduce the first
        inf2 = s.factorize(val, posdef=posdef)
        rec["again"] = (inf2.flag, inf2.num_neg, inf2.num_delay)
        x2 = s.solve(b)
        s.free()
        return rec, x, y, x2

    if world == 1:
        return [rank_body(0, 1)], (n, ptr, row, val, b)
    return sb.run_local_ranks(world, rank_body), (n, ptr, row, val, b)


CASES = [("lap27", 14, True), ("lap7", 20, True), ("lap7", 20, False), ("kkt", 10, False), ("kktd", 10, False)]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind,k,posdef", CASES, ids=lambda v: str(v))
def test_local_ranks_match_single_rank(lib, kind, k, posdef, world):
    sb.require_gpu()
    single, (n, ptr, row, val, b) = _run(kind, k, posdef, 1)
    ranks, _ = _run(kind, k, posdef, world)
    ref_rec, ref_x, _, _ = single[0]
    be_ref = gen.backward_error(n, ptr, row, val, ref_x, b)
    assert be_ref <= 1e-14
    for r, (rec, x, y, x2) in enumerate(ranks):
        # every front is factorized by exactly one rank with the same kernels: the statistics
        # are those of the single-rank run, on every rank
        for key in ("flag", "num_neg", "num_two", "num_delay", "matrix_rank", "maxfront"):
            assert rec[key] == ref_rec[key], (r, key, rec[key], ref_rec[key])
        assert rec["again"] == (ref_rec["flag"], ref_rec["num_neg"], ref_rec["num_delay"])
        be = gen.backward_error(n, ptr, row, val, x, b)
        assert be <= 1e-14 and be <= 10 * max(be_ref, 1e-16), (r, be, be_ref)
        assert np.abs(y - x).max() <= 1e-12 * max(1.0, np.abs(x).max())
        assert np.array_equal(x2, x)
        assert np.array_equal(x, ranks[0][1])        # replicated solution, bit-identical
    if kind == "kktd":
        assert ref_rec["num_delay"] > 0              # delayed columns travel up the tree


def _cross_rank_delays(n, ptr, row, val, order, worlds):
    """worlds for which some front that delays columns has its parent on another rank"""
    import ctypes as C
    s = sb.Solver()
    assert s.analyse(n, ptr, row, order).flag == 0
    s.factorize(val, posdef=False)
    L = sb.lib()
    tree = L.sylver_b200_fkeep_tree(s.fkeep)
    et = s.engine_tree()                   # the engine's (chain-coarsened) fronts
    nn, ncol, parent = et["nnodes"], et["ncol"], et["parent"]
    top = np.zeros(nn, dtype=np.int64)     # engine front -> its topmost reference front
    top[et["node_map"]] = np.arange(len(et["node_map"]))
    delayed = np.zeros(nn, dtype=bool)
    ndin = np.zeros(nn + 1, dtype=np.int64)
    for f in range(nn):
        ne = C.c_int(0)
        assert L.sylver_b200_numeric_tree_get_front_indef(tree, f, C.byref(ne), None, None) == 0
        nd = ncol[f] + ndin[f] - ne.value
        delayed[f] = nd > 0
        ndin[min(parent[f], nn)] += nd
    inner = parent < nn
    hit = []
    for world in worlds:
        own = sb.partition(s, world)[top]
        if (delayed & inner & (own != own[np.minimum(parent, nn - 1)])).any():
            hit.append(world)
    s.free()
    return hit


def test_delays_cross_rank_boundaries(lib):
    """The delayed-pivot hand-over (ghost fronts on the receiving rank) only runs when a front
    that delays columns has its parent on another rank.  Search a few scalings for such a
    case, then require the multi-rank run to reproduce the single-rank statistics."""
    sb.require_gpu()
    found = None
    for k, seed in [(8, 2), (10, 5), (12, 3), (10, 7), (12, 11), (14, 13)]:
        n, ptr, row, val, order = _kkt_with_delays(k, seed)
        worlds = _cross_rank_delays(n, ptr, row, val, order, (2, 3, 4, 8))
        if worlds:
            found = (k, seed, worlds)
            break
    if found is None:
        pytest.skip("no candidate delays a column across a rank boundary")
    k, seed, worlds = found
    print(f"\ncross-rank delays: kkt k={k} seed={seed} worlds={worlds}")
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))

    def body(rank, w):
        s = sb.Solver()
        assert s.analyse(n, ptr, row, order).flag == 0
        inf = s.factorize(val, posdef=False)
        rec = (inf.flag, inf.num_neg, inf.num_two, inf.num_delay, inf.matrix_rank)
        x = s.solve(b)
        s.free()
        return rec, x

    ref_rec, ref_x = body(0, 1)
    be_ref = gen.backward_error(n, ptr, row, val, ref_x, b)
    for world in worlds[:2]:
        for rec, x in sb.run_local_ranks(world, body):
            assert rec == ref_rec, (world, rec, ref_rec)
            be = gen.backward_error(n, ptr, row, val, x, b)
            assert be <= 1e-14 and be <= 10 * max(be_ref, 1e-16), (world, be, be_ref)


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("kind,k,exmax", [("lap27", 14, 0), ("lap7", 20, 0), ("lap27", 18, 0), ("lap27", 18, 2)])
def test_split_fronts_match_single_rank(lib, monkeypatch, kind, k, exmax, world):
    """Top-of-tree fronts split over their rank group (block-column cyclic ownership, panel
    broadcasts, contribution blocks gathered tile column by tile column -- SURVEY.md 8e).  The
    threshold is lowered so that the small test trees have split fronts several levels deep,
    including split children of split parents and groups of different sizes."""
    sb.require_gpu()
    single, (n, ptr, row, val, b) = _run(kind, k, True, 1)
    ref_x = single[0][1]
    be_ref = gen.backward_error(n, ptr, row, val, ref_x, b)
    monkeypatch.setenv("SYLVER_B200_SPLIT_MIN", "40")
    if exmax:
        # contribution pieces exchanged in many small groups (the fix for the 8-GPU hang, engine.cu)
        monkeypatch.setenv("SYLVER_B200_EX_MAX_OPS", str(exmax))
    order = _problem(kind, k)[4]

    def rank_body(rank, w):
        s = sb.Solver()
        assert s.analyse(n, ptr, row, order).flag == 0
        inf = s.factorize(val, posdef=True)
        info = s.split_info()
        x = s.solve(b)
        inf2 = s.factorize(val, posdef=True)
        x2 = s.solve(b)
        s.free()
        return inf.flag, inf2.flag, info, x, x2

    ranks = sb.run_local_ranks(world, rank_body)
    assert ranks[0][2][0] >= 1, ranks[0][2]          # the plan really has split fronts
    assert sum(r[2][1] for r in ranks) >= 2          # worked on by several ranks
    assert sum(r[2][2] for r in ranks) >= 1          # and contribution pieces cross ranks
    for flag, flag2, info, x, x2 in ranks:
        assert flag == 0 and flag2 == 0
        be = gen.backward_error(n, ptr, row, val, x, b)
        assert be <= 1e-14 and be <= 10 * max(be_ref, 1e-16), (be, be_ref)
        assert np.abs(x - ref_x).max() <= 1e-11 * max(1.0, np.abs(ref_x).max())
        assert np.array_equal(x, x2)
        assert np.array_equal(x, ranks[0][3])


def test_split_front_not_posdef_is_reported(lib, monkeypatch):
    """A non-positive pivot met inside a split front fails the factorization on every rank."""
    sb.require_gpu()
    n, ptr, row, val, order = _problem("lap27", 14)
    monkeypatch.setenv("SYLVER_B200_SPLIT_MIN", "40")
    bad = val.copy()
    # the last pivot order position is in the root front: make its diagonal entry negative
    col = int(np.argmax(order)) if order is not None else n - 1
    bad[ptr[col] - 1] = -50.0

    def rank_body(rank, w):
        s = sb.Solver()
        assert s.analyse(n, ptr, row, order).flag == 0
        inf = s.factorize(bad, posdef=True)
        s.free()
        return inf.flag

    flags = sb.run_local_ranks(2, rank_body)
    assert all(f == -6 for f in flags), flags
