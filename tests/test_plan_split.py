"""Host-side logic added for the top of the tree (no GPU needed): the engine-internal chain
coarsening of the assembly tree and the multi-GPU plan with fronts split over rank groups
(SURVEY.md 8e).  The numerical side of both is covered by tests/test_gpu_local_ranks.py."""
import os

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen


def analysed(kind, k, monkeypatch=None, amalgamate=None):
    if amalgamate is not None:
        monkeypatch.setenv("SYLVER_B200_AMALGAMATE", amalgamate)
    if kind == "kkt":
        n, ptr, row, val = gen.stokes_kkt(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
    else:
        n, ptr, row, val = (gen.laplacian_7pt if kind == "lap7" else gen.laplacian_27pt)(k)
        order = gen.nested_dissection_order(k)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, order).flag == 0
    return s, n


@pytest.mark.parametrize("kind,k", [("lap7", 16), ("lap7", 24), ("lap27", 12), ("kkt", 8)])
def test_chain_coarsening_invariants(lib, monkeypatch, kind, k):
    s, n = analysed(kind, k)
    sym, et = s.symbolic(), s.engine_tree()
    nref, G = sym["nnodes"], et["nnodes"]
    nmap = et["node_map"]
    ncol = np.diff(sym["sptr"]).astype(np.int64)
    nrow = np.diff(sym["rptr"]).astype(np.int64)
    par = sym["sparent"] - 1
    assert len(nmap) == nref and G <= nref
    # groups are runs of consecutive reference nodes, numbered in order
    assert nmap[0] == 0 and nmap[-1] == G - 1
    assert ((np.diff(nmap) == 0) | (np.diff(nmap) == 1)).all()
    assert int(et["ncol"].sum()) == n
    zeros_total = 0
    for g in range(G):
        members = np.nonzero(nmap == g)[0]
        a, b = members[0], members[-1]
        # a chain: every member but the top is the LAST child (= the node right before) of the next
        for i in range(a, b):
            assert par[i] == i + 1
        assert et["ncol"][g] == ncol[a:b + 1].sum()
        assert et["nrow"][g] == ncol[a:b + 1].sum() + (nrow[b] - ncol[b])
        # the engine parent is the group of the top member's reference parent
        assert et["parent"][g] == (nmap[par[b]] if par[b] < nref else G)
        # explicit zeros added to member j's columns: rows of the merged front it did not have
        for j in range(a, b + 1):
            rows_in_merged = et["nrow"][g] - (ncol[a:j].sum())
            zeros_total += int(ncol[j] * (rows_in_merged - nrow[j]))
            assert rows_in_merged >= nrow[j]
    panel = int((et["nrow"].astype(np.int64) * et["ncol"]).sum())
    assert zeros_total <= 0.05 * panel
    if kind == "lap7":
        assert G < nref                      # the separator chains of the 7-point stencil collapse
    # the reference-structure outputs are untouched by the coarsening
    cptr, cmap = s.cmap()
    s.free()
    s0, _ = analysed(kind, k, monkeypatch, "0")
    et0 = s0.engine_tree()
    assert et0["nnodes"] == nref and np.array_equal(et0["node_map"], np.arange(nref))
    cptr0, cmap0 = s0.cmap()
    assert np.array_equal(cptr, cptr0) and np.array_equal(cmap, cmap0)
    s0.free()


@pytest.mark.parametrize("exmax", [400, 2])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("kind,k", [("lap27", 16), ("lap7", 24)])
def test_split_plan_pairs_up_and_covers(lib, monkeypatch, kind, k, world, exmax):
    """Every piece a rank sends is received by its peer at the same level with the same extent,
    in the same order per ordered rank pair (what NCCL's grouped send/recv matching needs), every
    rank that works on a parent front ends up with every piece of each child's block exactly
    once, and the pieces of a block do not overlap."""
    monkeypatch.setenv("SYLVER_B200_SPLIT_MIN", "48")
    monkeypatch.setenv("SYLVER_B200_EX_MAX_OPS", str(exmax))      # 2: many small exchange groups
    s, n = analysed(kind, k)
    sym = s.symbolic()
    plans = [sb.plan_split(s, r, world) for r in range(world)]
    assert plans[0][0]["split_fronts"] >= 1
    assert sum(p[0]["split_member"] for p in plans) >= 2 * plans[0][0]["split_fronts"]
    sends = {}
    recvs = {}
    for r, (_, pieces) in enumerate(plans):
        for (l, f, peer, off, count, d, grp) in pieces.tolist():
            assert peer != r and count > 0 and off >= 0
            (sends if d == 0 else recvs).setdefault((r, peer) if d == 0 else (peer, r), []).append((l, f, off, count, grp))
    assert sends.keys() == recvs.keys() and len(sends) > 0
    for pair in sends:
        assert sends[pair] == recvs[pair], pair          # same order AND same exchange group on both sides
    # exchange groups (one NCCL group each): ids increase along every rank's list, never span
    # levels, and no rank has more than 400 operations in one (the bound that avoids the 8-GPU hang)
    for r, (_, pieces) in enumerate(plans):
        for d in (0, 1):
            g = pieces[pieces[:, 5] == d][:, 6]
            assert (np.diff(g) >= 0).all()
        if len(pieces):
            ids, cnts = np.unique(pieces[:, 6], return_counts=True)
            assert cnts.max() <= exmax
            for i in ids:
                assert len(set(pieces[pieces[:, 6] == i][:, 0].tolist())) == 1
    if exmax == 2 and world >= 4:      # some level was cut into several groups
        allp = np.concatenate([p[1] for p in plans])
        assert len(np.unique(allp[:, 6])) > len(np.unique(allp[:, 0]))
    # coverage: per (front, destination rank) the received pieces are disjoint intervals
    got = {}
    for (src, dst), lst in recvs.items():
        for (l, f, off, count, grp) in lst:
            got.setdefault((f, dst), []).append((off, off + count, src))
    for (f, dst), iv in got.items():
        iv.sort()
        for (a0, a1, _), (b0, b1, _) in zip(iv, iv[1:]):
            assert a1 <= b0, (f, dst)
    # an unsplit front's block travels as one piece of k * ldc doubles from its owner
    own = sb.partition(s, world)
    k_ref = (np.diff(sym["rptr"]) - np.diff(sym["sptr"])).astype(np.int64)
    for (f, dst), iv in got.items():
        if len(iv) == 1 and iv[0][0] == 0:
            ldc = (max(int(k_ref[f]), 1) + 3) // 4 * 4
            if iv[0][1] == k_ref[f] * ldc:
                assert iv[0][2] == own[f]
    # memory: split fronts are replicated over their group, nothing else is
    single = sb.plan_split(s, 0, 1)[0]
    assert single["split_fronts"] == 0 and single["sends"] == 0
    assert sum(p[0]["factor_bytes"] for p in plans) >= single["factor_bytes"]
    assert max(p[0]["factor_bytes"] for p in plans) < single["factor_bytes"]
    s.free()


def test_split_can_be_disabled(lib, monkeypatch):
    monkeypatch.setenv("SYLVER_B200_SPLIT", "0")
    s, _ = analysed("lap27", 12)
    for r in range(4):
        d, pieces = sb.plan_split(s, r, 4)
        assert d["split_fronts"] == 0
        # without split fronts every cross-rank edge is one whole-block piece
        assert all(off == 0 for (_, _, _, off, _, _, _) in pieces.tolist())
    s.free()


def _ceil(a, b):
    return -(-a // b)


@pytest.mark.parametrize("kind,k,world", [("lap27", 14, 1), ("lap7", 20, 1), ("lap27", 16, 2), ("lap27", 16, 4)])
def test_level_plan_tile_counts_match_brute_force(lib, monkeypatch, kind, k, world):
    """The tile-count prefix sums that size every batched launch of the positive definite path
    (panel solve, trailing update and its look-ahead / paired subsets, contribution update),
    recomputed from the engine tree by brute force."""
    monkeypatch.setenv("SYLVER_B200_SPLIT_MIN", "200")
    s, n = analysed(kind, k)
    et = s.engine_tree()
    G, nrow, ncol, par = et["nnodes"], et["nrow"], et["ncol"], et["parent"]
    top = np.zeros(G, dtype=np.int64)
    top[et["node_map"]] = np.arange(len(et["node_map"]))
    height = np.zeros(G + 1, dtype=np.int64)
    for g in range(G):
        height[par[g]] = max(height[par[g]], height[g] + 1)
    own = sb.partition(s, world)[top] if world > 1 else np.zeros(G, dtype=np.int64)
    NB = 128
    for rank in range(world):
        levels = sb.plan_levels(s, rank, world)
        d, _ = sb.plan_split(s, rank, world)
        nfront = 0
        for lv in levels:
            l = lv["level"]
            mine = [g for g in range(G) if height[g] == l and own[g] == rank]
            # split fronts are not in the batch; they are few: identify them by the count
            fronts = sorted(mine, key=lambda g: -ncol[g])
            assert lv["fronts"] <= len(fronts)
            if lv["fronts"] < len(fronts):
                assert world > 1 and d["split_fronts"] > 0
                continue                       # a level with a split front of this rank: covered by the split tests
            nfront += len(fronts)
            maxn = max([ncol[g] for g in fronts], default=0)
            assert len(lv["steps"]) == _ceil(maxn, NB)
            ctiles = 0
            for g in fronts:
                m_, n_ = int(nrow[g]), int(ncol[g])
                if m_ > n_:
                    tr = _ceil(m_ - (n_ & ~1), NB)
                    ctiles += tr * (tr + 1) // 2
            assert lv["contrib_tiles"] == ctiles
            for si, st in enumerate(lv["steps"]):
                p0 = si * NB
                act = [g for g in fronts if ncol[g] > p0]
                assert st["cnt"] == len(act)
                trsm = upd = updn = upd2n = 0
                for g in act:
                    m_, n_ = int(nrow[g]), int(ncol[g])
                    pw = min(NB, n_ - p0)
                    if m_ > p0 + pw:
                        trsm += _ceil(m_ - ((p0 + pw) & ~1), NB)
                    base = p0 + pw
                    if n_ > base:
                        tr, tc = _ceil(m_ - base, NB), _ceil(n_ - base, NB)
                        upd += sum(tr - tj for tj in range(tc))
                        updn += tr
                        upd2n += sum(tr - tj for tj in range(min(tc, 2)))
                assert (st["trsm"], st["upd"], st["updn"], st["upd2n"]) == (trsm, upd, updn, upd2n)
                assert st["updn"] + st["updr"] == st["upd"] == st["upd2n"] + st["upd2r"]
                assert st["wld"] % 4 == 0 and st["wld"] >= min(NB, maxn - p0)
        assert nfront > 0
    s.free()


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
os.environ["SYLVER_B200_SPLIT_MIN"] = "48"
import numpy as np, torch, torch.distributed as dist
import sylver_b200 as sb
from sylver_b200 import gen
world = {world}
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=world)
rank = dist.get_rank()
k = 14
n, ptr, row, val = gen.laplacian_27pt(k)
s = sb.Solver()
assert s.analyse(n, ptr, row, gen.nested_dissection_order(k)).flag == 0
summary, pieces = sb.plan_split(s, rank, world)
assert summary["split_fronts"] >= 1
# every rank derives the same global facts without talking to the others
facts = [None] * world
dist.all_gather_object(facts, (summary["split_fronts"], int(pieces.shape[0])))
assert len(set(f[0] for f in facts)) == 1
# replay: level by level post this level's receives and sends in plan order (what the library
# hands to one ncclGroupStart/End) and move a payload that encodes (front, offset, index)
def payload(f, off, count):
    return (f * 1.0e6 + off + np.arange(count)).astype(np.float64)
nlev = int(pieces[:, 0].max()) + 1 if len(pieces) else 0
got = 0
for l in range(nlev):
    reqs, bufs = [], []
    for i, (lv, f, peer, off, count, d, grp) in enumerate(pieces.tolist()):
        if lv != l: continue
        # gloo matches by (peer, tag): number the messages of an ordered pair in plan order
        if d == 1:
            b = torch.zeros(count, dtype=torch.float64); bufs.append((f, off, count, b))
            reqs.append(("r", peer, b))
        else:
            reqs.append(("s", peer, torch.from_numpy(payload(f, off, count))))
    seq = {{}}
    work = []
    for kind, peer, t in reqs:
        key = (kind, peer); seq[key] = seq.get(key, 0) + 1
        tag = l * 100000 + seq[key]
        work.append(dist.irecv(t, src=peer, tag=tag) if kind == "r" else dist.isend(t, dst=peer, tag=tag))
    for w in work: w.wait()
    for f, off, count, b in bufs:
        assert np.array_equal(b.numpy(), payload(f, off, count)), (f, off); got += 1
tot = torch.tensor([got]); dist.all_reduce(tot)
sent = torch.tensor([summary["sends"]]); dist.all_reduce(sent)
assert int(tot) == int(sent) > 0, (int(tot), int(sent))
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", got)
'''


@pytest.mark.parametrize("world", [2, 3])
def test_split_piece_exchange_over_gloo(lib, tmp_path, world):
    """The contribution-piece exchange of a plan with split fronts, replayed between real
    processes over gloo (the CPU stand-in for the per-level NCCL send/recv groups): every
    piece arrives where the plan says, in the order the plan says, with the right extent."""
    import socket
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER.format(root=ROOT, port=port, world=world))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out[-3000:]}"
        assert "ok" in out
