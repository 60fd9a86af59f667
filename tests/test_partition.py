"""Multi-GPU host logic on CPU: tree -> rank mapping and the per-level exchange schedule
(SURVEY.md 8e).  The world_size-2 test runs two real processes over gloo and replays the
schedule with point-to-point messages in the order the NCCL groups are issued."""
import os
import subprocess
import sys

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def analysed(kind="lap27", k=16):
    n, ptr, row, val = (gen.laplacian_27pt if kind == "lap27" else gen.laplacian_7pt)(k)
    s = sb.Solver()
    assert s.analyse(n, ptr, row, gen.nested_dissection_order(k)).flag == 0
    return s


def front_flops(sym):
    ncol = np.diff(sym["sptr"]).astype(np.float64)
    nrow = np.diff(sym["rptr"]).astype(np.float64)
    mm = nrow - ncol
    return ncol * mm * mm + mm * ncol * (ncol + 1) + ncol * (ncol + 1) * (2 * ncol + 1) / 6.0


@pytest.mark.parametrize("kind,k", [("lap27", 16), ("lap7", 24)])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_is_complete_and_balanced(lib, kind, k, world):
    s = analysed(kind, k)
    sym = s.symbolic()
    own = sb.partition(s, world)
    assert own.shape[0] == sym["nnodes"]
    assert own.min() >= 0 and own.max() < world
    w = front_flops(sym)
    assert abs(w.sum() - sym["num_flops"]) <= 1e-6 * sym["num_flops"]
    load = np.bincount(own, weights=w, minlength=world)
    assert (load > 0).all()
    # the subtree part balances well; the serial top of the tree is bounded by the root chain
    assert load.max() <= 2.2 * load.mean()
    # locality: below the cut whole subtrees stay on one rank -> few cross-rank edges
    parent = sym["sparent"] - 1
    inner = parent < sym["nnodes"]
    cross = int((own[inner] != own[parent[inner]]).sum())
    assert cross <= 6 * world
    s.free()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_schedule_pairs_up(lib, world):
    """Every send of rank a to b at level l has the matching receive on b (same level, same
    front): the per-level NCCL groups of all ranks fit together."""
    s = analysed("lap27", 16)
    sym = s.symbolic()
    own = sb.partition(s, world)
    plans = [sb.plan_exchanges(s, r, world) for r in range(world)]
    sends = {(int(l), int(f), r, int(p)) for r in range(world) for (l, f, p, d) in plans[r] if d == 0}
    recvs = {(int(l), int(f), int(p), r) for r in range(world) for (l, f, p, d) in plans[r] if d == 1}
    assert sends == recvs and len(sends) > 0
    parent = sym["sparent"] - 1
    for (l, f, a, b) in sends:
        assert own[f] == a and own[parent[f]] == b
    # every cross-rank edge with a non-empty contribution block is scheduled exactly once
    k = np.diff(sym["rptr"]) - np.diff(sym["sptr"])
    inner = (parent < sym["nnodes"]) & (k > 0)
    want = {int(f) for f in np.nonzero(inner)[0] if own[f] != own[parent[f]]}
    assert {f for (_, f, _, _) in sends} == want
    s.free()


WORKER = r'''
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import sylver_b200 as sb
from sylver_b200 import gen
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
L = sb.lib()
L.sylver_b200_comm_set_virtual(rank, 2)          # planning-only communicator
assert L.sylver_b200_comm_rank() == rank and L.sylver_b200_comm_world() == 2
k = 12
n, ptr, row, val = gen.laplacian_27pt(k)
s = sb.Solver()
assert s.analyse(n, ptr, row, gen.nested_dissection_order(k)).flag == 0
sym = s.symbolic()
own = sb.partition(s, 2)
# both ranks derive the same map without talking to each other
gathered = [None, None]
dist.all_gather_object(gathered, own.tolist())
assert gathered[0] == gathered[1]
plan = sb.plan_exchanges(s, rank, 2)
kk = np.diff(sym["rptr"]) - np.diff(sym["sptr"])
# replay: level by level, post this level's receives and sends (the order the library issues
# its NCCL groups) and move a recognisable payload per contribution block
nlev = int(plan[:, 0].max()) + 1 if len(plan) else 0
got = 0
for l in range(nlev):
    reqs, bufs = [], []
    for (lv, f, peer, d) in plan:
        if lv != l: continue
        if d == 1:
            b = torch.zeros(int(kk[f])); bufs.append((f, b)); reqs.append(dist.irecv(b, src=int(peer), tag=int(f)))
        else:
            reqs.append(dist.isend(torch.full((int(kk[f]),), float(f)), dst=int(peer), tag=int(f)))
    for r in reqs: r.wait()
    for f, b in bufs:
        assert float(b.min()) == float(f) == float(b.max()); got += 1
tot = torch.tensor([got]); dist.all_reduce(tot)
par = np.minimum(sym["sparent"] - 1, sym["nnodes"] - 1)
cross = int(((own != own[par]) & (sym["sparent"] - 1 < sym["nnodes"]) & (kk > 0)).sum())
assert int(tot) == cross, (int(tot), cross)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", got)
'''


def test_two_rank_schedule_over_gloo(lib, tmp_path):
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{out}"
        assert "ok" in out
