"""One rank of a multi-GPU factorization (launch with torchrun, one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/multi_gpu_worker.py lap27 40

Every rank runs the same analyse, factorizes its share of the assembly tree and takes part in
the distributed solve; rank 0 prints timings and the backward error."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb
from sylver_b200 import gen


def main():
    import faulthandler
    # a hang must leave evidence: every rank dumps its Python stack after SYLVER_WORKER_DUMP_S seconds
    faulthandler.dump_traceback_later(int(os.environ.get("SYLVER_WORKER_DUMP_S", "240")), exit=True)
    kind, k = sys.argv[1], int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    posdef = not (len(sys.argv) > 4 and sys.argv[4] == "indef")      # "indef": APTP LDL^T on the same matrix
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    sb.comm_init_from_torch(dist, local)
    if kind == "kkt":
        n, ptr, row, val = gen.stokes_kkt(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
        posdef = False
    else:
        n, ptr, row, val = (gen.laplacian_27pt if kind == "lap27" else gen.laplacian_7pt)(k)
        order = gen.nested_dissection_order(k)
    s = sb.Solver()
    t_a = time.perf_counter()
    inf = s.analyse(n, ptr, row, order)
    assert inf.flag == 0
    flops = int(inf.num_flops)
    print(f"[rank {rank}] analysed n={n} in {time.perf_counter() - t_a:.1f} s, flops {flops:.3e}", file=sys.stderr, flush=True)
    times = []
    for r in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        inf = s.factorize(val, posdef=posdef)
        torch.cuda.synchronize(); dist.barrier()
        times.append(time.perf_counter() - t0)
        print(f"[rank {rank}] factorization {r}: {times[-1]:.3f} s flag {inf.flag}", file=sys.stderr, flush=True)
        assert inf.flag == 0, inf.flag
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = s.solve(b)
    be = gen.backward_error(n, ptr, row, val, x, b)
    # partial solves: forward then backward must compose to the same solution on every rank
    y = s.solve(b, job=1)
    if not posdef:
        y = s.solve(y, job=2)
    y = s.solve(y, job=3)
    agree = float(np.abs(y - x).max())
    tm = s.timings()
    allt = [None] * world
    dist.all_gather_object(allt, tm["device_s"])
    if rank == 0:
        print(json.dumps(dict(kind=kind, k=k, posdef=posdef, world=world, n=n, num_flops=flops, wall_s=min(times),
                              num_neg=inf.num_neg, num_delay=inf.num_delay, split=s.split_info(),
                              gflops=flops / min(times) / 1e9, device_s_per_rank=allt, bwderr=be,
                              fwd_bwd_vs_full=agree, launches=tm["launches"])), flush=True)
    assert be <= 1e-14, be
    assert agree <= 1e-12
    s.free()                 # numeric trees (and any graph holding NCCL nodes) go before the communicator
    sb.lib().sylver_b200_comm_finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
