"""The reference's driver (drivers/spldlt_test.F90), against libsylver_b200.so: read a
Rutherford-Boeing matrix, b = A * 1, analyse, factorize, solve, report times and errors.

  python scripts/spldlt_test.py --mat matrix.rb [--posdef | --indef] [--nrhs 1] [--nemin 32]
        [--scale=none|mc64|auction|mc77] [--order=metis|rcm|natural|FILE] [--check] [--u 0.01]
        [--failed-pivot-method=tpp|pass] [--ngpu 1]

Like the Fortran driver the default ordering is METIS (options.ordering = 1, through the static
METIS of the CUDA toolkit); --order can also pick reverse Cuthill-McKee (scipy), the natural
order, or a file with one 1-based position per line.  --ncpu, --nb, --prune-tree,
--sched=*, --*-topology and --gpu-perf-coeff are accepted and ignored (every front runs on the GPU).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb
from sylver_b200 import gen, rb


def pick_order(spec, n, ptr, row):
    if spec in ("metis", "mc64-metis"):
        return None                      # options.ordering = 1 / 2: computed by analyse
    if spec == "natural":
        return np.arange(1, n + 1, dtype=np.int32)
    if spec == "rcm":
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        col = np.repeat(np.arange(n), np.diff(ptr[: n + 1]))
        a = sp.coo_matrix((np.ones(len(col)), (row[: len(col)] - 1, col)), shape=(n, n)).tocsr()
        perm = reverse_cuthill_mckee((a + a.T).tocsr(), symmetric_mode=True)
        order = np.empty(n, dtype=np.int32)
        order[perm] = np.arange(1, n + 1, dtype=np.int32)       # order[i] = position of variable i
        return order
    return np.loadtxt(spec, dtype=np.int64).astype(np.int32)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--mat", default="matrix.rb")
    g = ap.add_mutually_exclusive_group()
    g.add_argument("--posdef", action="store_true")
    g.add_argument("--indef", action="store_true")
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--nemin", type=int, default=32)
    ap.add_argument("--scale", default="none", choices=["none", "mc64", "auction", "mc77", "saved"])
    ap.add_argument("--order", "--ordering", default="metis",
                    help="metis | mc64-metis (matching-based, then --scale=saved is possible) | rcm | natural | FILE")
    ap.add_argument("--check", action="store_true", help="analyse with check=true (matrix cleaning)")
    ap.add_argument("--u", type=float, default=0.01)
    ap.add_argument("--failed-pivot-method", default="tpp", choices=["tpp", "pass"])
    ap.add_argument("--ngpu", type=int, default=1)
    for ignored in ("--ncpu", "--nb", "--gpu-perf-coeff", "--print-level"):
        ap.add_argument(ignored, default=None)
    for ignored in ("--prune-tree", "--no-prune-tree", "--flat-topology", "--numa-topology", "--sched=lws", "--sched=hp"):
        ap.add_argument(ignored, action="store_true")
    a = ap.parse_args(argv)

    print("Reading...")
    m = rb.read(a.mat)
    if m["type"][1] != "s":
        raise SystemExit(f"{a.mat}: type {m['type']} -- a symmetric matrix (lower triangle) is needed")
    n, ptr, row, val = m["n"], m["ptr"], m["row"], m["val"]
    if val is None:      # pattern only: rb_options%values = 2, "make up values" -- diagonally dominant here
        col = np.repeat(np.arange(1, n + 1), np.diff(ptr))
        val = np.where(row == col, float(np.diff(ptr).max() + 1), -1.0)
    print(f"ok  n = {n}  nnz = {int(ptr[n] - 1)}  type = {m['type']}")
    rhs = np.stack([gen.sym_matvec(n, ptr, row, val, np.ones(n)) for _ in range(a.nrhs)], axis=1)

    s = sb.Solver(ngpu=a.ngpu)
    s.options.nemin = a.nemin
    s.options.u = a.u
    s.options.scaling = {"none": 0, "mc64": 1, "auction": 2, "saved": 3, "mc77": 4}[a.scale]
    s.options.failed_pivot_method = 1 if a.failed_pivot_method == "tpp" else 2
    order = pick_order(a.order, n, ptr, row)
    if order is None:
        if sb.metis_order(3, np.array([1, 2, 3, 4]), np.array([1, 2, 3], dtype=np.int32)) is None:
            raise SystemExit("this build has no METIS: pass --order=rcm|natural|FILE")
        s.options.ordering = 2 if a.order == "mc64-metis" else 1
        order = np.zeros(n, dtype=np.int32)
    else:
        s.options.ordering = 0

    t = time.perf_counter()
    inf = s.analyse(n, ptr, row, order, val=val, check=a.check)
    t_analyse = time.perf_counter() - t
    print(f"Analyse: flag {inf.flag}  time {t_analyse:.3f} s  predicted nfact {inf.num_factor:.2e}  nflop {inf.num_flops:.2e}"
          f"  nsuper {inf.num_sup}  maxfront {inf.maxfront}")
    if inf.flag < 0:
        raise SystemExit(1)
    sb.require_gpu()
    t = time.perf_counter()
    inf = s.factorize(val, posdef=a.posdef)
    t_factor = time.perf_counter() - t
    tm = s.timings()
    print(f"Factor:  flag {inf.flag}  time {t_factor:.3f} s (device {tm['device_s']:.3f} s, "
          f"{inf.num_flops / max(tm['device_s'], 1e-12) / 1e9:.1f} GFLOP/s)  delays {inf.num_delay}  "
          f"neg {inf.num_neg}  2x2 {inf.num_two}  rank {inf.matrix_rank}  maxfront {inf.maxfront}")
    if inf.flag < 0:
        raise SystemExit(1)
    t = time.perf_counter()
    x = s.solve(rhs if a.nrhs > 1 else rhs[:, 0])
    t_solve = time.perf_counter() - t
    x2 = x.reshape(n, -1)
    print(f"Solve:   time {t_solve:.3f} s")
    for r in range(a.nrhs):
        print(f"  rhs {r + 1}: forward error {np.abs(x2[:, r] - 1).max():.3e}  "
              f"backward error {gen.backward_error(n, ptr, row, val, x2[:, r], rhs[:, r]):.3e}")
    s.free()


if __name__ == "__main__":
    main()
