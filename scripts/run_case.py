"""Run one synthetic case end to end on the GPU and print timings / parity figures.

usage: python scripts/run_case.py lap27 100 [--indef] [--oracle] [--reps 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb
from sylver_b200 import gen

KNAMES = ["scatter", "zero", "assemble", "potrf", "trsm", "update", "contrib"]


def dense(a):
    """config 2: single dense front m x ncol, random symmetric indefinite (glibc rand, seed 1)."""
    m, n = a.k, (a.ncol or a.k // 4)
    rng = gen.GlibcRand(1)
    A = gen.dense_sym_indef(m, rng=rng)
    if a.delays:
        A = gen.cause_delays(A, rng)
    flops = sum((m - n + j) ** 2 for j in range(1, n + 1))
    for r in range(a.reps):
        res = sb.factor_front_indef(A, n)
        st = res["stats"]
        print(json.dumps(dict(rep=r, m=m, n=n, nelim=res["nelim"], ms=res["ms"], gflops=flops / (res["ms"] * 1e-3) / 1e9,
                              num_neg=st.num_neg, num_two=st.num_two, num_delay=st.num_delay,
                              not_first_pass=st.not_first_pass, not_second_pass=st.not_second_pass)), flush=True)
    if a.oracle:
        from oracle import ref
        t = time.time()
        ro = ref.factor_front_indef(A, n)
        t = time.time() - t
        print(json.dumps(dict(oracle_s=t, oracle_gflops=flops / t / 1e9, nelim=int(ro["nelim"]),
                              num_neg=ro["stats"].num_neg, num_two=ro["stats"].num_two)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["lap7", "lap27", "kkt", "kktd", "dense"])
    ap.add_argument("k", type=int)
    ap.add_argument("--ncol", type=int, default=0, help="dense: fully-summed columns (default m/4)")
    ap.add_argument("--delays", action="store_true", help="dense: apply cause_delays")
    ap.add_argument("--indef", action="store_true")
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--nosolve", action="store_true")
    a = ap.parse_args()
    if a.kind == "dense":
        return dense(a)
    t0 = time.time()
    if a.kind == "lap7":
        n, ptr, row, val = gen.laplacian_7pt(a.k)
        order = gen.nested_dissection_order(a.k)
    elif a.kind == "lap27":
        n, ptr, row, val = gen.laplacian_27pt(a.k)
        order = gen.nested_dissection_order(a.k)
    else:
        n, ptr, row, val = (gen.stokes_kkt if a.kind == "kkt" else gen.stokes_kkt_delays)(a.k)
        order = gen.nested_dissection_order(a.k, dofs_per_cell=4)
        a.indef = True
    t1 = time.time()
    s = sb.Solver()
    inf = s.analyse(n, ptr, row, order)
    t2 = time.time()
    out = dict(kind=a.kind, k=a.k, n=n, nnz=int(ptr[-1] - 1), gen_s=t1 - t0, analyse_s=t2 - t1,
               flag=inf.flag, nnodes=inf.num_sup, num_flops=inf.num_flops, num_factor=inf.num_factor,
               maxfront=inf.maxfront)
    print(json.dumps(out), flush=True)
    posdef = not a.indef
    L = sb.lib()
    for r in range(a.reps):
        tw = time.time()
        inf = s.factorize(val, posdef=posdef)
        tw = time.time() - tw
        tm = s.timings()
        rec = dict(rep=r, flag=inf.flag, py_wall_s=tw, num_neg=inf.num_neg, num_two=inf.num_two,
                   num_delay=inf.num_delay, **(tm or {}))
        if tm:
            rec["gflops"] = out["num_flops"] / tm["device_s"] / 1e9
        print(json.dumps(rec), flush=True)
    tree = L.sylver_b200_fkeep_tree(s.fkeep)
    prof = (C.c_double * 32)()
    L.sylver_b200_numeric_tree_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    nk = L.sylver_b200_numeric_tree_profile(tree, prof, 32) if tree else 0
    if nk > 0:
        tot = sum(prof[3 * i] for i in range(nk))
        for i in range(nk):
            ms, nl, fl = prof[3 * i], prof[3 * i + 1], prof[3 * i + 2]
            print(f"  {KNAMES[i]:9s} {ms:10.3f} ms {100*ms/max(tot,1e-9):5.1f}%  launches {int(nl):6d}  "
                  f"alg GF {fl/1e9:12.2f}  -> {fl/max(ms,1e-9)/1e9:8.2f} TF/s")
    if nk > 0:
        lv = (C.c_double * 4096)()
        nl = L.sylver_b200_numeric_tree_profile_levels(tree, lv, 4096)
        et = s.engine_tree()
        print("  per level (ms): " + " ".join(f"{k:>9s}" for k in KNAMES))
        for l in range(nl):
            row_ = [lv[l * 7 + c] for c in range(7)]
            if sum(row_) > 0.05:
                print(f"  level {l:3d}      " + " ".join(f"{v:9.3f}" for v in row_))
    fb, cb = C.c_long(0), C.c_long(0)
    L.sylver_b200_numeric_tree_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.sylver_b200_numeric_tree_bytes.restype = C.c_long
    if tree:
        L.sylver_b200_numeric_tree_bytes(tree, C.byref(fb), C.byref(cb))
        print(f"  factor arena {fb.value/2**30:.2f} GiB, contribution arena {cb.value/2**30:.2f} GiB")
    if not a.nosolve and inf.flag >= 0:
        x0 = np.ones(n)
        b = gen.sym_matvec(n, ptr, row, val, x0)
        ts = time.time()
        x = s.solve(b)
        ts = time.time() - ts
        print(json.dumps(dict(solve_s=ts, bwderr=gen.backward_error(n, ptr, row, val, x, b),
                              fwderr=float(np.abs(x - 1).max()))), flush=True)
    if a.oracle:
        from oracle import ref
        ot = ref.OracleTree(s.symbolic())
        t = ot.factor(val, posdef)
        xo = ot.solve_original(gen.sym_matvec(n, ptr, row, val, np.ones(n)))
        print(json.dumps(dict(oracle_s=t, oracle_gflops=out["num_flops"] / t / 1e9, num_neg=ot.stats.num_neg,
                              num_two=ot.stats.num_two, num_delay=ot.stats.num_delay,
                              bwderr=gen.backward_error(n, ptr, row, val, xo, gen.sym_matvec(n, ptr, row, val, np.ones(n))))))


if __name__ == "__main__":
    main()
