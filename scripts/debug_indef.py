"""Debug helper: indefinite tree factorization on the GPU, solve re-done in numpy from the
fronts copied back (separates factorization bugs from solve-kernel bugs)."""
import ctypes as C
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylver_b200 as sb
from sylver_b200 import gen


def get_fronts(s, sym):
    L = sb.lib()
    tree = L.sylver_b200_fkeep_tree(s.fkeep)
    out = []
    for f in range(sym["nnodes"]):
        m, n, ne = C.c_int(), C.c_int(), C.c_int()
        L.sylver_b200_numeric_tree_get_front(tree, f, C.byref(m), C.byref(n), None, None)
        l = np.zeros((m.value, n.value), order="F")
        k = m.value - n.value
        cb = np.zeros((max(k, 1), max(k, 1)), order="F")
        L.sylver_b200_numeric_tree_get_front(tree, f, C.byref(m), C.byref(n), l.ctypes.data, cb.ctypes.data)
        d = np.zeros(2 * n.value + 2)
        perm = np.zeros(n.value, dtype=np.int32)
        L.sylver_b200_numeric_tree_get_front_indef(tree, f, C.byref(ne), d.ctypes.data, perm.ctypes.data)
        out.append(dict(m=m.value, n=n.value, nelim=ne.value, L=l, d=d, perm=perm, cb=cb[:k, :k]))
    return out


def numpy_solve(sym, fronts, b_perm, stages=None):
    x = b_perm.copy()
    nn = sym["nnodes"]
    sptr, rptr, rlist = sym["sptr"], sym["rptr"], sym["rlist"]
    def idx(f):
        fr = fronts[f]
        ncol0 = sptr[f + 1] - sptr[f]
        rows = rlist[rptr[f] - 1:rptr[f + 1] - 1]
        return np.concatenate([fr["perm"] - 1, rows[ncol0:] - 1])
    for f in range(nn):
        fr = fronts[f]; ne = fr["nelim"]; ix = idx(f)
        xl = x[ix]
        Lm = np.tril(fr["L"][:, :ne], -1)
        for j in range(ne):
            xl[j + 1:] -= Lm[j + 1:, j] * xl[j]
        x[ix] = xl
    if stages is not None:
        stages.append(x.copy())
    for f in range(nn):
        fr = fronts[f]; ne = fr["nelim"]; ix = idx(f); d = fr["d"]
        i = 0
        while i < ne:
            if i + 1 == ne or np.isfinite(d[2 * i + 2]):
                x[ix[i]] *= d[2 * i]; i += 1
            else:
                a, bb = x[ix[i]], x[ix[i + 1]]
                x[ix[i]] = d[2 * i] * a + d[2 * i + 1] * bb
                x[ix[i + 1]] = d[2 * i + 1] * a + d[2 * i + 3] * bb
                i += 2
    if stages is not None:
        stages.append(x.copy())
    for f in range(nn - 1, -1, -1):
        fr = fronts[f]; ne = fr["nelim"]; ix = idx(f)
        xl = x[ix]
        Lm = np.tril(fr["L"][:, :ne], -1)
        for j in range(ne - 1, -1, -1):
            xl[j] -= Lm[j + 1:, j] @ xl[j + 1:]
        x[ix] = xl
    return x


def main():
    kind, k = sys.argv[1], int(sys.argv[2])
    if kind == "kkt":
        n, ptr, row, val = gen.stokes_kkt(k); order = gen.nested_dissection_order(k, dofs_per_cell=4)
    else:
        n, ptr, row, val = gen.laplacian_7pt(k); order = gen.nested_dissection_order(k)
    s = sb.Solver()
    s.analyse(n, ptr, row, order)
    inf = s.factorize(val, posdef=False)
    sym = s.symbolic()
    print("flag", inf.flag, "neg", inf.num_neg, "two", inf.num_two, "delay", inf.num_delay, "nnodes", sym["nnodes"])
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = s.solve(b)
    print("gpu solve bwderr", gen.backward_error(n, ptr, row, val, x, b))
    fronts = get_fronts(s, sym)
    invp = sym["invp"].astype(np.int64) - 1
    stages = []
    xp = numpy_solve(sym, fronts, b[invp], stages)
    xn = np.empty(n); xn[invp] = xp
    print("numpy solve bwderr", gen.backward_error(n, ptr, row, val, xn, b))
    # per job comparison
    def unperm(v):
        o = np.empty(n); o[invp] = v; return o
    y1 = s.solve(b, job=1)
    print("fwd  gpu-vs-numpy max diff", np.abs(y1 - unperm(stages[0])).max(), "scale", np.abs(stages[0]).max())
    y2 = s.solve(unperm(stages[0]), job=2)
    print("diag gpu-vs-numpy max diff", np.abs(y2 - unperm(stages[1])).max(), "scale", np.abs(stages[1]).max())
    y3 = s.solve(unperm(stages[1]), job=3)
    print("bwd  gpu-vs-numpy max diff", np.abs(y3 - xn).max(), "scale", np.abs(xn).max())
    pos = sym["order"].astype(np.int64) - 1
    dd = np.abs(y1 - unperm(stages[0]))
    print("fwd worst elimination positions", np.sort(pos[np.argsort(-dd)[:12]]), "sptr", sym["sptr"][:6], "nelim", [f["nelim"] for f in fronts][:6])
    bad = np.argsort(-np.abs(x - xn))[:10]
    print("largest gpu-vs-numpy solve diffs at", bad, np.abs(x - xn)[bad])
    for f in range(min(sym["nnodes"], 400)):
        fr = fronts[f]
        if not np.all(np.isfinite(fr["L"])) or fr["nelim"] != fr["n"]:
            print("front", f, fr["m"], fr["n"], fr["nelim"], "nonfinite" if not np.all(np.isfinite(fr["L"])) else "")


if __name__ == "__main__":
    main()
