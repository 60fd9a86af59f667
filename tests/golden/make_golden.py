"""Regenerates the committed golden fixtures (run from the repo root IN THE BUILD
CONTAINER, where /root/reference exists and oracle/_ref has been built):

    python tests/golden/make_golden.py

symbolic.json : seam arrays from oracle/symbolic.py (restatement of the reference's
                Fortran analysis) for a few small cases.
numeric.json  : outputs of the REAL reference code (oracle/_ref/liboracle.so = SPRAL/SSIDS
                CPU engine compiled unmodified from /root/reference/spral/src): inertia,
                pivot statistics, backward errors and solutions for dense fronts, Laplacian
                and KKT trees.  GPU parity tests compare the B200 engine against these on
                the GPU box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref, symbolic as osym          # noqa: E402
from sylver_b200 import gen                       # noqa: E402
import test_symbolic as ts                        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def symbolic():
    out = {}
    for name, nemin in (("lap7_3", 8), ("rand_60", 32), ("kkt_3", 8), ("simple_mat", 32), ("tridiag3", 32)):
        n, ptr, row, order = ts.CASES[name]
        r = osym.analyse(n, ptr, row, order, nemin=nemin)
        rec = {k: r[k].tolist() for k in ts.KEYS}
        rec.update(nemin=nemin, num_flops=int(r["num_flops"]), num_factor=int(r["num_factor"]))
        out[name] = rec
    json.dump(out, open(os.path.join(HERE, "symbolic.json"), "w"))


def tree_case(kind, k, posdef):
    if kind == "lap7":
        n, ptr, row, val = gen.laplacian_7pt(k)
        order = gen.nested_dissection_order(k)
    elif kind == "lap27":
        n, ptr, row, val = gen.laplacian_27pt(k)
        order = gen.nested_dissection_order(k)
    else:
        n, ptr, row, val = gen.stokes_kkt(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
    sym = osym_fast(n, ptr, row, order)
    ot = ref.OracleTree(sym)
    ot.factor(val, posdef)
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = ot.solve_original(b)
    st = ot.stats
    rec = dict(kind=kind, k=k, posdef=posdef, n=n, flag=st.flag, num_neg=st.num_neg, num_two=st.num_two,
               num_delay=st.num_delay, num_zero=st.num_zero, maxfront=st.maxfront,
               bwderr=gen.backward_error(n, ptr, row, val, x, b))
    ot.close()
    return rec


def osym_fast(n, ptr, row, order):
    """Symbolic arrays through the product's C++ analysis (bit-identical to
    oracle/symbolic.py, see tests/test_symbolic.py) -- the Python restatement is too
    slow for n ~ 1e5."""
    import sylver_b200 as sb
    s = sb.Solver()
    inf = s.analyse(n, ptr, row, order)
    assert inf.flag == 0
    sym = s.symbolic()
    s.free()
    return sym


def dense_case(m, n, delays, seed=1):
    rng = gen.GlibcRand(seed)
    a = gen.dense_sym_indef(m, rng=rng)
    if delays:
        a = gen.cause_delays(a, rng)
    r = ref.factor_front_indef(a, n)
    return dict(m=m, n=n, delays=delays, nelim=int(r["nelim"]), num_neg=int(r["stats"].num_neg),
                num_two=int(r["stats"].num_two), num_delay=int(r["stats"].num_delay),
                not_first_pass=int(r["stats"].not_first_pass), not_second_pass=int(r["stats"].not_second_pass))


def numeric():
    out = {"trees": [], "dense": []}
    for kind, k, posdef in (("lap7", 8, True), ("lap7", 20, True), ("lap27", 12, True), ("lap7", 12, False),
                            ("lap7", 20, False), ("lap27", 16, False), ("kkt", 4, False), ("kkt", 8, False),
                            ("kkt", 12, False), ("kkt", 16, False)):
        out["trees"].append(tree_case(kind, k, posdef))
        print(out["trees"][-1], flush=True)
    for m, n, delays in ((32, 32, False), (64, 64, False), (128, 128, True), (200, 72, False), (500, 500, True),
                         (1000, 300, False), (1000, 300, True), (1301, 517, True), (2048, 1024, False),
                         (2048, 2048, True)):
        out["dense"].append(dense_case(m, n, delays))
        print(out["dense"][-1], flush=True)
    json.dump(out, open(os.path.join(HERE, "numeric.json"), "w"), indent=1)


if __name__ == "__main__":
    symbolic()
    numeric()
