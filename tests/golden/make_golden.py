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
    elif kind == "kktd":
        n, ptr, row, val = gen.stokes_kkt_delays(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
    else:
        n, ptr, row, val = gen.stokes_kkt(k)
        order = gen.nested_dissection_order(k, dofs_per_cell=4)
    sym = osym_fast(n, ptr, row, order)
    ot = ref.OracleTree(sym)
    secs = ot.factor(val, posdef)
    b = gen.sym_matvec(n, ptr, row, val, np.ones(n))
    x = ot.solve_original(b)
    st = ot.stats
    rec = dict(kind=kind, k=k, posdef=posdef, n=n, flag=st.flag, num_neg=st.num_neg, num_two=st.num_two,
               num_delay=st.num_delay, num_zero=st.num_zero, maxfront=st.maxfront,
               bwderr=gen.backward_error(n, ptr, row, val, x, b), oracle_seconds=secs)
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


def numeric_big():
    """numeric_big.json: the BASELINE.json configurations themselves (or the largest stand-in the
    CPU engine finishes in minutes here): config 2 (dense 8192 x 2048 APTP), config 3
    (27-point Laplacian 100^3 posdef), config 4's matrix family at 40^3 with and without
    delay-causing scaling, and 7-point LDL^T at 60^3.  Minutes of CPU time; run with `big`."""
    out = {"trees": [], "dense": []}
    path = os.path.join(HERE, "numeric_big.json")
    for m, n, delays in ((8192, 2048, False), (8192, 2048, True)):
        out["dense"].append(dense_case(m, n, delays))
        print(out["dense"][-1], flush=True)
    for kind, k, posdef in (("kktd", 12, False), ("kktd", 24, False), ("kkt", 40, False), ("kktd", 32, False),
                            ("lap7", 60, False), ("lap27", 100, True), ("kktd", 40, False)):
        out["trees"].append(tree_case(kind, k, posdef))
        print(out["trees"][-1], flush=True)
        json.dump(out, open(path, "w"), indent=1)


def host_preprocessing():
    """preprocess.json: outputs of the restatements of SPRAL's Fortran pre-processing
    (oracle/scaling.py, oracle/matrix_clean.py) on fixed inputs, floats as hex strings: the three
    scalings (vector, matching, counters) and the cleaned matrix + conversion map.  The product's
    C++ must reproduce them bit for bit (tests/test_scaling.py, tests/test_matrix_clean.py)."""
    from oracle import scaling as oscal, matrix_clean as oclean
    import test_scaling as tsc
    import test_matrix_clean as tmc
    out = {"scaling": [], "clean": []}
    for kind, k, seed in (("lap7", 5, 1), ("kkt", 4, 3)):
        n, ptr, row, val = tsc.badly_scaled(kind, k, seed)
        rec = dict(kind=kind, k=k, seed=seed)
        s, it = oscal.inf_norm_equilib_sym(n, ptr, row, val)
        rec["equilib"] = dict(scaling=[float(x).hex() for x in s], iterations=int(it))
        s, m, inf = oscal.auction_scale_sym(n, ptr, row, val)
        rec["auction"] = dict(scaling=[float(x).hex() for x in s], match=m.tolist(), inform=inf)
        s, m, inf = oscal.hungarian_scale_sym(n, ptr, row, val)
        rec["hungarian"] = dict(scaling=[float(x).hex() for x in s], match=m.tolist(), inform=inf)
        out["scaling"].append(rec)
    for seed in (501, 507):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(10, 40))
        ptr, row, val = tmc.dirty_random(n, rng)
        c = oclean.clean_cscl_oop_sym_indef(n, ptr, row)
        out["clean"].append(dict(seed=seed, n=n, ptr_in=ptr.tolist(), row_in=row.tolist(),
                                 val_in=[float(x).hex() for x in val], flag=c["flag"], noor=c["noor"], ndup=c["ndup"],
                                 ptr=c["ptr"].tolist(), row=c["row"].tolist(), map=c["map"].tolist(),
                                 val=[float(x).hex() for x in oclean.apply_conversion_map(c, val)]))
    json.dump(out, open(os.path.join(HERE, "preprocess.json"), "w"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        host_preprocessing()
    elif len(sys.argv) > 1 and sys.argv[1] == "big":
        numeric_big()
    else:
        symbolic()
        numeric()
        host_preprocessing()
