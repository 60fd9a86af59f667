"""Tuning aid for gen.stokes_kkt_delays: how many delayed pivots does the reference CPU engine
(oracle/_ref) see for a given fraction of scaled-down variables and exponent range?
Build container only (needs oracle/_ref).  Usage: python scripts/tune_kktd.py k frac lo hi ..."""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
import sylver_b200 as sb  # noqa: E402
from oracle import ref  # noqa: E402
from sylver_b200 import gen  # noqa: E402


def run(k, frac, lo, hi):
    n, ptr, row, val = gen.stokes_kkt(k)
    v = gen.scale_down_some(n, ptr, row, val, frac, lo, hi)
    order = gen.nested_dissection_order(k, dofs_per_cell=4)
    s = sb.Solver()
    s.analyse(n, ptr, row, order)
    sym = s.symbolic()
    s.free()
    ot = ref.OracleTree(sym)
    t = ot.factor(v, False)
    st = ot.stats
    b = gen.sym_matvec(n, ptr, row, v, np.ones(n))
    x = ot.solve_original(b)
    print(k, frac, lo, hi, "n", n, "delay", st.num_delay, "neg", st.num_neg, "two", st.num_two, "maxfront",
          st.maxfront, "bwd", gen.backward_error(n, ptr, row, v, x, b), "t", round(t, 2), flush=True)
    ot.close()


if __name__ == "__main__":
    a = sys.argv[1:]
    for i in range(0, len(a), 4):
        run(int(a[i]), float(a[i + 1]), float(a[i + 2]), float(a[i + 3]))
