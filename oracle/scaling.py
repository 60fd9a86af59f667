"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's norm-equilibration scaling.

Only tests/ may import this.  It restates SPRAL's inf_norm_equilib_sym
(/root/reference/spral/src/scaling.f90:480-521, the routine behind options%scaling >= 4,
/root/reference/src/spldlt_factorize_mod.F90:804-831): Algorithm 1 of Knight, Ruiz, Ucar,
"A Symmetry Preserving Algorithm for Matrix Scaling" with equilib_options' defaults
(max_iterations = 10, tol = 1e-8 as a default REAL; scaling.f90:29-32).

Parity: the Fortran cannot be compiled here (no Fortran compiler), so this restatement is
pinned by definition only -- PARITY UNPINNED against a run of the reference; the product's C++
(sylver_b200/csrc/api.cpp::equilib_scale_sym) is compared with it bit for bit
(tests/test_scaling.py), and both are checked against the property the algorithm converges to
(every row of |S A S| has infinity norm 1 within tol once the iteration has converged).
"""
from __future__ import annotations

import numpy as np


def inf_norm_equilib_sym(n: int, ptr: np.ndarray, row: np.ndarray, val: np.ndarray,
                         max_iterations: int = 10, tol: float = float(np.float32(1e-8))):
    """Lower triangle CSC, 1-based ptr/row (scaling.f90:480-521).  Returns (scaling, iterations)."""
    ptr = np.asarray(ptr, dtype=np.int64)
    r = np.asarray(row, dtype=np.int64)[: ptr[n] - 1] - 1
    v = np.asarray(val, dtype=np.float64)[: ptr[n] - 1]
    c = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr[: n + 1]))
    scaling = np.ones(n)                                   # scaling(1:n) = 1.0        (:497)
    itr = 1
    while itr <= max_iterations:                           # do itr = 1, max_iterations (:498)
        a = np.abs(scaling[r] * v * scaling[c])            # abs(scaling(r)*val(j)*scaling(c)) (:505)
        maxentry = np.zeros(n)
        np.maximum.at(maxentry, r, a)                      # maxentry(r) = max(maxentry(r), v) (:506)
        np.maximum.at(maxentry, c, a)                      # maxentry(c) = max(maxentry(c), v) (:507)
        nz = maxentry > 0                                  # beware empty cols          (:511)
        scaling[nz] = scaling[nz] / np.sqrt(maxentry[nz])
        if n == 0 or np.abs(1 - maxentry).max() < tol:     # convergence test           (:514)
            break
        itr += 1
    return scaling, itr - 1                                # inform%iterations = itr-1  (:516)
