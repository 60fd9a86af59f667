"""Argument checking and call-sequence errors of the public C API, after the list the
reference's lineage tests (spral/tests/ssids/ssids.f90:140-400,600-650 -- SyLVER's
spldlt_analyse / spldlt_factorize / spldlt_solve keep SSIDS's flags,
src/sylver_datatypes_mod.F90:13-45).  Everything here returns before any device work, so no GPU
is needed."""
import ctypes as C

import numpy as np
import pytest

import sylver_b200 as sb
from sylver_b200 import gen

SUCCESS, CALL_SEQUENCE, A_N_OOR, A_PTR, A_ALL_OOR = 0, -1, -2, -3, -4
PTR_ROW, ORDER, VAL, X_SIZE, JOB_OOR, NO_SAVED_SCALING, UNIMPLEMENTED = -7, -8, -9, -10, -11, -15, -98


def mat():
    n, ptr, row, val = gen.laplacian_7pt(3)
    return n, ptr, row, val, np.arange(1, n + 1, dtype=np.int32)


def test_analyse_argument_errors(lib):
    n, ptr, row, val, order = mat()
    s = sb.Solver()
    s.options.nemin = -1                                         # nemin out of range: default used
    assert s.analyse(n, ptr, row, order).flag == SUCCESS
    s.options.nemin = 8
    L = sb.lib()

    def raw_analyse(order_ptr, ordering=0, val_ptr=None, n_=n):
        s.options.ordering = ordering
        L.spldlt_analyse(n_, order_ptr, ptr.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p), val_ptr,
                         C.byref(s.akeep), False, C.byref(s.options), C.byref(s.inform))
        s.options.ordering = 0
        return s.inform.flag

    o = order.copy()
    assert raw_analyse(None) == ORDER                            # order absent
    o[0] = n + 1
    assert raw_analyse(o.ctypes.data_as(C.c_void_p)) == ORDER    # order out of range above
    o[0] = 0
    assert raw_analyse(o.ctypes.data_as(C.c_void_p)) == ORDER    # order out of range below
    o = order.copy(); o[1] = o[0]
    assert raw_analyse(o.ctypes.data_as(C.c_void_p)) == ORDER    # repeated entry
    o = order.copy()
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), ordering=-1) == ORDER
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), ordering=3) == ORDER
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), ordering=2) == VAL          # matching ordering needs val
    # METIS ordering: available when the library was built with the CUDA toolkit's static METIS
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), ordering=1) == (SUCCESS if sb.metis_order(n, ptr, row) else UNIMPLEMENTED)
    assert raw_analyse(None, ordering=1) == (SUCCESS if sb.metis_order(n, ptr, row) else UNIMPLEMENTED)   # order may be absent
    val_p = val.ctypes.data_as(C.c_void_p)
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), ordering=2, val_ptr=val_p) == (
        SUCCESS if sb.metis_order(n, ptr, row) else UNIMPLEMENTED)       # matching-based ordering needs METIS too
    assert raw_analyse(o.ctypes.data_as(C.c_void_p), n_=-1) == A_N_OOR
    # factorize after a failed analyse: call sequence error
    assert s.factorize(val, posdef=True).flag == CALL_SEQUENCE
    s.free()


def test_factorize_and_solve_call_sequence(lib):
    n, ptr, row, val, order = mat()
    L = sb.lib()
    s = sb.Solver()
    # factorize without analyse
    L.spldlt_factorize(True, None, None, val.ctypes.data_as(C.c_void_p), None, None, C.byref(s.fkeep),
                       C.byref(s.options), C.byref(s.inform))
    assert s.inform.flag == CALL_SEQUENCE
    assert s.analyse(n, ptr, row, order).flag == SUCCESS
    # solve without factorize
    x = np.ones(n)
    L.spldlt_solve(0, 1, x.ctypes.data_as(C.c_void_p), n, s.akeep, None, C.byref(s.options), C.byref(s.inform))
    assert s.inform.flag == CALL_SEQUENCE
    # val absent
    L.spldlt_factorize(True, None, None, None, None, s.akeep, C.byref(s.fkeep), C.byref(s.options), C.byref(s.inform))
    assert s.inform.flag == VAL
    # a scaling computed at factorize needs ptr and row when the analyse did not check (and keep) them
    for sc in (1, 2, 4):
        s.options.scaling = sc
        L.spldlt_factorize(True, None, row.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p), None, s.akeep,
                           C.byref(s.fkeep), C.byref(s.options), C.byref(s.inform))
        assert s.inform.flag == PTR_ROW
        L.spldlt_factorize(True, ptr.ctypes.data_as(C.c_void_p), None, val.ctypes.data_as(C.c_void_p), None, s.akeep,
                           C.byref(s.fkeep), C.byref(s.options), C.byref(s.inform))
        assert s.inform.flag == PTR_ROW
    s.options.scaling = 3
    assert s.factorize(val, posdef=True).flag == NO_SAVED_SCALING
    s.free()


def test_n_zero(lib):
    """n = 0 (ssids.f90:1172-1190): every phase succeeds and does nothing."""
    s = sb.Solver()
    ptr = np.array([1], dtype=np.int64)
    row = np.zeros(1, dtype=np.int32)
    inf = s.analyse(0, ptr, row, np.zeros(1, dtype=np.int32))
    assert inf.flag == SUCCESS
    inf = s.factorize(np.zeros(1), posdef=False)
    assert inf.flag == SUCCESS and inf.matrix_rank == 0
    s.free()
