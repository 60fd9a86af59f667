"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY (checker, never the product path).

ctypes driver for ``oracle/_ref/liboracle.so``: the UNMODIFIED SPRAL/SSIDS CPU
multifrontal engine compiled from /root/reference/spral/src (recipe:
oracle/Makefile).  It is the code SyLVER itself delegates subtrees to
(/root/reference/src/spldlt_factorize_mod.F90:395-471) and shares SyLVER's
pivoting kernels, so it stands in for the (unbuildable) full reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle.so")


class CpuFactorOptions(C.Structure):
    """spral::ssids::cpu::cpu_factor_options (spral/src/ssids/cpu/cpu_iface.hxx:23-33)."""
    _fields_ = [("print_level", C.c_int), ("action", C.c_bool), ("small", C.c_double),
                ("u", C.c_double), ("multiplier", C.c_double),
                ("small_subtree_threshold", C.c_long), ("cpu_block_size", C.c_int),
                ("pivot_method", C.c_int), ("failed_pivot_method", C.c_int)]


class ThreadStats(C.Structure):
    """spral::ssids::cpu::ThreadStats (spral/src/ssids/cpu/ThreadStats.hxx:47-58)."""
    _fields_ = [("flag", C.c_int), ("num_delay", C.c_int), ("num_neg", C.c_int),
                ("num_two", C.c_int), ("num_zero", C.c_int), ("maxfront", C.c_int),
                ("not_first_pass", C.c_int), ("not_second_pass", C.c_int)]


def default_options() -> CpuFactorOptions:
    """SyLVER defaults (src/sylver_datatypes_mod.F90:97-198; cpu_block_size forced to
    256 for subtrees, src/tasks/tasks.hxx:503)."""
    o = CpuFactorOptions()
    o.print_level = 0
    o.action = True
    o.small = 1e-20
    o.u = 0.01
    o.multiplier = 1.1
    o.small_subtree_threshold = 4 * 10 ** 6
    o.cpu_block_size = 256
    o.pivot_method = 2
    o.failed_pivot_method = 1
    return o


_lib = None


OMP_LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_omp.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def select_task_parallel() -> bool:
    """CPU-baseline timing only: load liboracle_omp.so (the same sources built with OpenMP tasks,
    `make -C oracle omp`) instead of the sequential library.  Must be called before the first
    lib(); the caller sets OMP_NUM_THREADS / OPENBLAS_NUM_THREADS.  Returns False if that build
    is absent (the sequential library is used then)."""
    global LIB_PATH
    if _lib is not None or not os.path.exists(OMP_LIB_PATH):
        return False
    LIB_PATH = OMP_LIB_PATH
    return True


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        # the wheel-bundled OpenBLAS needs its own libgfortran: preload from the same dir
        pth = os.path.join(_HERE, "_ref", "openblas_path.txt")
        if os.path.exists(pth):
            ob = open(pth).read().strip()
            d = os.path.dirname(ob)
            for prefix in ("libquadmath", "libgfortran", "libopenblas"):
                for f in sorted(os.listdir(d)):
                    if f.startswith(prefix):
                        try:
                            C.CDLL(os.path.join(d, f), mode=C.RTLD_GLOBAL)
                        except OSError:
                            pass
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.spral_ssids_cpu_create_symbolic_subtree.restype = vp
        L.spral_ssids_cpu_create_symbolic_subtree.argtypes = [
            C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp,
            C.POINTER(CpuFactorOptions)]
        L.spral_ssids_cpu_destroy_symbolic_subtree.argtypes = [vp]
        L.spral_ssids_cpu_create_num_subtree_dbl.restype = vp
        L.spral_ssids_cpu_create_num_subtree_dbl.argtypes = [
            C.c_bool, vp, vp, vp, vp, C.POINTER(CpuFactorOptions), C.POINTER(ThreadStats)]
        L.spral_ssids_cpu_destroy_num_subtree_dbl.argtypes = [C.c_bool, vp]
        L.oracle_create_num_subtree_parallel.restype = vp
        L.oracle_create_num_subtree_parallel.argtypes = [
            C.c_bool, vp, vp, vp, C.POINTER(CpuFactorOptions), C.POINTER(ThreadStats)]
        for nm in ("fwd", "diag", "diag_bwd", "bwd"):
            f = getattr(L, f"spral_ssids_cpu_subtree_solve_{nm}_dbl")
            f.argtypes = [C.c_bool, vp, C.c_int, vp, C.c_int]
            f.restype = C.c_int
        L.spral_ssids_cpu_subtree_enquire_dbl.argtypes = [C.c_bool, vp, vp, vp]
        L.oracle_align_lda.restype = C.c_long
        L.oracle_align_lda.argtypes = [C.c_long]
        L.oracle_factor_front_indef.restype = C.c_int
        L.oracle_factor_front_indef.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp,
                                                C.POINTER(CpuFactorOptions), C.POINTER(ThreadStats)]
        L.oracle_factor_front_posdef.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int,
                                                 C.POINTER(C.c_int)]
        L.oracle_ldlt_tpp_factor.restype = C.c_int
        L.oracle_ldlt_tpp_factor.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int,
                                             C.c_bool, C.c_double, C.c_double]
        L.oracle_ldlt_solve.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int]
        L.oracle_chol_solve.argtypes = [C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleTree:
    """Whole assembly tree factorized as ONE SSIDS CPU subtree
    (spral/src/ssids/cpu/SymbolicSubtree.cxx:10-20, NumericSubtree.cxx:29-59)."""

    def __init__(self, sym: dict, options: CpuFactorOptions | None = None, last_node: int | None = None):
        """last_node: factorize only the nodes 1..last_node (a prefix of the postorder = a forest
        of complete subtrees; bounded CPU samples of a large problem -- no solve then)."""
        self.L = lib()
        self.sym = sym
        self.opt = options or default_options()
        self.n = int(sym["n"])
        nn = int(sym["nnodes"]) if last_node is None else int(last_node)
        # keep the arrays alive: SSIDS borrows the pointers
        self._keep = [np.ascontiguousarray(sym[k]) for k in ("sptr", "sparent", "rptr", "rlist", "nptr", "nlist")]
        self._contrib_idx = np.zeros(1, dtype=np.int32)
        self.symb = self.L.spral_ssids_cpu_create_symbolic_subtree(
            self.n, 1, nn + 1, *[_p(a) for a in self._keep], 0, _p(self._contrib_idx), C.byref(self.opt))
        self.num = None
        self.posdef = None
        self.stats = ThreadStats()

    def factor(self, val: np.ndarray, posdef: bool, scaling=None) -> float:
        if self.num is not None:
            self.L.spral_ssids_cpu_destroy_num_subtree_dbl(self.posdef, self.num)
        val = np.ascontiguousarray(val, dtype=np.float64)
        self._val = val
        self.posdef = posdef
        t0 = time.perf_counter()
        if LIB_PATH == OMP_LIB_PATH:
            # task-parallel build: the OpenMP parallel/single region SSIDS's Fortran caller provides
            self.num = self.L.oracle_create_num_subtree_parallel(
                posdef, self.symb, _p(val), _p(scaling), C.byref(self.opt), C.byref(self.stats))
        else:
            self.num = self.L.spral_ssids_cpu_create_num_subtree_dbl(
                posdef, self.symb, _p(val), _p(scaling), None, C.byref(self.opt), C.byref(self.stats))
        return time.perf_counter() - t0

    def solve(self, b_perm: np.ndarray) -> np.ndarray:
        """Solve with x already in elimination order (as the seam expects)."""
        x = np.array(b_perm, dtype=np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        self.L.spral_ssids_cpu_subtree_solve_fwd_dbl(self.posdef, self.num, nrhs, _p(x), self.n)
        self.L.spral_ssids_cpu_subtree_solve_diag_bwd_dbl(self.posdef, self.num, nrhs, _p(x), self.n) \
            if not self.posdef else \
            self.L.spral_ssids_cpu_subtree_solve_bwd_dbl(self.posdef, self.num, nrhs, _p(x), self.n)
        return x

    def solve_original(self, b: np.ndarray) -> np.ndarray:
        invp = np.asarray(self.sym["invp"], dtype=np.int64) - 1
        xp = self.solve(b[invp])
        x = np.empty_like(xp)
        x[invp] = xp
        return x

    def close(self):
        if self.num is not None:
            self.L.spral_ssids_cpu_destroy_num_subtree_dbl(self.posdef, self.num)
            self.num = None
        if self.symb is not None:
            self.L.spral_ssids_cpu_destroy_symbolic_subtree(self.symb)
            self.symb = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def aligned_zeros(shape, align: int = 64) -> np.ndarray:
    """Fortran-ordered float64 zeros whose data pointer is `align`-byte aligned.  SSIDS picks
    block_ldlt (aligned 32x32 tile) or ldlt_tpp_factor per block by pointer alignment
    (spral/src/ssids/cpu/kernels/ldlt_app.cxx, Block::factor), so pivot sequences -- hence
    num_two -- are only reproducible with a fixed alignment."""
    cnt = int(np.prod(shape))
    raw = np.zeros(cnt + align // 8, dtype=np.float64)
    off = (-raw.ctypes.data % align) // 8
    return raw[off:off + cnt].reshape(shape, order="F")


def align_lda(m: int) -> int:
    return int(lib().oracle_align_lda(m))


def factor_front_posdef(a: np.ndarray, n: int, blksz: int = 256):
    """cholesky_factor on an m x m symmetric array's first n columns.
    Returns (L panel m x n, contrib (m-n)^2, info) -- info == -1 on success."""
    m = a.shape[0]
    lda = align_lda(m)
    buf = aligned_zeros((lda, n))
    buf[:m, :] = np.tril(a)[:, :n]
    k = m - n
    contrib = aligned_zeros((max(k, 1), max(k, 1)))
    info = C.c_int(0)
    lib().oracle_factor_front_posdef(m, n, _p(buf), lda, _p(contrib), blksz, C.byref(info))
    return buf[:m, :].copy(), contrib[:k, :k].copy(), info.value


def factor_front_indef(a: np.ndarray, n: int, options: CpuFactorOptions | None = None):
    """factor_node_indef on a dense front. Returns dict(nelim, L, d, perm, contrib, stats)."""
    opt = options or default_options()
    m = a.shape[0]
    lda = align_lda(m)
    buf = aligned_zeros((lda, n))
    buf[:m, :] = np.tril(a)[:, :n]
    d = np.zeros(2 * n + 2)
    perm = np.arange(1, n + 1, dtype=np.int32)
    k = m - n
    contrib = aligned_zeros((max(k, 1), max(k, 1)))
    # the contribution block must hold A22 (the reference front code receives it zeroed and
    # applies the Schur update with beta = 0; add A22 afterwards in the caller)
    stats = ThreadStats()
    nelim = lib().oracle_factor_front_indef(m, n, _p(perm), _p(buf), lda, _p(d), _p(contrib),
                                            C.byref(opt), C.byref(stats))
    return dict(nelim=nelim, L=buf[:m, :].copy(), d=d[:2 * n].copy(), perm=perm, contrib=contrib[:k, :k].copy(),
                stats=stats, lda=lda)
