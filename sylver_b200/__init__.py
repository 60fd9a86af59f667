"""sylver_b200 -- B200-native numeric factorization behind SyLVER's C ABI.

This package is a thin ctypes binding over ``libsylver_b200.so`` (built in-tree
by :mod:`sylver_b200.build`).  The Python layer mirrors the reference's C API
one to one (``spldlt_analyse / spldlt_factorize / spldlt_solve``,
/root/reference/include/sylver/sylver.h:73-123); all numeric work happens in
the hand-written sm_100a kernels inside the shared library.  There is no CPU
fallback: numeric calls on a machine without a CUDA device return
``flag == -51`` and :func:`require_gpu` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SYLVER_B200_LIB selects a kernel-variant build for A/B measurements (sylver_b200/build.py)
LIB_PATH = os.environ.get("SYLVER_B200_LIB") or os.path.join(_HERE, "libsylver_b200.so")


class Inform(C.Structure):
    _fields_ = [("flag", C.c_int), ("matrix_dup", C.c_int), ("matrix_missing_diag", C.c_int),
                ("matrix_outrange", C.c_int), ("matrix_rank", C.c_int), ("maxdepth", C.c_int),
                ("maxfront", C.c_int), ("num_delay", C.c_int), ("num_factor", C.c_long),
                ("num_flops", C.c_long), ("num_neg", C.c_int), ("num_sup", C.c_int),
                ("num_two", C.c_int), ("stat", C.c_int), ("cuda_error", C.c_int),
                ("cublas_error", C.c_int), ("unused", C.c_char * 80)]


class Options(C.Structure):
    _fields_ = [("array_base", C.c_int), ("print_level", C.c_int), ("unit_diagnostics", C.c_int),
                ("unit_error", C.c_int), ("unit_warning", C.c_int), ("ordering", C.c_int),
                ("nemin", C.c_int), ("prune_tree", C.c_bool), ("min_gpu_work", C.c_long),
                ("scaling", C.c_int), ("pivot_method", C.c_int), ("small", C.c_double),
                ("u", C.c_double), ("small_subtree_threshold", C.c_long), ("nb", C.c_int),
                ("cpu_topology", C.c_int), ("action", C.c_bool), ("use_gpu", C.c_bool),
                ("gpu_perf_coeff", C.c_double), ("failed_pivot_method", C.c_int),
                ("scheduler", C.c_int)]


class OptionsC(C.Structure):
    """sylver::options_c (reference src/sylver_ciface.hxx:39-55)."""
    _fields_ = [("print_level", C.c_int), ("action", C.c_bool), ("small", C.c_double),
                ("u", C.c_double), ("multiplier", C.c_double),
                ("small_subtree_threshold", C.c_long), ("nb", C.c_int),
                ("pivot_method", C.c_int), ("failed_pivot_method", C.c_int),
                ("cpu_topology", C.c_int)]


class InformC(C.Structure):
    """sylver::inform_c (reference src/sylver_ciface.hxx:72-87)."""
    _fields_ = [("flag", C.c_int), ("num_delay", C.c_int), ("num_neg", C.c_int),
                ("num_two", C.c_int), ("num_zero", C.c_int), ("maxfront", C.c_int),
                ("not_first_pass", C.c_int), ("not_second_pass", C.c_int)]


class SymbolicView(C.Structure):
    _fields_ = [("n", C.c_int), ("nnodes", C.c_int), ("sptr", C.POINTER(C.c_int)),
                ("sparent", C.POINTER(C.c_int)), ("rptr", C.POINTER(C.c_long)),
                ("rlist", C.POINTER(C.c_int)), ("nptr", C.POINTER(C.c_long)),
                ("nlist", C.POINTER(C.c_long)), ("order", C.POINTER(C.c_int)),
                ("invp", C.POINTER(C.c_int)), ("num_factor", C.c_long), ("num_flops", C.c_long)]


_lib = None

# every symbol include/sylver_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "sylver_init", "sylver_finalize", "sylver_default_options", "spldlt_analyse",
    "spldlt_factorize", "spldlt_solve", "spldlt_free_akeep", "spldlt_free_fkeep",
    "spldlt_create_symbolic_tree", "spldlt_destroy_symbolic_tree",
    "spldlt_create_numeric_tree_dbl", "spldlt_create_numeric_tree_posdef_dbl",
    "spldlt_destroy_numeric_tree_dbl", "spldlt_destroy_numeric_tree_posdef_dbl",
    "spldlt_tree_solve_fwd_dbl", "spldlt_tree_solve_bwd_dbl", "spldlt_tree_solve_diag_dbl",
    "spldlt_tree_solve_diag_bwd_dbl", "spldlt_tree_solve_fwd_posdef_dbl",
    "spldlt_tree_solve_bwd_posdef_dbl", "sylver_b200_device_count", "sylver_b200_version",
    "sylver_b200_akeep_view", "sylver_b200_symbolic_tree_cmap", "sylver_b200_numeric_tree_timings", "sylver_b200_numeric_tree_split_info",
    "sylver_b200_symbolic_tree_view",
    "sylver_b200_fkeep_tree", "sylver_b200_factor_front_posdef", "sylver_b200_factor_front_indef",
    "sylver_b200_bench_dmma", "sylver_b200_bench_copy", "sylver_b200_akeep_tree",
    "sylver_b200_numeric_tree_profile", "sylver_b200_numeric_tree_profile_levels", "sylver_b200_numeric_tree_bytes", "sylver_b200_set_stream",
    "sylver_b200_numeric_tree_get_front", "sylver_b200_numeric_tree_get_front_indef",
    "sylver_b200_comm_unique_id", "sylver_b200_comm_init", "sylver_b200_comm_finalize",
    "sylver_b200_comm_rank", "sylver_b200_comm_world", "sylver_b200_comm_set_virtual",
    "sylver_b200_comm_init_local",
    "sylver_b200_partition", "sylver_b200_plan_exchanges", "sylver_b200_plan_split", "sylver_b200_plan_levels", "sylver_b200_metis_order", "sylver_b200_match_order", "sylver_b200_equilib_scale", "sylver_b200_auction_scale", "sylver_b200_hungarian_scale", "sylver_b200_clean_matrix", "sylver_b200_apply_conversion_map",
]


def lib() -> C.CDLL:
    """Load libsylver_b200.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m sylver_b200.build` (or __graft_entry__.build()). "
            "There is no CPU/PyTorch fallback for the factorization path.")
    L = C.CDLL(LIB_PATH)
    vp, ip, lp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_double)
    L.sylver_init.argtypes = [C.c_int, C.c_int]
    L.sylver_default_options.argtypes = [C.POINTER(Options)]
    L.spldlt_analyse.argtypes = [C.c_int, vp, vp, vp, vp, C.POINTER(vp), C.c_bool,
                                 C.POINTER(Options), C.POINTER(Inform)]
    L.spldlt_factorize.argtypes = [C.c_bool, vp, vp, vp, vp, vp, C.POINTER(vp),
                                   C.POINTER(Options), C.POINTER(Inform)]
    L.spldlt_solve.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp, C.POINTER(Options),
                               C.POINTER(Inform)]
    L.spldlt_free_akeep.argtypes = [C.POINTER(vp)]
    L.spldlt_free_fkeep.argtypes = [C.POINTER(vp)]
    L.spldlt_create_symbolic_tree.restype = vp
    L.spldlt_create_symbolic_tree.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                              C.c_int, vp, vp, vp, vp]
    L.spldlt_destroy_symbolic_tree.argtypes = [vp]
    L.spldlt_create_numeric_tree_dbl.restype = vp
    L.spldlt_create_numeric_tree_dbl.argtypes = [C.c_bool, vp, vp, vp, vp, vp,
                                                 C.POINTER(OptionsC), C.POINTER(InformC)]
    L.spldlt_create_numeric_tree_posdef_dbl.restype = vp
    L.spldlt_create_numeric_tree_posdef_dbl.argtypes = [vp, vp, vp, vp, vp, C.POINTER(OptionsC),
                                                        C.POINTER(InformC)]
    L.spldlt_destroy_numeric_tree_dbl.argtypes = [C.c_bool, vp]
    L.spldlt_destroy_numeric_tree_posdef_dbl.argtypes = [vp]
    for name in ("fwd", "bwd", "diag", "diag_bwd"):
        f = getattr(L, f"spldlt_tree_solve_{name}_dbl")
        f.argtypes = [C.c_bool, vp, C.c_int, vp, C.c_int]
        f.restype = C.c_int
    for name in ("fwd", "bwd"):
        f = getattr(L, f"spldlt_tree_solve_{name}_posdef_dbl")
        f.argtypes = [vp, C.c_int, vp, C.c_int]
        f.restype = C.c_int
    L.sylver_b200_version.restype = C.c_char_p
    L.sylver_b200_akeep_view.argtypes = [vp, C.POINTER(SymbolicView)]
    L.sylver_b200_akeep_tree.restype = vp
    L.sylver_b200_akeep_tree.argtypes = [vp]
    L.sylver_b200_symbolic_tree_cmap.argtypes = [vp, C.POINTER(lp), C.POINTER(ip)]
    L.sylver_b200_fkeep_tree.restype = vp
    L.sylver_b200_fkeep_tree.argtypes = [vp]
    L.sylver_b200_numeric_tree_timings.argtypes = [vp, dp]
    L.sylver_b200_numeric_tree_split_info.argtypes = [vp, vp]
    L.sylver_b200_symbolic_tree_view.argtypes = [vp, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(ip), C.POINTER(ip)]
    L.sylver_b200_numeric_tree_get_front.argtypes = [vp, C.c_int, ip, ip, vp, vp]
    L.sylver_b200_factor_front_posdef.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_int,
                                                  C.POINTER(C.c_float)]
    L.sylver_b200_factor_front_indef.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp,
                                                 C.POINTER(OptionsC), C.POINTER(InformC),
                                                 C.POINTER(C.c_float)]
    L.sylver_b200_numeric_tree_get_front_indef.argtypes = [vp, C.c_int, ip, vp, vp]
    L.sylver_b200_numeric_tree_profile.argtypes = [vp, dp, C.c_int]
    L.sylver_b200_numeric_tree_profile_levels.argtypes = [vp, dp, C.c_int]
    L.sylver_b200_numeric_tree_bytes.restype = C.c_long
    L.sylver_b200_numeric_tree_bytes.argtypes = [vp, lp, lp]
    L.sylver_b200_set_stream.argtypes = [vp, C.c_int]
    L.sylver_b200_comm_unique_id.argtypes = [vp]
    L.sylver_b200_comm_init.argtypes = [C.c_int, C.c_int, vp]
    L.sylver_b200_comm_set_virtual.argtypes = [C.c_int, C.c_int]
    L.sylver_b200_comm_init_local.argtypes = [C.c_int, C.c_int, C.c_int]
    L.sylver_b200_partition.argtypes = [vp, C.c_int, vp]
    L.sylver_b200_plan_exchanges.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp]
    L.sylver_b200_plan_split.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.sylver_b200_metis_order.argtypes = [C.c_int, vp, vp, vp, vp]
    L.sylver_b200_match_order.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp]
    L.sylver_b200_plan_levels.argtypes = [vp, C.c_int, C.c_int, C.c_long, vp]
    L.sylver_b200_plan_levels.restype = C.c_long
    L.sylver_b200_equilib_scale.argtypes = [C.c_int, vp, vp, vp, vp]
    L.sylver_b200_auction_scale.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp]
    L.sylver_b200_hungarian_scale.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_int, vp]
    L.sylver_b200_clean_matrix.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    L.sylver_b200_apply_conversion_map.argtypes = [C.c_long, C.c_long, vp, vp, vp]
    L.sylver_b200_bench_dmma.restype = C.c_double
    L.sylver_b200_bench_dmma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.sylver_b200_bench_copy.restype = C.c_double
    L.sylver_b200_bench_copy.argtypes = [C.c_long, C.c_int]
    _lib = L
    return L


def device_count() -> int:
    return int(lib().sylver_b200_device_count())


def require_gpu() -> None:
    if device_count() == 0:
        raise RuntimeError("sylver_b200: no CUDA device visible; the factorization path has no CPU fallback")


def default_options() -> Options:
    o = Options()
    lib().sylver_default_options(C.byref(o))
    o.ordering = 0     # the order is always an input here (METIS is not part of this path)
    return o


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)       # raw (device) address
    return a.ctypes.data_as(C.c_void_p)


def comm_init_from_torch(dist, device_index: int) -> None:
    """Create the library's NCCL communicator inside a torch.distributed job (one process
    per GPU): rank 0 draws the NCCL unique id, torch.distributed broadcasts its 128 bytes."""
    import torch
    L = lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        if L.sylver_b200_comm_unique_id(buf) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.to(torch.device("cuda", device_index))
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
    if L.sylver_b200_comm_init(rank, world, idbuf) != 0:
        raise RuntimeError("sylver_b200_comm_init failed")


_fabric_seq = [0]


def run_local_ranks(world: int, fn, timeout: float = 600.0):
    """Run ``fn(rank, world)`` on `world` threads of this process, each joined to a fresh
    in-process communicator (sylver_b200_comm_init_local): the multi-rank factorization and
    solve on ONE device.  Returns the list of results; re-raises the first exception."""
    import threading
    L = lib()
    _fabric_seq[0] += 1
    fid = _fabric_seq[0]
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            if L.sylver_b200_comm_init_local(r, world, fid) != 0:
                raise RuntimeError("sylver_b200_comm_init_local failed")
            try:
                out[r] = fn(r, world)
            finally:
                L.sylver_b200_comm_finalize()
        except BaseException as e:      # noqa: BLE001 - reported to the caller below
            err[r] = e

    ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout)
        if t.is_alive():
            raise TimeoutError("a local rank did not finish (deadlocked exchange?)")
    for e in err:
        if e is not None:
            raise e
    return out


def partition(solver: "Solver", world: int) -> np.ndarray:
    nn = solver.symbolic()["nnodes"]
    own = np.zeros(max(nn, 1), dtype=np.int32)
    lib().sylver_b200_partition(solver.akeep, world, _ptr(own))
    return own[:nn]


def equilib_scale(n: int, ptr, row, val):
    """Norm-equilibration scaling of a lower-triangle CSC matrix (sylver_b200_equilib_scale):
    (scaling, iterations)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    sc = np.zeros(max(n, 1))
    it = lib().sylver_b200_equilib_scale(n, _ptr(ptr), _ptr(row), _ptr(val), _ptr(sc))
    if it < 0:
        raise RuntimeError("sylver_b200_equilib_scale failed")
    return sc[:n], it


def auction_scale(n: int, ptr, row, val):
    """Auction-matching scaling of a lower-triangle CSC matrix (sylver_b200_auction_scale):
    (scaling, match, inform dict)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    sc = np.zeros(max(n, 1))
    match = np.zeros(max(n, 1), dtype=np.int32)
    inf = np.zeros(4, dtype=np.int32)
    if lib().sylver_b200_auction_scale(n, _ptr(ptr), _ptr(row), _ptr(val), _ptr(sc), _ptr(match), _ptr(inf)) != 0:
        raise RuntimeError("sylver_b200_auction_scale failed")
    return sc[:n], match[:n], dict(flag=int(inf[0]), matched=int(inf[1]), iterations=int(inf[2]), unmatchable=int(inf[3]))


def clean_matrix(n: int, ptr, row):
    """What analyse(check=True) does to the structure (sylver_b200_clean_matrix):
    dict(flag, noor, ndup, ptr, row, map)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    cap = max(len(row), 1)
    po = np.zeros(max(n, 0) + 1, dtype=np.int64)
    ro = np.zeros(cap, dtype=np.int32)
    mp = np.zeros(2 * cap, dtype=np.int64)
    cnt = np.zeros(5, dtype=np.int64)
    flag = lib().sylver_b200_clean_matrix(n, _ptr(ptr), _ptr(row), cap, _ptr(po), _ptr(ro), _ptr(mp), _ptr(cnt))
    if flag < 0:
        return dict(flag=flag)
    return dict(flag=flag, noor=int(cnt[1]), ndup=int(cnt[2]), ptr=po, row=ro[: cnt[3]], map=mp[: cnt[4]])


def apply_conversion_map(cm: dict, val) -> np.ndarray:
    """Values of the cleaned matrix `cm` (from clean_matrix) for the caller's values."""
    val = np.ascontiguousarray(val, dtype=np.float64)
    mp = np.ascontiguousarray(cm["map"], dtype=np.int64)
    ne = len(cm["row"])
    out = np.zeros(max(ne, 1))
    if lib().sylver_b200_apply_conversion_map(ne, len(mp), _ptr(mp), _ptr(val), _ptr(out)) != 0:
        raise RuntimeError("sylver_b200_apply_conversion_map failed")
    return out[:ne]


def hungarian_scale(n: int, ptr, row, val, scale_if_singular: bool = False):
    """Hungarian (MC64-like) scaling of a lower-triangle CSC matrix
    (sylver_b200_hungarian_scale): (scaling, match, inform dict)."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    sc = np.zeros(max(n, 1))
    match = np.zeros(max(n, 1), dtype=np.int32)
    inf = np.zeros(2, dtype=np.int32)
    flag = lib().sylver_b200_hungarian_scale(n, _ptr(ptr), _ptr(row), _ptr(val), _ptr(sc), _ptr(match),
                                             1 if scale_if_singular else 0, _ptr(inf))
    if flag == -1:
        raise RuntimeError("sylver_b200_hungarian_scale failed")
    return sc[:n], match[:n], dict(flag=int(inf[0]), matched=int(inf[1]))


def metis_order(n: int, ptr, row):
    """METIS nested-dissection order of a lower-triangle CSC pattern (sylver_b200_metis_order):
    (order, invp), 1-based; None if the library was built without METIS."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    order = np.zeros(max(n, 1), dtype=np.int32)
    invp = np.zeros(max(n, 1), dtype=np.int32)
    rc = lib().sylver_b200_metis_order(n, _ptr(ptr), _ptr(row), _ptr(order), _ptr(invp))
    if rc == -2:
        return None
    if rc != 0:
        raise RuntimeError(f"sylver_b200_metis_order failed ({rc})")
    return order[:n], invp[:n]


def match_order(n: int, ptr, row, val):
    """Matching-based ordering (sylver_b200_match_order): (flag, order, scale, pairs); None if
    the library was built without METIS."""
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    row = np.ascontiguousarray(row, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    order = np.zeros(max(n, 1), dtype=np.int32)
    scale = np.zeros(max(n, 1))
    pairs = np.zeros(max(n, 1), dtype=np.int32)
    rc = lib().sylver_b200_match_order(n, _ptr(ptr), _ptr(row), _ptr(val), _ptr(order), _ptr(scale), _ptr(pairs))
    if rc == -2:
        return None
    if rc < 0:
        raise RuntimeError(f"sylver_b200_match_order failed ({rc})")
    return rc, order[:n], scale[:n], pairs[:n]


def plan_levels(solver: "Solver", rank: int = 0, world: int = 1):
    """Host-only level plan (sylver_b200_plan_levels): list of dicts per level with
    fronts, contrib_tiles and steps (list of dicts cnt, trsm, upd, updn, updr, upd2n, upd2r, wld)."""
    cnt = lib().sylver_b200_plan_levels(solver.akeep, rank, world, 0, None)
    if cnt < 0:
        raise RuntimeError("sylver_b200_plan_levels failed")
    buf = np.zeros(max(cnt, 1), dtype=np.int64)
    lib().sylver_b200_plan_levels(solver.akeep, rank, world, cnt, _ptr(buf))
    out, k = [], 0
    keys = ("cnt", "trsm", "upd", "updn", "updr", "upd2n", "upd2r", "wld")
    while k < cnt:
        lvl, fronts, nsteps, ctiles = (int(v) for v in buf[k:k + 4])
        k += 4
        steps = []
        for _ in range(nsteps):
            steps.append(dict(zip(keys, (int(v) for v in buf[k:k + 8]))))
            k += 8
        out.append(dict(level=lvl, fronts=fronts, contrib_tiles=ctiles, steps=steps))
    return out


def plan_split(solver: "Solver", rank: int, world: int):
    """Host-only positive definite plan of `rank` among `world` ranks (sylver_b200_plan_split):
    (summary dict, pieces array of rows (level, front, peer, offset, count, direction, group)) --
    group = id of the NCCL group the operation is issued in (identical on both sides of a pair)."""
    out = np.zeros(8, dtype=np.int64)
    cnt = lib().sylver_b200_plan_split(solver.akeep, rank, world, _ptr(out), 0, None)
    if cnt < 0:
        raise RuntimeError("sylver_b200_plan_split failed")
    pieces = np.zeros((max(cnt, 1), 6), dtype=np.int64)
    lib().sylver_b200_plan_split(solver.akeep, rank, world, _ptr(out), 6 * cnt, _ptr(pieces))
    keys = ("split_fronts", "split_member", "factor_bytes", "contrib_bytes", "stage_bytes", "sends", "recvs", "max_ops_level")
    pieces = pieces[:cnt]
    pieces = np.column_stack([pieces[:, :5], pieces[:, 5] & 1, pieces[:, 5] >> 1]) if cnt else np.zeros((0, 7), dtype=np.int64)
    return dict(zip(keys, (int(v) for v in out))), pieces


def plan_exchanges(solver: "Solver", rank: int, world: int) -> np.ndarray:
    """(level, front, peer, dir) rows; dir 0 = send, 1 = recv."""
    cnt = lib().sylver_b200_plan_exchanges(solver.akeep, rank, world, 0, None)
    out = np.zeros((max(cnt, 1), 4), dtype=np.int32)
    lib().sylver_b200_plan_exchanges(solver.akeep, rank, world, 4 * cnt, _ptr(out))
    return out[:cnt]


def default_options_c() -> OptionsC:
    """sylver::options_c with SyLVER's defaults (src/sylver_datatypes_mod.F90:97-198)."""
    o = OptionsC()
    o.print_level = 0
    o.action = True
    o.small = 1e-20
    o.u = 0.01
    o.multiplier = 1.1
    o.small_subtree_threshold = 4 * 10 ** 6
    o.nb = 256
    o.pivot_method = 2
    o.failed_pivot_method = 1
    o.cpu_topology = 1
    return o


def factor_front_indef(a: np.ndarray, n: int, options: OptionsC | None = None):
    """APTP LDL^T of the first n columns of the dense symmetric front ``a`` (m x m) on the
    GPU (reference harness shape: tests/testing_factor_node_indef.hxx:44-460).
    Returns dict(nelim, L (m x n), d (2n), perm (n, 1-based), contrib ((m-n)^2), stats, ms)."""
    require_gpu()
    opt = options or default_options_c()
    m = a.shape[0]
    lda = m
    buf = np.zeros((lda, n), order="F")
    buf[:m, :] = np.tril(a)[:, :n]
    d = np.zeros(2 * n + 2)
    perm = np.arange(1, n + 1, dtype=np.int32)
    k = m - n
    contrib = np.zeros((max(k, 1), max(k, 1)), order="F")
    stats = InformC()
    ms = C.c_float(0)
    nelim = lib().sylver_b200_factor_front_indef(m, n, _ptr(perm), _ptr(buf), lda, _ptr(d), _ptr(contrib),
                                                 C.byref(opt), C.byref(stats), C.byref(ms))
    return dict(nelim=nelim, L=buf, d=d[:2 * n].copy(), perm=perm, contrib=contrib[:k, :k].copy(),
                stats=stats, ms=ms.value)


class Solver:
    """Mirror of the reference C example's call sequence
    (/root/reference/examples/C/spldlt_simple_example_c.c:27-75)."""

    def __init__(self, ngpu: int = 1):
        self.L = lib()
        self.L.sylver_init(1, ngpu)
        self.akeep = C.c_void_p(None)
        self.fkeep = C.c_void_p(None)
        self.options = default_options()
        self.inform = Inform()
        self.n = 0
        self.ptr = self.row = None

    def analyse(self, n, ptr, row, order, val=None, check=False):
        self.n = n
        self.ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        self.row = np.ascontiguousarray(row, dtype=np.int32)
        self.order = np.ascontiguousarray(order, dtype=np.int32).copy()
        self.L.spldlt_analyse(n, _ptr(self.order), _ptr(self.ptr), _ptr(self.row), _ptr(val),
                              C.byref(self.akeep), check, C.byref(self.options), C.byref(self.inform))
        return self.inform

    def symbolic(self):
        """Return the seam arrays as numpy copies (1-based values, as in Fortran)."""
        v = SymbolicView()
        if self.L.sylver_b200_akeep_view(self.akeep, C.byref(v)) != 0:
            raise RuntimeError("analyse has not been run")
        nn = v.nnodes
        out = dict(n=v.n, nnodes=nn, num_factor=v.num_factor, num_flops=v.num_flops)
        out["sptr"] = np.ctypeslib.as_array(v.sptr, (nn + 1,)).copy()
        out["sparent"] = np.ctypeslib.as_array(v.sparent, (max(nn, 1),)).copy()[:nn]
        out["rptr"] = np.ctypeslib.as_array(v.rptr, (nn + 1,)).copy()
        nr = int(out["rptr"][nn] - 1) if nn else 0
        out["rlist"] = np.ctypeslib.as_array(v.rlist, (max(nr, 1),)).copy()[:nr]
        out["nptr"] = np.ctypeslib.as_array(v.nptr, (nn + 1,)).copy()
        ne = int(out["nptr"][nn] - 1) if nn else 0
        out["nlist"] = np.ctypeslib.as_array(v.nlist, (max(2 * ne, 1),)).copy()[:2 * ne]
        out["order"] = np.ctypeslib.as_array(v.order, (max(v.n, 1),)).copy()[:v.n]
        out["invp"] = np.ctypeslib.as_array(v.invp, (max(v.n, 1),)).copy()[:v.n]
        return out

    def cmap(self):
        tree = self.L.sylver_b200_akeep_tree(self.akeep)
        cptr = C.POINTER(C.c_long)()
        cm = C.POINTER(C.c_int)()
        self.L.sylver_b200_symbolic_tree_cmap(tree, C.byref(cptr), C.byref(cm))
        nn = self.symbolic()["nnodes"]
        p = np.ctypeslib.as_array(cptr, (nn + 1,)).copy()
        if int(p[nn]) == 0:
            return p, np.zeros(0, dtype=np.int32)
        return p, np.ctypeslib.as_array(cm, (int(p[nn]),)).copy()

    def factorize(self, val, posdef: bool, scale=None):
        """val: numpy array (host) or an int device address."""
        if not isinstance(val, int):
            val = np.ascontiguousarray(val, dtype=np.float64)
        self._val = val
        self.L.spldlt_factorize(posdef, _ptr(self.ptr), _ptr(self.row), _ptr(val), _ptr(scale),
                                self.akeep, C.byref(self.fkeep), C.byref(self.options),
                                C.byref(self.inform))
        return self.inform

    def timings(self):
        out = (C.c_double * 4)()
        tree = self.L.sylver_b200_fkeep_tree(self.fkeep)
        if not tree or self.L.sylver_b200_numeric_tree_timings(tree, out) != 0:
            return None
        return dict(device_s=out[0], h2d_s=out[1], wall_s=out[2], launches=int(out[3]))

    def engine_tree(self):
        """The chain-coarsened tree the engine factorizes: dict(nnodes, nrow, ncol, parent
        (0-based, nnodes = virtual root), node_map (reference front -> engine front))."""
        tree = self.L.sylver_b200_akeep_tree(self.akeep)
        nn = C.c_int(0)
        pr, pc, pp, pm = (C.POINTER(C.c_int)() for _ in range(4))
        nref = self.L.sylver_b200_symbolic_tree_view(tree, C.byref(nn), C.byref(pr), C.byref(pc), C.byref(pp), C.byref(pm))
        if nref < 0:
            return None
        g = nn.value
        return dict(nnodes=g, nrow=np.ctypeslib.as_array(pr, (g,)).copy(), ncol=np.ctypeslib.as_array(pc, (g,)).copy(),
                    parent=np.ctypeslib.as_array(pp, (g,)).copy(),
                    node_map=np.ctypeslib.as_array(pm, (nref,)).copy() if nref else np.zeros(0, np.int32))

    def split_info(self):
        """(split fronts in the tree, split fronts this rank works on, pieces sent) of the last factorization."""
        out = (C.c_int * 3)()
        tree = self.L.sylver_b200_fkeep_tree(self.fkeep)
        if not tree or self.L.sylver_b200_numeric_tree_split_info(tree, out) != 0:
            return None
        return int(out[0]), int(out[1]), int(out[2])

    def solve(self, b, job: int = 0):
        x = np.array(b, dtype=np.float64, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        self.L.spldlt_solve(job, nrhs, _ptr(x), self.n, self.akeep, self.fkeep,
                            C.byref(self.options), C.byref(self.inform))
        return x

    def free(self):
        self.L.spldlt_free_fkeep(C.byref(self.fkeep))
        self.L.spldlt_free_akeep(C.byref(self.akeep))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
