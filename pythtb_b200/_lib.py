"""ctypes binding of libtbk_b200.so (``include/tbk.h``).

Loading fails loudly: there is no CPU fallback.  The library is built in-tree
by ``python -m pythtb_b200.build`` (``__graft_entry__.build()``).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PYTHTB_B200_LIB: an alternative build of the same library (A/B experiments under profiles/)
SO_PATH = os.environ.get("PYTHTB_B200_LIB") or os.path.join(HERE, "libtbk_b200.so")

c_int32, c_int64, c_size_t, c_void_p = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)

# every symbol include/tbk.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "tbk_version", "tbk_last_error", "tbk_model_create", "tbk_model_destroy", "tbk_gen_ham",
    "tbk_eigh_workspace", "tbk_eigh_batched", "tbk_solve_workspace", "tbk_solve_k", "tbk_solve_grid",
    "tbk_impose_boundary", "tbk_flux_workspace", "tbk_flux_plane", "tbk_berry_workspace",
    "tbk_berry_strings", "tbk_position_matrix", "tbk_position_hwf_workspace", "tbk_position_hwf",
    "tbk_flush_l2", "tbk_halo_pack", "tbk_last_kernel", "tbk_launch_count", "tbk_peer_create", "tbk_peer_connect",
    "tbk_peer_destroy", "tbk_solve_grid_x", "tbk_flux_plane_x", "tbk_stream_sync", "tbk_debug_profile", "tbk_debug_cta_trace", "tbk_kmesh_uniform", "tbk_solve_grid_prepare", "tbk_flux_plane_prepare", "tbk_prepared_run", "tbk_prepared_destroy",
    "tbk_peer_barrier", "tbk_peer_defer", "tbk_peer_flush", "tbk_wilson_products", "tbk_wilson_workspace", "tbk_wilson_phases", "tbk_wilson_chain", "tbk_bench_fp64", "tbk_bench_fp64_flops",
]


class ModelDesc(ctypes.Structure):
    """``tbk_model_desc``"""
    _fields_ = [("dim_k", c_int32), ("nsta", c_int32), ("nph", c_int32), ("nel", c_int32),
                ("nterm", c_int32), ("convention", c_int32),
                ("ph_R", c_void_p), ("tau", c_void_p), ("el_ptr", c_void_p), ("el_row", c_void_p),
                ("el_col", c_void_p), ("t_ph", c_void_p), ("t_amp", c_void_p), ("pm_ptr", c_void_p),
                ("pm_el", c_void_p), ("pm_amp", c_void_p)]


class WfView(ctypes.Structure):
    """``tbk_wf_view``"""
    _fields_ = [("wfs_dev", c_void_p), ("n", c_int32), ("nsta_arr", c_int32), ("nocc", c_int32),
                ("occ_dev", c_void_p), ("state_stride", c_int64)]


class TbkError(Exception):
    pass


ERR_UNSUPPORTED = -5


_lib = None


def load():
    """Return the loaded library with argtypes set; raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "pythtb_b200: %s is missing. Build it with `python -m pythtb_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % SO_PATH)
    lib = ctypes.CDLL(SO_PATH)
    V, I32, I64, SZ = c_void_p, c_int32, c_int64, c_size_t
    sig = {
        "tbk_version": (ctypes.c_int, []),
        "tbk_last_error": (ctypes.c_char_p, []),
        "tbk_model_create": (ctypes.c_int, [ctypes.POINTER(ModelDesc), ctypes.POINTER(V)]),
        "tbk_model_destroy": (ctypes.c_int, [V]),
        "tbk_gen_ham": (ctypes.c_int, [V, V, I64, V, V]),
        "tbk_eigh_workspace": (SZ, [I32, I64, I32]),
        "tbk_eigh_batched": (ctypes.c_int, [V, I32, I64, V, V, V, SZ, V]),
        "tbk_solve_workspace": (SZ, [I32, I64, I32]),
        "tbk_solve_k": (ctypes.c_int, [V, V, I64, V, I64, I64, V, I64, I64, V, SZ, V]),
        "tbk_solve_grid": (ctypes.c_int, [V, c_double_p, c_int32_p, I32, I32, I32, I32, V, V, V, V, SZ, V]),
        "tbk_impose_boundary": (ctypes.c_int, [V, I64, I64, I64, I32, I32, V, V]),
        "tbk_flux_workspace": (SZ, [I32, I32, I64, I64, I64]),
        "tbk_flux_plane": (ctypes.c_int, [ctypes.POINTER(WfView), V, I64, I64, I64, I64, I64, V, V, V, SZ, V]),
        "tbk_berry_workspace": (SZ, [I32, I32, I64, I64, I32]),
        "tbk_berry_strings": (ctypes.c_int, [ctypes.POINTER(WfView), V, I64, I64, I64, I32, V, V, SZ, V]),
        "tbk_position_matrix": (ctypes.c_int, [V, I64, I32, I32, V, V, V]),
        "tbk_position_hwf_workspace": (SZ, [I32, I32, I64]),
        "tbk_position_hwf": (ctypes.c_int, [V, I64, I32, I32, V, V, V, I32, V, SZ, V]),
        "tbk_flush_l2": (ctypes.c_int, [V, SZ, V]),
        "tbk_stream_sync": (ctypes.c_int, [V]),
        "tbk_wilson_products": (ctypes.c_int, [ctypes.POINTER(WfView), V, I64, I64, I64, V, V, SZ, V]),
        "tbk_wilson_workspace": (SZ, [I32, I64, I64]),
        "tbk_wilson_phases": (ctypes.c_int, [V, I64, I64, I32, V, V, SZ, V]),
        "tbk_wilson_chain": (ctypes.c_int, [V, I64, I64, I32, V, V, SZ, V]),
        "tbk_bench_fp64": (ctypes.c_int, [I32, I32, V, V]),
        "tbk_bench_fp64_flops": (ctypes.c_double, [I32, I32]),
        "tbk_debug_profile": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), I32]),
        "tbk_debug_cta_trace": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), I64, I32]),
        "tbk_peer_barrier": (ctypes.c_int, [c_void_p, c_void_p]),
        "tbk_solve_grid_prepare": (ctypes.c_int, [V, c_double_p, c_int32_p, I32, I32, I32, I32, V, V, V, V, SZ, I64, V, ctypes.POINTER(V)]),
        "tbk_flux_plane_prepare": (ctypes.c_int, [ctypes.POINTER(WfView), V, I64, I64, I64, I64, I64, V, V, V, SZ, V, ctypes.POINTER(V)]),
        "tbk_prepared_run": (ctypes.c_int, [V, V, I32]),
        "tbk_kmesh_uniform": (ctypes.c_int, [c_int32_p, I32, V, V]),
        "tbk_prepared_destroy": (ctypes.c_int, [V]),
        "tbk_peer_defer": (ctypes.c_int, [c_void_p, c_int32]),
        "tbk_peer_flush": (ctypes.c_int, [c_void_p, c_void_p]),
        "tbk_halo_pack": (ctypes.c_int, [V, V, I64, I32, I32, V, V]),
        "tbk_peer_create": (ctypes.c_int, [I32, I32, ctypes.POINTER(V), V]),
        "tbk_peer_connect": (ctypes.c_int, [V, V]),
        "tbk_peer_destroy": (ctypes.c_int, [V]),
        "tbk_solve_grid_x": (ctypes.c_int, [V, c_double_p, c_int32_p, I32, I32, I32, I32, V, V, V, V, SZ, I64, V, V]),
        "tbk_flux_plane_x": (ctypes.c_int, [ctypes.POINTER(WfView), V, I64, I64, I64, I64, I64, V, V, V, SZ, V, V]),
        "tbk_last_kernel": (ctypes.c_char_p, []),
        "tbk_launch_count": (c_int64, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Turn a non-zero status into an Exception carrying tbk_last_error()."""
    if rc != 0:
        msg = load().tbk_last_error()
        raise TbkError("\n\nlibtbk_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def last_kernel(lib):
    """Name of the kernel family the last solve call on this thread dispatched to."""
    name = lib.tbk_last_kernel()
    return name.decode() if name else "?"
