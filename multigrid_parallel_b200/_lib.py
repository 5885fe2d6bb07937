"""Loader + ctypes prototypes for libmgb.so (include/mgb.h)."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class MgbError(RuntimeError):
    """A libmgb call returned non-zero; the message is mgb_last_error()."""


def lib_path():
    # MGB_LIB: another build of the same library (A/B timing of a kernel variant)
    return os.environ.get("MGB_LIB") or os.path.join(_HERE, "libmgb.so")


def header_path():
    return os.path.join(_ROOT, "include", "mgb.h")


def declared_symbols():
    """Every function name include/mgb.h declares (used by the CPU tests)."""
    text = open(header_path()).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", text)))


_LIB = None


def load_library():
    """dlopen libmgb.so; raises if it was not built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MgbError(
            f"{path} is missing: build it with `make -C multigrid_parallel_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.mgb_last_error.restype = C.c_char_p
    L.mgb_version.restype = C.c_char_p
    L.mgb_device_count.argtypes = [c_ip]
    L.mgb_create.argtypes = [C.POINTER(vp), i, i, i, i, i, i]
    L.mgb_destroy.argtypes = [vp]
    L.mgb_levels.argtypes = [vp]
    L.mgb_dims.argtypes = [vp, i, c_ip, c_ip, c_ip]
    L.mgb_spacing.restype = d
    L.mgb_spacing.argtypes = [vp, i]
    L.mgb_set_option.argtypes = [vp, i, i]
    L.mgb_sync.argtypes = [vp]
    L.mgb_set_global.argtypes = [i, C.c_longlong]
    L.mgb_tile_plan.argtypes = [i, i, i, i, C.POINTER(C.c_longlong)]
    L.mgb_upload.argtypes = [vp, i, i, C.c_void_p]
    L.mgb_download.argtypes = [vp, i, i, C.c_void_p]
    L.mgb_upload_range.argtypes = [vp, i, i, C.c_longlong, C.c_longlong, C.c_void_p]
    L.mgb_download_range.argtypes = [vp, i, i, C.c_longlong, C.c_longlong, C.c_void_p]
    L.mgb_zero.argtypes = [vp, i, i]
    L.mgb_set_dirichlet.argtypes = [vp, i, i]
    L.mgb_sumsq.argtypes = [vp, i, i, c_dp]
    L.mgb_edge_values.argtypes = [vp, i, i]
    L.mgb_set_spacing.argtypes = [vp, d]
    L.mgb_pin_host.argtypes = [C.c_void_p, C.c_ulonglong]
    L.mgb_unpin_host.argtypes = [C.c_void_p]
    L.mgb_error_sumsq.argtypes = [vp, c_dp]
    L.mgb_half_sweep.argtypes = [vp, i, i]
    L.mgb_smooth.argtypes = [vp, i, i, i]
    L.mgb_gs_lex.argtypes = [vp, i, i]
    L.mgb_debug_half_sweep_range.argtypes = [vp, i, i, i, i]
    L.mgb_residual.argtypes = [vp, i, i, c_dp]
    L.mgb_restrict.argtypes = [vp, i]
    L.mgb_residual_restrict.argtypes = [vp, i]
    L.mgb_prolong_correct.argtypes = [vp, i]
    L.mgb_sweep_residual_restrict.argtypes = [vp, i, i]
    L.mgb_sweep_residual.argtypes = [vp, i, i, c_dp]
    L.mgb_coarse_solve.argtypes = [vp]
    L.mgb_coarse_lu_download.argtypes = [vp, C.c_void_p]
    L.mgb_coarse_info.argtypes = [vp, c_ip, c_ip, c_dp]
    L.mgb_vcycle.argtypes = [vp, c_dp]
    L.mgb_vcycles.argtypes = [vp, i, c_dp]
    L.mgb_fmg_init.argtypes = [vp, c_dp]
    L.mgb_solve.argtypes = [vp, d, i, c_dp, c_ip]
    L.mgb_timing.argtypes = [vp, i, i, c_ip, c_dp]
    L.mgb_timing_reset.argtypes = [vp]
    L.mgb_launch_count.restype = C.c_longlong
    L.mgb_launch_count.argtypes = [vp]
    L.mgb_timer_start.argtypes = [vp]
    L.mgb_timer_stop.argtypes = [vp, c_dp]
    L.mgb_stream.restype = C.c_void_p
    L.mgb_stream.argtypes = [vp]
    L.mgb_nccl_unique_id.argtypes = [C.c_void_p]
    L.mgb_create_dist.argtypes = [C.POINTER(vp), i, i, i, i, i, i, i, i, C.c_void_p, i,
                                  C.c_longlong]
    L.mgb_dist_info.argtypes = [vp, c_ip, c_ip, c_ip]
    L.mgb_local_range.argtypes = [vp, i, c_ip, c_ip, c_ip, c_ip]
    L.mgb_plan_slab.argtypes = [i, i, i, c_ip, c_ip]
    L.mgb_plan_first_dist_level.argtypes = [i, i, i, i, i, i, C.c_longlong]
    L.mgb_host_smooth.argtypes = [c_dp, c_dp, i, i, i, d, i, i]
    L.mgb_host_gs_lex.argtypes = [c_dp, c_dp, i, i, i, d, i, i]
    L.mgb_vtk_open.argtypes = [C.POINTER(vp), c_dp, i, i, i, d, i]
    L.mgb_vtk_next.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_longlong)]
    L.mgb_vtk_host_chunks.argtypes = [vp, C.POINTER(C.c_longlong)]
    L.mgb_vtk_close.argtypes = [vp]
    L.mgb_host_residual.argtypes = [c_dp, c_dp, i, i, i, d, c_dp, c_dp]
    L.mgb_host_restrict.argtypes = [c_dp, i, i, i, c_dp, i, i, i]
    L.mgb_host_prolong_correct.argtypes = [c_dp, i, i, i, c_dp, i, i, i]
    L.mgb_host_coarse_matrix.argtypes = [c_dp, i, i, i, d]
    L.mgb_host_lu_factor.argtypes = [c_dp, i]
    L.mgb_host_lu_solve.argtypes = [c_dp, i, c_dp, c_dp]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        msg = load_library().mgb_last_error()
        raise MgbError(msg.decode() if msg else f"libmgb error {rc}")
