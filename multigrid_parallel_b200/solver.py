"""Python mirror of the C ABI (include/mgb.h): a handle class + array helpers.

Everything here forwards to libmgb.so; arrays cross the boundary as float64
numpy arrays in the reference's natural layout (shape (ni, nj, nk), C order).
"""
import ctypes as C
import math

import numpy as np

from ._lib import c_dp, check, load_library

MGB_U, MGB_D, MGB_R = 0, 1, 2
OPT_GRAPH, OPT_PROFILE, OPT_FUSE, OPT_GRAPH_LEVELS, OPT_TAIL, OPT_ZERO_GUESS = 0, 1, 2, 3, 4, 5
OPT_PROLONG_MASK = 6
# mg_3d.h:136-137
STAGE_NAMES = ("Smoother1", "CalcResidual1", "Restrict Residual", "Recurse, Direct Solve",
               "Prolongate&Correct", "Smoother2", "CalcResidual2")


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need C-contiguous float64"
    return a.ctypes.data_as(c_dp)


class Solver:
    """One multigrid hierarchy on one GPU (mgb_create ... mgb_destroy).

    Mirrors the reference's solver state (mg_3d.h:19-28): `coarse` points per
    side on level 0, `levels` grids, `gs` smoothing iterations per leg.
    """

    def __init__(self, coarse, levels, gs, device=0, rank=0, nranks=1, nccl_uid=None,
                 min_planes=16, min_points=-1):
        self.L = load_library()
        if isinstance(coarse, int):
            coarse = (coarse,) * 3
        self.h_ = C.c_void_p()
        self.rank, self.nranks = rank, nranks
        if nranks == 1:
            check(self.L.mgb_create(C.byref(self.h_), *coarse, levels, gs, device))
        else:
            # slab-partitioned over i, one process per GPU (mgb_create_dist)
            assert nccl_uid is not None and len(nccl_uid) == 128
            buf = C.create_string_buffer(bytes(nccl_uid), 128)
            check(self.L.mgb_create_dist(C.byref(self.h_), *coarse, levels, gs, device, rank,
                                         nranks, buf, min_planes, min_points))
        self.levels = levels
        self.gs = gs

    @staticmethod
    def nccl_unique_id():
        """rank 0: the 128-byte NCCL id every rank passes to the constructor"""
        L = load_library()
        buf = C.create_string_buffer(128)
        check(L.mgb_nccl_unique_id(buf))
        return buf.raw

    def local_range(self, level=None):
        """(i0, li, own_lo, own_hi): local planes [i0, i0+li), owned [own_lo, own_hi)"""
        level = self.levels - 1 if level is None else level
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.L.mgb_local_range(self.h_, level, a, b, c, d))
        return a.value, b.value, c.value, d.value

    def local_shape(self, level=None):
        level = self.levels - 1 if level is None else level
        ni, nj, nk = self.dims(level)
        return (self.local_range(level)[1], nj, nk)

    @property
    def first_dist_level(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.L.mgb_dist_info(self.h_, a, b, c))
        return c.value

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "h_", None) is not None and self.h_:
            self.L.mgb_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- geometry ----------------------------------------------------------
    def dims(self, level=None):
        level = self.levels - 1 if level is None else level
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.L.mgb_dims(self.h_, level, a, b, c))
        return a.value, b.value, c.value

    def spacing(self, level=None):
        level = self.levels - 1 if level is None else level
        return self.L.mgb_spacing(self.h_, level)

    def set_option(self, key, value):
        check(self.L.mgb_set_option(self.h_, key, int(value)))

    def sync(self):
        check(self.L.mgb_sync(self.h_))

    # -- arrays ------------------------------------------------------------
    def upload(self, level, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.shape == self.local_shape(level), (host.shape, self.local_shape(level))
        check(self.L.mgb_upload(self.h_, level, which, host.ctypes.data))

    def upload_ptr(self, level, which, ptr):
        check(self.L.mgb_upload(self.h_, level, which, ptr))

    def download(self, level, which, out=None):
        if out is None:
            out = np.empty(self.local_shape(level), dtype=np.float64)
        check(self.L.mgb_download(self.h_, level, which, out.ctypes.data))
        return out

    def download_ptr(self, level, which, ptr):
        check(self.L.mgb_download(self.h_, level, which, ptr))

    def zero(self, level, which):
        check(self.L.mgb_zero(self.h_, level, which))

    def set_dirichlet(self, level, which):
        check(self.L.mgb_set_dirichlet(self.h_, level, which))

    def edge_values(self, level, which):
        """updateEdgeValues (mg_3d.h:304-430) on the device array"""
        check(self.L.mgb_edge_values(self.h_, level, which))

    def set_spacing(self, h):
        check(self.L.mgb_set_spacing(self.h_, h))

    def sumsq(self, level, which):
        v = C.c_double()
        check(self.L.mgb_sumsq(self.h_, level, which, v))
        return v.value

    def error_sumsq(self):
        v = C.c_double()
        check(self.L.mgb_error_sumsq(self.h_, v))
        return v.value

    # -- operators ---------------------------------------------------------
    def half_sweep(self, level, colour):
        check(self.L.mgb_half_sweep(self.h_, level, colour))

    def smooth(self, level, iters, first_red):
        check(self.L.mgb_smooth(self.h_, level, iters, int(first_red)))

    def gs_lex(self, level, iters):
        """`iters` lexicographic Gauss-Seidel sweeps (GaussSeidelSmoother, mg_3d.h:546-634)"""
        check(self.L.mgb_gs_lex(self.h_, level, iters))

    def residual(self, level, store_r=False):
        """returns sqrt(sum of squares) like calculateResidual on one thread"""
        v = C.c_double()
        check(self.L.mgb_residual(self.h_, level, int(store_r), v))
        return math.sqrt(v.value)

    def restrict(self, level):
        check(self.L.mgb_restrict(self.h_, level))

    def residual_restrict(self, level):
        check(self.L.mgb_residual_restrict(self.h_, level))

    def sweep_residual_restrict(self, level, colour):
        check(self.L.mgb_sweep_residual_restrict(self.h_, level, colour))

    def sweep_residual(self, level, colour):
        v = C.c_double()
        check(self.L.mgb_sweep_residual(self.h_, level, colour, v))
        return math.sqrt(v.value)

    def prolong_correct(self, level):
        check(self.L.mgb_prolong_correct(self.h_, level))

    def coarse_solve(self):
        check(self.L.mgb_coarse_solve(self.h_))

    def coarse_lu(self):
        n = int(np.prod(self.dims(0)))
        a = np.empty((n, n))
        check(self.L.mgb_coarse_lu_download(self.h_, a.ctypes.data))
        return a

    def coarse_info(self):
        """(n, half bandwidth, seconds the build + factorisation took at create)"""
        n, bw, t = C.c_int(), C.c_int(), C.c_double()
        check(self.L.mgb_coarse_info(self.h_, n, bw, t))
        return n.value, bw.value, t.value

    # -- cycles ------------------------------------------------------------
    def vcycle(self):
        """one V-cycle; returns the residual 2-norm (SolverLinSolve)"""
        v = C.c_double()
        check(self.L.mgb_vcycle(self.h_, v))
        return math.sqrt(v.value)

    def vcycles(self, n):
        """n cycles back to back, one read-back of the last norm (mgb_vcycles)"""
        v = C.c_double()
        check(self.L.mgb_vcycles(self.h_, int(n), v))
        return math.sqrt(v.value)

    def fmg_init(self):
        """SolverFMGInitialize (mg_3d.h:1364-1404); returns the residual 2-norm after it"""
        v = C.c_double()
        check(self.L.mgb_fmg_init(self.h_, v))
        return math.sqrt(v.value)

    def solve(self, threshold, max_cycles=100):
        hist = np.zeros(max_cycles)
        n = C.c_int()
        check(self.L.mgb_solve(self.h_, threshold, max_cycles, _dp(hist), n))
        return hist[: n.value].copy()

    def timing(self, level, stage):
        c, s = C.c_int(), C.c_double()
        check(self.L.mgb_timing(self.h_, level, stage, c, s))
        return c.value, s.value

    def timing_reset(self):
        check(self.L.mgb_timing_reset(self.h_))

    def timer_start(self):
        check(self.L.mgb_timer_start(self.h_))

    def timer_stop(self):
        """device seconds since timer_start (CUDA events on the solver's stream)"""
        v = C.c_double()
        check(self.L.mgb_timer_stop(self.h_, v))
        return v.value

    @property
    def launch_count(self):
        return self.L.mgb_launch_count(self.h_)


G_TILE, G_TILE_MIN_PLANE, G_GSLEX_TILE = 0, 1, 2


def set_global(key, value):
    """process-wide kernel selection (mgb_set_global)"""
    check(load_library().mgb_set_global(key, int(value)))


# ---- stateless array entry points (the reference's raw-pointer API) -------
def host_smooth(v, d, h, iters, first_red):
    L = load_library()
    check(L.mgb_host_smooth(_dp(v), _dp(d), *v.shape, h, iters, int(first_red)))


def host_gs_lex(v, d, h, iters, edges=True):
    """GaussSeidelSmoother(v, d, N, h, iters) on host arrays (mg_3d.h:546-637)"""
    L = load_library()
    check(L.mgb_host_gs_lex(_dp(v), _dp(d), *v.shape, h, iters, int(edges)))


def vtk_stream(values, h, device=0):
    """writeOutputData (postprocess.h:5-47) through mgb_vtk_*: yields the byte chunks of the
    reference's ASCII VTK file for the grid `values` (shape (ni, nj, nk)); the generator's
    `host_chunks` attribute is not available -- use vtk_bytes for the count"""
    L = load_library()
    w = C.c_void_p()
    check(L.mgb_vtk_open(C.byref(w), _dp(values), *values.shape, h, device))
    try:
        p = C.c_void_p()
        n = C.c_longlong()
        nxt = L.mgb_vtk_next
        nxt.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
        while True:
            check(nxt(w, C.byref(p), C.byref(n)))
            if n.value == 0:
                break
            yield C.string_at(p.value, n.value)
        hc = C.c_longlong()
        check(L.mgb_vtk_host_chunks(w, C.byref(hc)))
        vtk_stream.last_host_chunks = hc.value
    finally:
        check(L.mgb_vtk_close(w))


def vtk_bytes(values, h, device=0):
    """the whole file; returns (bytes, number of chunks the host's snprintf had to format)"""
    data = b"".join(vtk_stream(values, h, device))
    return data, vtk_stream.last_host_chunks


def host_residual(v, d, h, res=None):
    L = load_library()
    ss = C.c_double()
    check(L.mgb_host_residual(_dp(v), _dp(d), *v.shape, h,
                              _dp(res) if res is not None else None, ss))
    return math.sqrt(ss.value)


def host_restrict(r, dc):
    L = load_library()
    check(L.mgb_host_restrict(_dp(r), *r.shape, _dp(dc), *dc.shape))


def host_prolong_correct(ec, ef):
    L = load_library()
    check(L.mgb_host_prolong_correct(_dp(ec), *ec.shape, _dp(ef), *ef.shape))


def host_coarse_matrix(shape, h):
    L = load_library()
    n = int(np.prod(shape))
    A = np.zeros((n, n))
    check(L.mgb_host_coarse_matrix(_dp(A), *shape, h))
    return A


def host_lu_factor(a):
    L = load_library()
    check(L.mgb_host_lu_factor(_dp(a), a.shape[0]))


def host_lu_solve(lu, b):
    L = load_library()
    x = np.zeros_like(b)
    check(L.mgb_host_lu_solve(_dp(lu), lu.shape[0], _dp(b), _dp(x)))
    return x
