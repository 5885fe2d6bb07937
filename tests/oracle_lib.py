"""ctypes bindings for the CPU checker under oracle/ (test infrastructure only).

`orc`  = oracle/liborc.so        -- the repo's own plain-C restatement
`ref`  = oracle/_ref/libmg_ref.so -- the UNMODIFIED reference header compiled
         from /root/reference (present wherever `make -C oracle` ran with the
         reference mounted; the built .so travels to the GPU box)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


class Orc:
    """The repo's restatement (oracle/mg_oracle.c)."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liborc.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        d, i, p = C.c_double, C.c_int, c_dp
        L.orc_bcfunc.restype = d
        L.orc_bcfunc.argtypes = [d, d, d]
        L.orc_set_dirichlet.argtypes = [p, i, i, i, d]
        L.orc_smooth.argtypes = [p, p, i, i, i, d, i, i]
        L.orc_half_sweep.argtypes = [p, p, i, i, i, d, i]
        L.orc_gs_lex.argtypes = [p, p, i, i, i, d, i, i]
        L.orc_edge_values.argtypes = [p, i, i, i]
        L.orc_residual.restype = d
        L.orc_residual.argtypes = [p, p, i, i, i, d, p]
        L.orc_restrict.argtypes = [p, i, i, i, p, i, i, i]
        L.orc_prolong_correct.argtypes = [p, i, i, i, p, i, i, i]
        L.orc_coarse_matrix.argtypes = [p, i, i, i, d]
        L.orc_lu_factor.argtypes = [p, i]
        L.orc_lu_solve.argtypes = [p, i, p, p]
        L.orc_write_vtk.restype = i
        L.orc_write_vtk.argtypes = [C.c_char_p, p, i, i, i, d]
        L.orc_l2norm.restype = d
        L.orc_l2norm.argtypes = [p, C.c_long]
        L.orc_mg_create.restype = C.c_void_p
        L.orc_mg_create.argtypes = [i, i, i, i, i]
        L.orc_mg_destroy.argtypes = [C.c_void_p]
        L.orc_mg_dims.argtypes = [C.c_void_p, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
        for f in ("orc_mg_u", "orc_mg_d", "orc_mg_r"):
            getattr(L, f).restype = p
            getattr(L, f).argtypes = [C.c_void_p, i]
        L.orc_mg_h.restype = d
        L.orc_mg_h.argtypes = [C.c_void_p]
        L.orc_mg_vcycle.restype = d
        L.orc_mg_vcycle.argtypes = [C.c_void_p]
        L.orc_mg_solve.restype = i
        L.orc_mg_solve.argtypes = [C.c_void_p, d, i, p, p]
        L.orc_mg_fmg_init.argtypes = [C.c_void_p]
        L.orc_mg_setup_problem.restype = d
        L.orc_mg_setup_problem.argtypes = [C.c_void_p]

    # array-level helpers (arrays are float64, shape (ni,nj,nk), C order)
    def set_dirichlet(self, v, h):
        self.L.orc_set_dirichlet(_p(v), *v.shape, h)

    def smooth(self, v, d, h, iters, first_red):
        self.L.orc_smooth(_p(v), _p(d), *v.shape, h, iters, int(first_red))

    def half_sweep(self, v, d, h, colour):
        self.L.orc_half_sweep(_p(v), _p(d), *v.shape, h, colour)

    def gs_lex(self, v, d, h, iters, edges=True):
        """GaussSeidelSmoother (mg_3d.h:546-637)"""
        self.L.orc_gs_lex(_p(v), _p(d), *v.shape, h, iters, int(edges))

    def edge_values(self, v):
        self.L.orc_edge_values(_p(v), *v.shape)

    def residual(self, v, d, h, res=None):
        return self.L.orc_residual(_p(v), _p(d), *v.shape, h, _p(res))

    def restrict(self, r, dc):
        self.L.orc_restrict(_p(r), *r.shape, _p(dc), *dc.shape)

    def prolong_correct(self, ec, ef):
        self.L.orc_prolong_correct(_p(ec), *ec.shape, _p(ef), *ef.shape)

    def coarse_matrix(self, shape, h):
        n = int(np.prod(shape))
        A = np.zeros((n, n))
        self.L.orc_coarse_matrix(_p(A), *shape, h)
        return A

    def lu_factor(self, a):
        self.L.orc_lu_factor(_p(a), a.shape[0])

    def lu_solve(self, lu, b):
        x = np.zeros_like(b)
        self.L.orc_lu_solve(_p(lu), lu.shape[0], _p(b), _p(x))
        return x

    def l2norm(self, d):
        return self.L.orc_l2norm(_p(d), d.size)

    def write_vtk(self, path, grid, h):
        """writeOutputData (postprocess.h:5-47) for a box"""
        assert self.L.orc_write_vtk(str(path).encode(), _p(grid), *grid.shape, h) == 0


class OrcMG:
    """Multilevel driver of the restatement."""

    def __init__(self, orc, coarse, levels, gs):
        if isinstance(coarse, int):
            coarse = (coarse,) * 3
        self.L = orc.L
        self.h_ = self.L.orc_mg_create(*coarse, levels, gs)
        self.levels = levels

    def close(self):
        if self.h_:
            self.L.orc_mg_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        self.close()

    def dims(self, level):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_mg_dims(self.h_, level, a, b, c)
        return a.value, b.value, c.value

    def _arr(self, fn, level):
        shape = self.dims(level)
        ptr = fn(self.h_, level)
        return np.ctypeslib.as_array(ptr, shape=shape)

    def u(self, level):
        return self._arr(self.L.orc_mg_u, level)

    def d(self, level):
        return self._arr(self.L.orc_mg_d, level)

    def r(self, level):
        return self._arr(self.L.orc_mg_r, level)

    @property
    def h(self):
        return self.L.orc_mg_h(self.h_)

    def vcycle(self):
        return self.L.orc_mg_vcycle(self.h_)

    def setup_problem(self):
        """test_mg_3d.c:17-29; returns ||d||"""
        return self.L.orc_mg_setup_problem(self.h_)

    def fmg_init(self):
        self.L.orc_mg_fmg_init(self.h_)

    def solve(self, tol=1e-8, max_cycles=100):
        hist = np.zeros(max_cycles)
        init = np.zeros(1)
        n = self.L.orc_mg_solve(self.h_, tol, max_cycles, _p(hist), _p(init))
        return hist[:n].copy(), float(init[0])


class Ref:
    """The reference itself (cubes only), if oracle/_ref was built."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libmg_ref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/mg_3d.h"):
            build()
        self.L = L = _load(path)
        if L is None:
            return
        d, i, p = C.c_double, C.c_int, c_dp
        L.ref_set_threads.argtypes = [i]
        L.ref_max_threads.restype = i
        L.ref_set_dirichlet.argtypes = [p, i, d]
        L.ref_smooth.argtypes = [p, p, i, d, i, i]
        L.GaussSeidelSmoother.argtypes = [p, p, i, d, i]  # the reference's own symbol
        L.updateEdgeValues.argtypes = [p, i]
        L.ref_residual.restype = d
        L.ref_residual.argtypes = [p, p, i, d, p]
        L.ref_restrict.argtypes = [p, i, p, i]
        L.ref_prolong_correct.argtypes = [p, i, p, i]
        L.ref_coarse_matrix.argtypes = [p, i, d]
        L.ref_lu_factor.argtypes = [p, i]
        L.ref_lu_solve.argtypes = [p, i, p, p]
        if hasattr(L, "ref_write_vtk"):
            L.ref_write_vtk.argtypes = [C.c_char_p, p, d, i]
        L.ref_l2norm.restype = d
        L.ref_l2norm.argtypes = [p, i]
        L.ref_solve.restype = i
        L.ref_solve.argtypes = [i, i, i, d, i, p, p, p, p]
        L.ref_solve_fmg.restype = i
        L.ref_solve_fmg.argtypes = [i, i, i, d, i, p, p, p, p]
        L.ref_setup_problem.restype = d
        L.ref_solver_open.restype = i
        L.ref_solver_open.argtypes = [i, i, i, C.POINTER(p), C.POINTER(p), p]
        L.ref_vcycle.restype = d
        for f in ("ref_level_u", "ref_level_d", "ref_level_r"):
            getattr(L, f).restype = p
            getattr(L, f).argtypes = [i]

    @property
    def available(self):
        return self.L is not None

    def set_threads(self, n):
        self.L.ref_set_threads(n)

    def set_dirichlet(self, v, h):
        self.L.ref_set_dirichlet(_p(v), v.shape[0], h)

    def smooth(self, v, d, h, iters, first_red):
        self.L.ref_smooth(_p(v), _p(d), v.shape[0], h, iters, int(first_red))

    def gs_lex(self, v, d, h, iters):
        """the reference's GaussSeidelSmoother (sweeps + updateEdgeValues), cubes only"""
        self.L.GaussSeidelSmoother(_p(v), _p(d), v.shape[0], h, iters)

    def edge_values(self, v):
        self.L.updateEdgeValues(_p(v), v.shape[0])

    def write_vtk(self, path, grid, h):
        """the reference's own writeOutputData (postprocess.h:5-47), cubes only"""
        self.L.ref_write_vtk(str(path).encode(), _p(grid), h, grid.shape[0])

    def residual(self, v, d, h, res=None):
        return self.L.ref_residual(_p(v), _p(d), v.shape[0], h, _p(res))

    def restrict(self, r, dc):
        self.L.ref_restrict(_p(r), r.shape[0], _p(dc), dc.shape[0])

    def prolong_correct(self, ec, ef):
        self.L.ref_prolong_correct(_p(ec), ec.shape[0], _p(ef), ef.shape[0])

    def coarse_matrix(self, N, h):
        n = N ** 3
        A = np.zeros((n, n))
        self.L.ref_coarse_matrix(_p(A), N, h)
        return A

    def lu_factor(self, a):
        self.L.ref_lu_factor(_p(a), a.shape[0])

    def lu_solve(self, lu, b):
        x = np.zeros_like(b)
        self.L.ref_lu_solve(_p(lu), lu.shape[0], _p(b), _p(x))
        return x

    def solve(self, coarse, levels, gs, tol=1e-8, max_cycles=100, want_u=True, fmg=False):
        N = (coarse - 1) * (1 << (levels - 1)) + 1
        hist = np.zeros(max_cycles)
        init = np.zeros(1)
        secs = np.zeros(1)
        u = np.zeros((N, N, N)) if want_u else None
        fn = self.L.ref_solve_fmg if fmg else self.L.ref_solve
        n = fn(coarse, levels, gs, tol, max_cycles, _p(hist), _p(init), _p(u), _p(secs))
        return hist[:n].copy(), float(init[0]), u, float(secs[0])

    def fmg_state(self, coarse, levels, gs):
        """every level's u and d right after the set-up + FMG initialisation"""
        grid, rhs, h = c_dp(), c_dp(), C.c_double()
        self.L.ref_solver_open(coarse, levels, gs, C.byref(grid), C.byref(rhs), C.byref(h))
        self.L.ref_setup_problem()
        self.L.ref_fmg_init()
        out = []
        for l in range(levels):
            n = (coarse - 1) * (1 << l) + 1
            out.append((np.ctypeslib.as_array(self.L.ref_level_u(l), shape=(n, n, n)).copy(),
                        np.ctypeslib.as_array(self.L.ref_level_d(l), shape=(n, n, n)).copy()))
        self.L.ref_solver_close()
        return out


def seeded(shape, seed):
    """uniform(-1,1) float64 array from a fixed seed (PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 1.0, size=shape)
