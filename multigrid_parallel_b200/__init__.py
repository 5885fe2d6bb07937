"""multigrid_parallel_b200 -- B200-native 3D geometric-multigrid V-cycle.

The product is ``libmgb.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/mgb.h``) plus the C drop-in headers under ``compat/`` that mirror the
reference's ``mg_3d.h`` API.  This Python package is only the thin ctypes
binding the tests and ``bench.py`` use; there is no CPU fallback anywhere.
"""
from ._lib import MgbError, lib_path, load_library  # noqa: F401
from .solver import (  # noqa: F401
    MGB_D, MGB_R, MGB_U, STAGE_NAMES, Solver, host_coarse_matrix, host_lu_factor,
    host_lu_solve, host_prolong_correct, host_residual, host_restrict, host_smooth,
    host_gs_lex, vtk_bytes, vtk_stream,
    set_global, G_TILE, G_TILE_MIN_PLANE, G_GSLEX_TILE,
)

__all__ = [
    "Solver", "MgbError", "load_library", "lib_path", "MGB_U", "MGB_D", "MGB_R",
    "STAGE_NAMES", "set_global", "G_TILE", "G_TILE_MIN_PLANE", "G_GSLEX_TILE", "host_smooth", "host_gs_lex", "vtk_bytes", "vtk_stream", "host_residual", "host_restrict",
    "host_prolong_correct", "host_coarse_matrix", "host_lu_factor", "host_lu_solve",
]
