"""CPU: the C-ABI library loads, exports every symbol include/mgb.h declares,
fails loudly without a GPU (no CPU fallback), and the host-side headers are
self-consistent.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(mgb):
    from multigrid_parallel_b200._lib import declared_symbols
    lib = mgb.load_library()
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert b"sm_100a" in lib.mgb_version()


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "mgb.h")).read()
    for needle in ("mg_3d.h:107-144", "mg_3d.h:794-842", "mg_3d.h:844-998", "mg_3d.h:1000-1145",
                   "1242-1362", "gauss_elim.h:31-60", "mg_3d.h:640-709"):
        assert needle in text, needle


def test_no_cpu_fallback_without_gpu(mgb):
    if _have_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(mgb.MgbError):
        mgb.Solver(3, 3, 2)
    import numpy as np
    v = np.zeros((5, 5, 5))
    with pytest.raises(mgb.MgbError):
        mgb.host_smooth(v, v.copy(), 0.25, 1, True)
    with pytest.raises(mgb.MgbError):
        mgb.host_lu_factor(np.eye(3))


def test_argument_validation_messages(mgb):
    lib = mgb.load_library()
    h = C.c_void_p()
    assert lib.mgb_create(C.byref(h), 3, 3, 3, 0, 2, 0) != 0
    assert b"levels" in lib.mgb_last_error()
    assert lib.mgb_create(C.byref(h), 4, 4, 4, 3, 2, 0) != 0  # 4-1 not a power of two
    assert b"power" in lib.mgb_last_error()
    assert lib.mgb_create(C.byref(h), 2, 3, 3, 3, 2, 0) != 0
    assert lib.mgb_vcycle(None, None) != 0
    assert b"null solver" in lib.mgb_last_error()


def test_product_never_touches_the_oracle():
    """the shipped package must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "multigrid_parallel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, \
                    os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libmgb.so")], capture_output=True, text=True)
    assert "liborc" not in out.stdout and "libmg_ref" not in out.stdout


def test_compat_headers_declare_reference_api():
    """every symbol the reference's drivers use is defined by the
    drop-in headers with the reference's signature"""
    d = os.path.join(ROOT, "multigrid_parallel_b200", "compat")
    text = open(os.path.join(d, "mg_3d.h")).read()
    for sig in [
        r"void SolverInitialize\(int argc, char \*\*argv\)",
        r"int SolverGetDetails\(double \*\*grid, double \*\*rhs, double \*h\)",
        r"void SolverSetupBoundaryConditions\(\)", r"double SolverLinSolve\(\)",
        r"void SolverSmoothenEdgeValues\(\)", r"double SolverGetResidual\(\)",
        r"double SolverGetInitialResidual\(\)", r"void SolverResetTimingInfo\(\)",
        r"void SolverPrintTimingInfo\(\)", r"void SolverFinalize\(\)",
        r"double BCFunc\(double x, double y, double z\)",
        r"void setupBoundaryConditions\(double \*v, int levelN, double spacing\)",
        r"void preSmoother\(", r"void postSmoother\(", r"double calculateResidual\(",
        r"void restrictResidual\(", r"void prolongateAndCorrectError\(",
        r"void GaussSeidelSmoother\(", r"void updateEdgeValues\(",
        r"void constructCoarseMatrixA\(double \*A, int N, const double h\)",
        r"void allocGridLevels\(", r"void deAllocGridLevels\(", r"double vcycle\(",
    ]:
        assert re.search(sig, text), sig
    for g in ("TimingInfo **tInfo", "int coarseGridNum", "int finestOneSideNum", "int numLevels",
              "int gsIterNum", "double **u, **d, **r", "double *A", "double spacing"):
        assert g in text, g
    ge = open(os.path.join(d, "gauss_elim.h")).read()
    assert "void convertToLU_InPlace(double *a, int n)" in ge and "void solveWithLU(" in ge
    ti = open(os.path.join(d, "timing_info.h")).read()
    for f in ("allocTimingInfo", "resetTimingInfo", "printTimingInfo", "deAllocTimingInfo"):
        assert f in ti
    assert "void writeOutputData(const char *fileName, const double *grid, const double h, const int N)" \
        in open(os.path.join(d, "postprocess.h")).read()


def test_compat_host_helpers_match_reference(tmp_path, ref):
    """updateEdgeValues / setupBoundaryConditions / VTK writer of the drop-in
    headers are host C: compile them WITHOUT libmgb calls being reached and
    compare with the reference's (oracle/_ref).  (GaussSeidelSmoother runs on the
    GPU: tests/test_gs_lex.py.)"""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    import numpy as np
    from oracle_lib import c_dp, seeded
    src = tmp_path / "h.c"
    src.write_text('#define GRID_LENGTH (1.)\n#include "mg_3d.h"\n#include "postprocess.h"\n')
    so = tmp_path / "libcompat_host.so"
    inc = [f"-I{ROOT}/multigrid_parallel_b200/compat", f"-I{ROOT}/include"]
    subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
                    *inc, "-o", str(so), str(src), f"-L{ROOT}/multigrid_parallel_b200", "-lmgb",
                    f"-Wl,-rpath,{ROOT}/multigrid_parallel_b200", "-lm"], check=True)
    L = C.CDLL(str(so))
    N = 9
    h = 1.0 / (N - 1)
    a = seeded((N,) * 3, 51)
    b = a.copy()
    L.updateEdgeValues.argtypes = [c_dp, C.c_int]
    ref.L.updateEdgeValues.argtypes = [c_dp, C.c_int]
    L.updateEdgeValues(a.ctypes.data_as(c_dp), N)
    ref.L.updateEdgeValues(b.ctypes.data_as(c_dp), N)
    assert np.array_equal(a, b)
    L.setupBoundaryConditions.argtypes = [c_dp, C.c_int, C.c_double]
    L.setupBoundaryConditions(a.ctypes.data_as(c_dp), N, h)
    ref.set_dirichlet(b, h)
    assert np.array_equal(a, b)
    L.writeOutputData.argtypes = [C.c_char_p, c_dp, C.c_double, C.c_int]
    os.environ["MGB_VTK_GPU"] = "0"  # the host formatter (the GPU one: tests/test_gpu_vtk.py)
    f1 = tmp_path / "a.vtk"
    L.writeOutputData(str(f1).encode(), a.ctypes.data_as(c_dp), h, N)
    # reference writer: not in libmg_ref (postprocess.h is not included there);
    # check the format against its printf strings (postprocess.h:13-19,31,44)
    lines = f1.read_text().splitlines()
    assert lines[4] == "DIMENSIONS 9 9 9" and lines[5] == "POINTS 729 float"
    assert lines[6 + 10] == "%10.8e %10.8e %10.8e" % (0.0, h * 1, h * 1)
    assert lines[6 + 729 + 4 + 5] == "%10.8e" % a.reshape(-1)[5]
