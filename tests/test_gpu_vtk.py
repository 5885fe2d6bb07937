"""GPU: writeOutputData (/root/reference/postprocess.h:5-47) with the text produced on the
device (mgb_vtk_*, csrc/vtk.cu) -- byte for byte the file the reference's fprintf loop
writes.  Expected bytes come from the C library's own snprintf (ctypes), the reference's
formatter."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import seeded

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
libc = C.CDLL(None)
libc.snprintf.restype = C.c_int


def c_e8(x):
    buf = C.create_string_buffer(64)
    libc.snprintf(buf, C.c_size_t(64), b"%10.8e", C.c_double(float(x)))
    return buf.value


def reference_file(values, h):
    """postprocess.h:13-44, line by line"""
    ni, nj, nk = values.shape
    total = values.size
    out = [b"# vtk DataFile Version 2.0\nPotential data\nASCII\nDATASET STRUCTURED_GRID\n"
           b"DIMENSIONS %d %d %d\nPOINTS %d float\n" % (ni, nj, nk, total)]
    xs = [[c_e8(h * i) for i in range(n)] for n in (ni, nj, nk)]
    for i in range(ni):
        for j in range(nj):
            pre = xs[0][i] + b" " + xs[1][j] + b" "
            out.extend(pre + xs[2][k] + b"\n" for k in range(nk))
    out.append(b"\nPOINT_DATA %d\nSCALARS data float 1\nLOOKUP_TABLE default\n" % total)
    flat = values.reshape(-1)
    # '%.8e' % x is correctly rounded like glibc's printf for finite x (spot-checked below)
    out.extend((("%.8e\n" % x).encode() if np.isfinite(x) else c_e8(x) + b"\n") for x in flat)
    return b"".join(out)


def test_python_and_glibc_agree_on_finite_values():
    rng = np.random.default_rng(5)
    for x in np.concatenate([rng.uniform(-1, 1, 2000) * 10.0 ** rng.integers(-30, 30, 2000),
                             [0.0, -0.0, 10000000.25, 1000000005.0, 103 / 1024, 5e-324]]):
        assert ("%.8e" % x).encode() == c_e8(x)


@pytest.mark.parametrize("shape,h", [((5, 4, 3), 0.25), ((9, 9, 9), 1 / 8), ((33, 33, 33), 1 / 32),
                                     ((17, 6, 1025), 1 / 1024)])
def test_vtk_stream_small_grids(mgb, shape, h):
    v = seeded(shape, 31) * 10.0 ** np.random.default_rng(1).integers(-12, 2, shape)
    data, host_chunks = mgb.vtk_bytes(v, h)
    assert data == reference_file(v, h)
    assert host_chunks == 0  # everything formatted on the device


def test_vtk_stream_crosses_chunks_and_signs(mgb):
    """more than one 2^20-point chunk per section; negatives (16-byte lines) mixed with
    positives (15) so that the block offsets come from the scan"""
    shape = (130, 129, 65)  # 1.09 M points
    rng = np.random.default_rng(3)
    v = rng.uniform(-1, 1, shape) * 10.0 ** rng.integers(-15, 3, shape)
    v[rng.uniform(size=shape) < 0.1] = 0.0
    v[0, 0, :8] = [-0.0, 1.0, -1.0, 9.999999995, 9.9999999949, 0.1, 123456789.0, 1e-20]
    data, host_chunks = mgb.vtk_bytes(v, 1 / 128)
    assert hashlib.sha256(data).hexdigest() == hashlib.sha256(reference_file(v, 1 / 128)).hexdigest()
    assert host_chunks == 0


def test_vtk_exceptions_go_through_the_c_library(mgb):
    """values the device formatter must not decide itself: ties through a truncated power
    of ten, non-finite values, three-digit exponents, subnormals"""
    shape = (4, 8, 8)
    v = seeded(shape, 9)
    v.reshape(-1)[:10] = [1000000005.0, 1000000015.0, np.inf, -np.inf, np.nan, 1e-120, -3e150,
                          5e-324, 2.5e-310, 1e100]
    data, host_chunks = mgb.vtk_bytes(v, 0.5)
    assert data == reference_file(v, 0.5)
    assert host_chunks == 1  # the one value chunk fell back to snprintf
    # ... while exact ties through an exact power of ten stay on the device
    w = seeded(shape, 10)
    w.reshape(-1)[:6] = [10000000.25, -10000000.75, 103 / 1024, 0.5, 2.5e-7, 1.0000000005]
    data, host_chunks = mgb.vtk_bytes(w, 0.5)
    assert data == reference_file(w, 0.5) and host_chunks == 0


def test_dropin_write_output_data_uses_the_gpu(tmp_path):
    """compat/postprocess.h on its own (no mg_3d.h): same bytes as its host formatter"""
    src = tmp_path / "w.c"
    src.write_text('#include "postprocess.h"\n')
    so = tmp_path / "libw.so"
    inc = [f"-I{ROOT}/multigrid_parallel_b200/compat", f"-I{ROOT}/include"]
    subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", *inc, "-o", str(so),
                    str(src), f"-L{ROOT}/multigrid_parallel_b200", "-lmgb",
                    f"-Wl,-rpath,{ROOT}/multigrid_parallel_b200", "-lm"], check=True)
    code = f"""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, {HERE!r})
from oracle_lib import c_dp, seeded
L = C.CDLL({str(so)!r})
L.writeOutputData.argtypes = [C.c_char_p, c_dp, C.c_double, C.c_int]
N = 21
v = seeded((N,) * 3, 77)
os.environ["MGB_VTK_TIMING"] = "1"
L.writeOutputData({str(tmp_path / "gpu.vtk")!r}.encode(), v.ctypes.data_as(c_dp), 1.0 / (N - 1), N)
os.environ["MGB_VTK_GPU"] = "0"
L.writeOutputData({str(tmp_path / "host.vtk")!r}.encode(), v.ctypes.data_as(c_dp), 1.0 / (N - 1), N)
"""
    p = subprocess.run(["python", "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    assert "formatted on the GPU" in p.stderr
    a, b = (tmp_path / "gpu.vtk").read_bytes(), (tmp_path / "host.vtk").read_bytes()
    assert a == b and len(a) > 21 ** 3 * 60


def test_vtk_stream_equals_the_oracle_file(mgb, orc, tmp_path):
    """the device-formatted stream against the file oracle/mg_oracle.c writes (which the CPU
    suite pins to the reference's own writeOutputData)"""
    shape = (21, 34, 19)
    v = seeded(shape, 41) * 10.0 ** np.random.default_rng(2).integers(-14, 2, shape)
    orc.write_vtk(tmp_path / "o.vtk", v, 1 / 32)
    data, host_chunks = mgb.vtk_bytes(v, 1 / 32)
    assert data == (tmp_path / "o.vtk").read_bytes() and host_chunks == 0


def test_vtk_stream_golden(mgb):
    """... and against the hash of the reference's own file for the seeded 9^3 grid"""
    import json
    ops = json.load(open(os.path.join(HERE, "golden", "operators.json")))
    g = seeded((9,) * 3, 21) * 10.0 ** np.random.default_rng(21).integers(-12, 3, (9,) * 3)
    data, host_chunks = mgb.vtk_bytes(g, 0.125)
    assert hashlib.sha256(data).hexdigest() == ops["vtk_9_sha256"] and host_chunks == 0
