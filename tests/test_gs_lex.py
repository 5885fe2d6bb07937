"""Lexicographic Gauss-Seidel (GaussSeidelSmoother, /root/reference/mg_3d.h:546-637;
driver test_gs_3d.c) -- SURVEY section 8 row f4.

CPU: the oracle's restatement against the reference's own compiled routine and the
committed golden hashes; a numpy model of the pipelined hyperplane schedule the GPU
kernel runs (csrc/gslex.cu) against the serial loop, bit for bit.
GPU: mgb_gs_lex / mgb_host_gs_lex / the drop-in GaussSeidelSmoother against the oracle,
bit for bit, cubes and boxes."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import c_dp, seeded

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---------------------------------------------------------------- CPU
@pytest.mark.parametrize("N,iters", [(9, 1), (17, 2), (33, 3)])
def test_oracle_gs_lex_vs_compiled_reference(orc, ref, N, iters):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 61), seeded((N,) * 3, 62)
    w = v.copy()
    orc.gs_lex(v, d, h, iters, edges=True)
    ref.gs_lex(w, d, h, iters)
    assert np.array_equal(v, w)
    a = seeded((N,) * 3, 63)
    b = a.copy()
    orc.edge_values(a)
    ref.edge_values(b)
    assert np.array_equal(a, b)


def test_oracle_gs_lex_golden(orc):
    """hashes produced by the reference's GaussSeidelSmoother (oracle/gen_golden.py)"""
    ops = json.load(open(os.path.join(GOLD, "operators.json")))
    N = 17
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 1), seeded((N,) * 3, 2)
    assert sha(v) == ops["inputs"]["v"] and sha(d) == ops["inputs"]["d"]
    orc.gs_lex(v, d, h, 2, edges=True)
    assert sha(v) == ops["gauss_seidel_smoother_2"]
    orc.gs_lex(v, d, h, 1, edges=True)
    assert sha(v) == ops["then_gauss_seidel_smoother_1"]


def wavefront_model(v, d, h, iters):
    """the schedule of k_gs_lex in numpy: at step tau sweep s updates ALL points of the
    hyperplane i+j+k = 3 + tau - 2s at once from a snapshot (same expression, same order
    of additions as gs_point)"""
    ni, nj, nk = v.shape
    hSq, sixth = h * h, 1.0 / 6
    hmin, hmax = 3, ni + nj + nk - 6
    I, J, K = np.meshgrid(np.arange(1, ni - 1), np.arange(1, nj - 1), np.arange(1, nk - 1),
                          indexing="ij")
    H = I + J + K
    planes = {hh: (I[H == hh], J[H == hh], K[H == hh]) for hh in range(hmin, hmax + 1)}
    for tau in range(hmax - hmin + 1 + 2 * (iters - 1)):
        snap = v.copy()  # every sweep of this step reads the state before the step
        for s in range(iters):
            hh = hmin + tau - 2 * s
            if hh < hmin or hh > hmax:
                continue
            i, j, k = planes[hh]
            t = snap[i - 1, j, k] + snap[i + 1, j, k]
            t = t + snap[i, j - 1, k]
            t = t + snap[i, j + 1, k]
            t = t + snap[i, j, k - 1]
            t = t + snap[i, j, k + 1]
            v[i, j, k] = sixth * (t - hSq * d[i, j, k])


@pytest.mark.parametrize("shape,iters", [((9, 9, 9), 1), ((7, 12, 5), 3), ((17, 9, 13), 2),
                                         ((5, 5, 21), 4)])
def test_wavefront_schedule_equals_serial_sweep(orc, shape, iters):
    h = 0.37
    v, d = seeded(shape, 71), seeded(shape, 72)
    w = v.copy()
    orc.gs_lex(v, d, h, iters, edges=False)
    wavefront_model(w, d, h, iters)
    assert np.array_equal(v, w)


def tile_schedule_model(v, d, h, iters, tile, workers, seed):
    """the TILE-level schedule of k_gs_lex_tile in Python: work items (sweep, a, b, c) are
    handed out by a ticket counter in the order of a+b+c+2*sweep; a worker that holds an
    item runs it only once the flags say so -- the three "minus" tiles at this sweep, the
    three "plus" tiles and the tile itself at the previous one -- and otherwise waits while
    other workers (picked at random) go on.  A tile is swept lexicographically from the
    array as it is at that moment.  Returns the number of scheduling steps."""
    ni, nj, nk = v.shape
    TI, TJ, TK = tile
    nt = [-(-(n - 2) // t) for n, t in zip((ni, nj, nk), tile)]
    items = sorted(((s, a, b, c) for s in range(iters) for a in range(nt[0])
                    for b in range(nt[1]) for c in range(nt[2])),
                   key=lambda w: w[1] + w[2] + w[3] + 2 * w[0])
    done = np.zeros(nt, dtype=int)
    rng = np.random.default_rng(seed)
    hSq, sixth = h * h, 1.0 / 6
    ticket, holding, steps, running = 0, {}, 0, set()

    def ready(s, a, b, c):
        for axis in range(3):
            for plus in (0, 1):
                n = [a, b, c]
                n[axis] += 1 if plus else -1
                if 0 <= n[axis] < nt[axis] and done[tuple(n)] < (s if plus else s + 1):
                    return False
        return done[a, b, c] >= s

    while ticket < len(items) or holding:
        for wk in range(workers):  # idle workers draw tickets, in worker order
            if wk not in holding and ticket < len(items):
                holding[wk] = items[ticket]
                ticket += 1
        # items are in flight for a while: starting one must never find the same tile or a
        # face neighbour still being worked on (that is what makes a tile's sweep atomic)
        startable = [wk for wk, it in holding.items() if wk not in running and ready(*it)]
        if startable and (not running or rng.integers(2)):
            wk = startable[rng.integers(len(startable))]
            _, a, b, c = holding[wk]
            for other in running:
                _, a2, b2, c2 = holding[other]
                assert abs(a - a2) + abs(b - b2) + abs(c - c2) > 1, "overlapping tiles in flight"
            running.add(wk)
            continue
        assert running, "deadlock: every worker waits"
        wk = sorted(running)[rng.integers(len(running))]
        running.discard(wk)
        s, a, b, c = holding.pop(wk)
        for i in range(1 + a * TI, min(1 + (a + 1) * TI, ni - 1)):
            for j in range(1 + b * TJ, min(1 + (b + 1) * TJ, nj - 1)):
                for k in range(1 + c * TK, min(1 + (c + 1) * TK, nk - 1)):
                    t = v[i - 1, j, k] + v[i + 1, j, k]
                    t = t + v[i, j - 1, k]
                    t = t + v[i, j + 1, k]
                    t = t + v[i, j, k - 1]
                    t = t + v[i, j, k + 1]
                    v[i, j, k] = sixth * (t - hSq * d[i, j, k])
        done[a, b, c] = s + 1
        steps += 1
    return steps


@pytest.mark.parametrize("shape,tile,iters,workers", [((9, 8, 11), (2, 3, 4), 3, 4),
                                                      ((6, 6, 6), (8, 8, 8), 4, 3),   # one tile
                                                      ((12, 5, 7), (3, 2, 2), 2, 16),
                                                      ((5, 13, 6), (1, 4, 2), 3, 2)])
def test_tile_schedule_equals_serial_sweep(orc, shape, tile, iters, workers):
    """whatever order the flags allow, the ticket-ordered tile wavefront gives the serial
    result, and no set of workers can end up waiting for each other"""
    h = 0.21
    for seed in range(3):
        v, d = seeded(shape, 75), seeded(shape, 76)
        w = v.copy()
        orc.gs_lex(v, d, h, iters, edges=False)
        tile_schedule_model(w, d, h, iters, tile, workers, seed)
        assert np.array_equal(v, w), (shape, tile, seed)


# ---------------------------------------------------------------- GPU
LEX_MODES = {"auto": 1, "hyperplanes": 0, "t8x16x32": 2, "t8x8x32": 3, "t8x32x32": 4, "t16x16x32": 5}


@pytest.fixture(params=list(LEX_MODES))
def lex_kernel(request, mgb):
    """every GPU test of the sweep runs with the kernel the library would pick and with
    each kernel forced (global hyperplanes, the tile wavefront in four tile shapes)"""
    mgb.set_global(mgb.G_GSLEX_TILE, LEX_MODES[request.param])
    yield request.param
    mgb.set_global(mgb.G_GSLEX_TILE, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("coarse,levels", [((3, 3, 3), 5), ((5, 5, 5), 3), ((3, 5, 9), 3),
                                           ((9, 3, 5), 2), ((3, 3, 3), 6)])
@pytest.mark.parametrize("iters", [1, 2, 5])
def test_gpu_gs_lex_bitwise(mgb, orc, lex_kernel, coarse, levels, iters):
    with mgb.Solver(coarse, levels, 2) as s:
        for lvl in range(levels):
            shape = s.dims(lvl)
            if min(shape) < 3:
                continue
            h = s.spacing(lvl)
            v, d = seeded(shape, 81 + lvl), seeded(shape, 91 + lvl)
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            s.gs_lex(lvl, iters)
            orc.gs_lex(v, d, h, iters, edges=False)
            assert np.array_equal(s.download(lvl, mgb.MGB_U), v), (shape, iters)
            # ... and the updateEdgeValues that ends the reference routine
            s.edge_values(lvl, mgb.MGB_U)
            orc.edge_values(v)
            assert np.array_equal(s.download(lvl, mgb.MGB_U), v)


@pytest.mark.gpu
def test_gpu_gs_lex_large_plane_many_blocks(mgb, orc, lex_kernel):
    """a level whose hyperplanes need every SM's block (grid-wide barrier path)"""
    with mgb.Solver((257, 129, 37), 1, 1) as s:
        shape = s.dims(0)
        v, d = seeded(shape, 5), seeded(shape, 6)
        s.upload(0, mgb.MGB_U, v)
        s.upload(0, mgb.MGB_D, d)
        s.gs_lex(0, 3)
        orc.gs_lex(v, d, s.spacing(0), 3, edges=False)
        assert np.array_equal(s.download(0, mgb.MGB_U), v)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(9, 9, 9), (12, 7, 20)])
def test_gpu_host_gs_lex(mgb, orc, shape):
    h = 0.125
    v, d = seeded(shape, 7), seeded(shape, 8)
    w = v.copy()
    mgb.host_gs_lex(w, d, h, 2, edges=True)
    orc.gs_lex(v, d, h, 2, edges=True)
    assert np.array_equal(v, w)


@pytest.mark.gpu
def test_gpu_gs_lex_golden(mgb):
    """against the hashes the reference's own GaussSeidelSmoother produced"""
    ops = json.load(open(os.path.join(GOLD, "operators.json")))
    N = 17
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 1), seeded((N,) * 3, 2)
    mgb.host_gs_lex(v, d, h, 2, edges=True)
    assert sha(v) == ops["gauss_seidel_smoother_2"]
    mgb.host_gs_lex(v, d, h, 1, edges=True)
    assert sha(v) == ops["then_gauss_seidel_smoother_1"]


@pytest.mark.gpu
@pytest.mark.parametrize("N,lazy", [(9, "1"), (65, "1"), (65, "0")])
def test_gpu_dropin_gauss_seidel_smoother(tmp_path, orc, N, lazy):
    """the drop-in header's GaussSeidelSmoother called like test_gs_3d.c:56 does --
    repeatedly on caller-owned arrays (device-resident session for the large case,
    staged for the small one)"""
    src = tmp_path / "h.c"
    src.write_text('#define GRID_LENGTH (1.)\n#include "mg_3d.h"\n')
    so = tmp_path / "libcompat_gs.so"
    inc = [f"-I{ROOT}/multigrid_parallel_b200/compat", f"-I{ROOT}/include"]
    subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
                    *inc, "-o", str(so), str(src), f"-L{ROOT}/multigrid_parallel_b200", "-lmgb",
                    f"-Wl,-rpath,{ROOT}/multigrid_parallel_b200", "-lm"], check=True)
    code = f"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, {HERE!r})
from oracle_lib import Orc, c_dp, seeded
L = C.CDLL({str(so)!r})
L.GaussSeidelSmoother.argtypes = [c_dp, c_dp, C.c_int, C.c_double, C.c_int]
L.calculateResidual.restype = C.c_double
L.calculateResidual.argtypes = [c_dp, c_dp, C.c_int, C.c_double, c_dp]
N = {N}; h = 1.0 / (N - 1)
# page-aligned like a large calloc, so that the session path can watch the pages
def alloc(seed):
    raw = np.zeros(N ** 3 + 1024); off = (-raw.ctypes.data // 8) % 512
    a = raw[off:off + N ** 3].reshape(N, N, N); a[...] = seeded((N,) * 3, seed); return raw, a
r1, v = alloc(11); r2, d = alloc(12)
w = v.copy(); orc = Orc()
for it in range(3):
    L.GaussSeidelSmoother(v.ctypes.data_as(c_dp), d.ctypes.data_as(c_dp), N, h, 1)
    n = L.calculateResidual(v.ctypes.data_as(c_dp), d.ctypes.data_as(c_dp), N, h, None)
    orc.gs_lex(w, d, h, 1, edges=True)
    ne = orc.residual(w, d, h)
    assert abs(n - ne) <= 1e-12 * ne, (n, ne)
assert np.array_equal(v, w)
# the red-black smoothers on the same resident pair (test_rb_gs_3d.c:56-81 flow; neither
# driver ever calls SolverInitialize)
L.preSmoother.argtypes = L.postSmoother.argtypes = [c_dp, c_dp, C.c_int, C.c_double, C.c_int]
L.preSmoother(v.ctypes.data_as(c_dp), d.ctypes.data_as(c_dp), N, h, 1)
L.postSmoother(v.ctypes.data_as(c_dp), d.ctypes.data_as(c_dp), N, h, 1)
orc.smooth(w, d, h, 1, True); orc.smooth(w, d, h, 1, False)
assert np.array_equal(v, w)
print("ok")
"""
    env = dict(os.environ, MGB_LAZY_SYNC=lazy, OMP_NUM_THREADS="1")
    p = subprocess.run(["python", "-X", "faulthandler", "-c", code], capture_output=True, text=True,
                       env=env, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
