"""CPU-side check of the launch plans of the TMA tile kernels (tile.cu) through
mgb_tile_plan: for every level size a hierarchy can have, a plan that says `ok`
must fit the kernels' compile-time limits (384 threads, 112 KB of dynamic shared
memory so that two blocks share an SM, TMA box extents <= 256) and its tiles and
chunks must cover the whole level.  No GPU needed."""
import ctypes as C

import pytest

KINDS = {0: "residual norm", 1: "residual+restrict", 2: "half-sweep", 3: "prolongation"}
SIZES = [(n, n, n) for n in (9, 17, 33, 65, 129, 257, 513, 1025, 2049)] + [
    (4097, 513, 513), (1025, 513, 513), (33, 65, 129), (129, 65, 65), (33, 33, 257),
    (257, 1025, 129), (65, 17, 4097), (50, 50, 50), (6, 9, 12)]


def _plan(lib, kind, shape):
    out = (C.c_longlong * 12)()
    rc = lib.mgb_tile_plan(kind, *shape, out)
    return rc, list(out)


@pytest.mark.parametrize("shape", SIZES)
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_tile_plans_cover_the_level(mgb, kind, shape):
    lib = mgb.load_library()
    ni, nj, nk = shape
    if kind == 1 and not (ni & 1 and nj & 1 and nk & 1):
        rc, _ = _plan(lib, kind, shape)
        assert rc != 0  # restriction needs odd extents: refused with a message
        return
    rc, (ok, gx, gy, gz, threads, smem, trt, tqt, tro, tqo, ty, chunk) = _plan(lib, kind, shape)
    assert rc == 0
    if not ok:
        return  # the launcher falls back to the marching kernels
    assert 32 <= threads <= 384 and threads % 32 == 0
    assert trt * tqt <= threads
    assert smem <= 112 * 1024
    assert 2 * (tqt + 2) <= 256 and trt + 2 <= 256  # TMA box extents
    nq = (nk + 3) // 4
    if kind == 1:
        ncj, nck, nci = (nj + 1) // 2, (nk + 1) // 2, (ni + 1) // 2
        assert gx * tqo >= (nck + 1) // 2       # coarse quads
        assert gy * ty >= ncj and trt == 2 * ty + 1 and tqt == tqo + 1
        assert gz * chunk >= nci
    else:
        assert gx * tqo >= nq and gy * tro >= nj
        planes = ni if kind == 3 else ni - 2
        assert gz * chunk >= planes
        if kind == 3:
            assert chunk % 2 == 0  # whole (even, odd) plane pairs
    assert gx >= 1 and gy >= 1 and 1 <= gz <= 65535 and gy <= 65535


def test_headline_levels_use_the_fixed_shapes(mgb):
    """513^3 and 1025^3 get the tile extents the kernels are specialised for"""
    lib = mgb.load_library()
    for n in (513, 1025):
        assert _plan(lib, 0, (n, n, n))[1][6:8] == [5, 43]
        assert _plan(lib, 1, (n, n, n))[1][6:8] == [11, 34]
        assert _plan(lib, 2, (n, n, n))[1][6:8] == [6, 43]
        assert _plan(lib, 3, (n, n, n))[1][6:8] == [6, 43]
