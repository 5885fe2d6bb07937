"""CPU proof (oracle + numpy) of the two exact short-cuts the GPU cycle takes
relative to the reference's vcycle:

 1. a coarse level is not zeroed before pre-smoothing (mg_3d.h:1258-1259): its
    first RED half-sweep is evaluated with the six neighbours literally 0, and
    the following BLACK half-sweep overwrites every interior black point, so
    whatever the interior held before never matters -- provided the faces are 0;
 2. the prolongation corrects only the RED points (mg_3d.h:1000-1145 corrects all
    of them): the post-smoother's first half-sweep is BLACK, which overwrites
    every interior black point without reading it; black face points only get
    the reference's `+= 0.`.

Both must give the same BITS as the reference's sequence of operations."""
import numpy as np
import pytest

from oracle_lib import seeded

SHAPES = [(9, 9, 9), (17, 9, 33), (33, 33, 33)]


def _colour_masks(shape):
    i, j, k = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    red = ((i + j + k) & 1) == 1
    interior = np.ones(shape, bool)
    for ax, n in enumerate(shape):
        idx = [slice(None)] * 3
        idx[ax] = [0, n - 1]
        interior[tuple(idx)] = False
    return red, interior


@pytest.mark.parametrize("shape", SHAPES)
def test_zero_guess_first_sweep_needs_no_memset(orc, shape):
    h = 1.0 / (shape[2] - 1)
    d = seeded(shape, 1)
    red, interior = _colour_masks(shape)
    # the reference: zero the level, then RED, BLACK, RED, BLACK
    want = np.zeros(shape)
    orc.smooth(want, d, h, 2, True)
    # the GPU cycle: interior left over from the previous cycle (garbage), faces 0
    got = seeded(shape, 2) * 1e3
    got[~interior] = 0.0
    # first RED half-sweep with the guess taken as 0: gs_point(0,0,0,0,0,0,hSq,d,1/6),
    # same operation order as mg_3d.h:437-442
    s = np.zeros(shape)
    for _ in range(5):
        s = s + 0.0
    first = (1.0 / 6) * (s - (h * h) * d)
    sel = red & interior
    got[sel] = first[sel]
    orc.half_sweep(got, d, h, 0)       # BLACK reads red only, overwrites all interior black
    orc.smooth(got, d, h, 1, True)     # second iteration as usual
    assert np.array_equal(got, want)
    assert np.array_equal(np.signbit(got), np.signbit(want))


@pytest.mark.parametrize("shape", SHAPES)
def test_red_only_prolongation_then_black_sweep(orc, shape):
    coarse = tuple((n + 1) // 2 for n in shape)
    h = 1.0 / (shape[2] - 1)
    ec = seeded(coarse, 3)
    _, cint = _colour_masks(coarse)
    ec[~cint] = 0.0                    # coarse error has homogeneous Dirichlet faces
    ef0, d = seeded(shape, 4), seeded(shape, 5)
    ef0[0, 0, 0] = -0.0                # a face value the reference's `+= 0.` flips to +0.
    red, interior = _colour_masks(shape)
    # the reference: correct every point, then post-smooth (BLACK, RED) twice
    want = ef0.copy()
    orc.prolong_correct(ec, want)
    orc.smooth(want, d, h, 2, False)
    # the GPU cycle: correct the RED points only, black faces get `+ 0.`
    full = ef0.copy()
    orc.prolong_correct(ec, full)
    got = ef0.copy()
    got[red] = full[red]
    black_faces = ~red & ~interior
    got[black_faces] = got[black_faces] + 0.0
    orc.smooth(got, d, h, 2, False)
    assert np.array_equal(got, want)
    assert np.array_equal(np.signbit(got), np.signbit(want))
