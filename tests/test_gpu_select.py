"""GPU parity of the pixel selector (cmlsel_*, SURVEY.md 8f NEXT #4, PixelSelector part) against the reference's golden vectors and the numpy
restatement.  Integer outputs: corners, types and potentials are compared exactly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_selector_matches_reference():
    from libcml_b200 import CaptureImageGenerator, PixelSelector
    w = cmlw.load(os.path.join(GOLDEN, "select_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "select_golden.cmlw"))
    W, H = w["size"]
    cap = CaptureImageGenerator(W, H).generate(w["gray"])
    sel = PixelSelector(W, H)
    for d, dens in enumerate(w["densities"]):
        assert sel.currentPotential == g[f"sel_pot_before{d}"][0]
        xy, ty = sel.compute(cap, dens)
        if d == 0:
            assert np.array_equal(sel.read("ths", (H // 32, W // 32)), g["sel_ths"])
            assert np.array_equal(sel.read("ths_smoothed", (H // 32, W // 32)), g["sel_ths_smoothed"])
        assert np.array_equal(xy, g[f"sel_xy{d}"]) and np.array_equal(ty, g[f"sel_type{d}"]), d
        assert sel.currentPotential == g[f"sel_pot_after{d}"][0]


def test_selector_matches_oracle_on_quantised_image():
    """8-bit image: many exactly-zero derivative components, i.e. candidates whose gradient is exactly orthogonal to an axis-aligned random
    direction -- the case in which the per-block selection counts depend on the directions and the prefix sum has to be iterated."""
    import prepare_oracle as P
    import select_oracle as S
    from libcml_b200 import CaptureImageGenerator, PixelSelector, CmlbaError, synth
    W, H = 200, 136                       # not multiples of 32: the flat threshold indexing of the last partial column is exercised
    win = synth.make_window(W, H, 2, 10, 1, False, seed=21, low_freq=True)
    gray = np.rint(win["gray"][0] / 4) * 4                    # coarse quantisation
    cap = CaptureImageGenerator(W, H).generate(gray.astype(np.float32))
    lv = P.prepare(gray.astype(np.float32), None, None, None, 5)
    levels = [(lv[l][1], lv[l][2]) for l in range(3)]
    ora = S.PixelSelector(W, H)
    sel = PixelSelector(W, H)
    for dens in (400.0, 80.0, 3000.0):
        xy, ty = sel.compute(cap, dens)
        oxy, oty = ora.compute(levels, dens)
        assert np.array_equal(xy, oxy) and np.array_equal(ty, oty) and sel.currentPotential == ora.pot, dens
    xy2, _ = sel.compute(cap, 3000.0, capacity=10)            # truncated output
    assert xy2.shape == (10, 2)
    with pytest.raises(CmlbaError):
        PixelSelector(32, 32)
    with pytest.raises(ValueError):
        PixelSelector(W + 32, H).compute(cap, 100.0)


@pytest.mark.parametrize("recursions,th_factor", [(0, 1.0), (2, 1.0), (1, 2.0), (1, 0.5)])
def test_selector_options_against_oracle(recursions, th_factor):
    """recursionsLeft and thFactor other than the defaults, 1-pixel potential (thread-per-block kernel) and large potentials (warp kernel)."""
    import prepare_oracle as P
    import select_oracle as S
    from libcml_b200 import CaptureImageGenerator, PixelSelector, synth
    W, H = 224, 160
    win = synth.make_window(W, H, 2, 10, 1, False, seed=33, low_freq=True, with_gradients=False)
    gray = win["gray"][0]
    cap = CaptureImageGenerator(W, H).generate(gray)
    lv = P.prepare(gray, None, None, None, 5)
    levels = [(lv[l][1], lv[l][2]) for l in range(3)]
    for pot0, dens in ((1, 5000.0), (2, 900.0), (3, 300.0), (7, 60.0)):
        sel = PixelSelector(W, H); sel.setPotential(pot0)
        ora = S.PixelSelector(W, H); ora.pot = pot0
        xy, ty = sel.compute(cap, dens, recursionsLeft=recursions, thFactor=th_factor)
        out = ora.make_maps(levels, dens, recursions, th_factor)
        sub = out[32:H - 32, 32:W - 32]
        ys, xs = np.nonzero(sub != 0)
        order = np.lexsort((ys, xs))
        oxy = np.stack([xs[order] + 32, ys[order] + 32], 1).astype(np.float32)
        assert np.array_equal(xy, oxy) and np.array_equal(ty, out[ys[order] + 32, xs[order] + 32]) and sel.currentPotential == ora.pot, (pot0, dens)
