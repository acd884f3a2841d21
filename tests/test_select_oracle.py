"""The numpy restatement of the DSO pixel selector (oracle/select_oracle.py) against the reference's golden vectors (tests/golden/select_*.cmlw,
made by oracle/make_golden.py select from the unmodified reference).  CPU only.  Integer outputs: everything is compared exactly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402
import prepare_oracle as P  # noqa: E402
import select_oracle as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_selector_matches_reference():
    w = cmlw.load(os.path.join(GOLDEN, "select_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "select_golden.cmlw"))
    W, H = w["size"]
    lv = P.prepare(w["gray"], None, None, None, 5)
    levels = [(lv[l][1], lv[l][2]) for l in range(3)]
    ths, sm = S.make_hists(levels[0][1])
    assert np.array_equal(ths, g["sel_ths"]) and np.array_equal(sm, g["sel_ths_smoothed"])
    sel = S.PixelSelector(W, H)
    for d, dens in enumerate(w["densities"]):
        assert sel.pot == g[f"sel_pot_before{d}"][0]
        xy, ty = sel.compute(levels, dens)
        assert np.array_equal(xy, g[f"sel_xy{d}"]) and np.array_equal(ty, g[f"sel_type{d}"]), d
        assert sel.pot == g[f"sel_pot_after{d}"][0]


def test_random_pattern_is_the_reference_lcg():
    p = S.random_pattern(5)
    state, want = 777, []
    for _ in range(5):
        state = (state * 1664525 + 1013904223) % 2 ** 32
        want.append(state >> 24)
    assert list(p) == want
