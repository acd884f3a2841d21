"""GPU parity of the FAST-9 detector (cmlfast_*, SURVEY.md 8f NEXT #4, first unit of the ORB extractor) against the reference's golden vectors and the
numpy restatement.  Integer outputs: exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402

pytestmark = pytest.mark.gpu


def test_fast_matches_reference():
    from libcml_b200 import FAST
    g = cmlw.load(os.path.join(ROOT, "tests", "golden", "fast_golden.cmlw"))
    H, W = g["gray_u8"].shape
    det = FAST(W, H)
    for k, th in enumerate(g["thresholds"]):
        xy, sc = det.compute(g["gray_u8"], int(th))
        assert np.array_equal(xy, g[f"fast_xy{k}"]) and np.array_equal(sc, g[f"fast_score{k}"]), k


def test_fast_random_images_against_oracle():
    """Noise images (dense corners, many score ties -> mutual suppression), tiny and odd sizes, saturated pixels, truncated output, errors."""
    import fast_oracle as F
    from libcml_b200 import FAST, CmlbaError
    rng = np.random.default_rng(11)
    det = FAST(333, 257)
    for (h, w, lo, hi, th) in ((257, 333, 0, 256, 20), (64, 64, 100, 140, 7), (7, 7, 0, 256, 5), (9, 40, 0, 256, 1), (120, 50, 0, 2, 1), (90, 90, 250, 256, 3)):
        img = rng.integers(lo, hi, (h, w)).astype(np.uint8)
        img[rng.integers(0, h, 20), rng.integers(0, w, 20)] = 255
        xy, sc = det.compute(img, th)
        oxy, osc = F.compute(img, th)
        assert np.array_equal(xy, oxy) and np.array_equal(sc, osc), (h, w, th)
    img = rng.integers(0, 256, (200, 300)).astype(np.uint8)
    full_xy, _ = det.compute(img, 10)
    part_xy, _ = det.compute(img, 10, capacity=5)
    assert part_xy.shape == (5, 2) and np.array_equal(part_xy, full_xy[:5]) and det.last_count == full_xy.shape[0]
    with pytest.raises(CmlbaError):
        det.compute(np.zeros((300, 400), np.uint8), 10)          # larger than the handle
    with pytest.raises(CmlbaError):
        FAST(4, 4)
