"""The numpy restatement of FAST-9 (oracle/fast_oracle.py) against the reference's golden vectors (tests/golden/fast_golden.cmlw, made by
oracle/make_golden.py fast from the unmodified reference).  CPU only; integer outputs: exact."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402
import fast_oracle as F  # noqa: E402


def test_fast_matches_reference():
    g = cmlw.load(os.path.join(ROOT, "tests", "golden", "fast_golden.cmlw"))
    for k, th in enumerate(g["thresholds"]):
        xy, sc = F.compute(g["gray_u8"], int(th))
        assert np.array_equal(xy, g[f"fast_xy{k}"]) and np.array_equal(sc, g[f"fast_score{k}"]), k
        assert (sc >= th).all()


def test_score_is_the_largest_threshold_that_still_detects():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (40, 50)).astype(np.uint8)
    xy, sc = F.compute(img, 10)
    ring, c = F._ring(img)
    for (x, y), s in list(zip(xy, sc))[:40]:
        r = ring[:, y - 3:y - 2, x - 3:x - 2]; cc = c[y - 3:y - 2, x - 3:x - 2]
        assert F.is_corner(r, cc, np.full((1, 1), s))[0, 0] and not F.is_corner(r, cc, np.full((1, 1), s + 1))[0, 0]
