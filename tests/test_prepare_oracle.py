"""The numpy restatement of the reference's image preparation (oracle/prepare_oracle.py) against the reference's golden vectors
(tests/golden/prepare_*.cmlw, made by oracle/make_golden.py prepare from the unmodified reference).  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402
import prepare_oracle as P  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
# the reference binary is built with FMA contraction (gcc -O2 -march=x86-64-v3): its LUT / interpolation sums differ from plain fp32 by a few ulp
ATOL = 1e-4          # on a 0..255 intensity scale = 4e-7 relative


def test_prepare_matches_reference():
    w = cmlw.load(os.path.join(GOLDEN, "prepare_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "prepare_golden.cmlw"))
    L = g["prep_levels_wh"].size // 2
    res = P.prepare(w["raw"], w["lut"], w["inv_vignette"], g["prep_map"], L)
    assert L == 5
    for l, (gray, grad, wgn) in enumerate(res):
        assert gray.shape == (g["prep_levels_wh"][2 * l + 1], g["prep_levels_wh"][2 * l])
        assert np.abs(gray - g[f"prep_gray{l}"]).max() <= ATOL
        assert np.abs(grad - g[f"prep_grad{l}"]).max() <= ATOL
        np.testing.assert_allclose(wgn, g[f"prep_wgn{l}"], rtol=1e-4, atol=1e-3)
    outside = ~np.isfinite(g["prep_map"][..., 0])
    assert outside.any() and (res[0][0][outside] == 0).all()           # pixels the map leaves undefined are zero


def test_lut_semantics():
    lut = (255.0 * (np.arange(256, dtype=np.float32) / 255.0) ** 0.8).astype(np.float32)
    x = np.array([0.0, 0.5, 17.25, 254.75, 255.0], np.float32)
    y = P.lut_apply(lut, x)
    assert y[0] == lut[0] and y[4] == lut[255]                         # 255 -> upper index wraps to 0 with weight 0
    assert abs(y[2] - (lut[17] * 0.75 + lut[18] * 0.25)) < 1e-5
    inv = P.lut_inverse(lut)
    assert inv[0] == 0 and inv[255] == 255 and np.all(np.diff(inv[5:252]) > 0)     # entries the search cannot bracket keep the identity default (reference quirk)
    assert np.abs(P.lut_apply(lut, inv[5:250]) - np.arange(5, 250)).max() < 1e-3
