"""GPU parity of the image preparation (cmlimg_*, SURVEY.md 8f NEXT #3) against the reference's golden vectors and the numpy restatement."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")
ATOL = 1e-4          # 0..255 intensity scale; the reference binary contracts FMAs, plain fp32 differs by a few ulp (tests/test_prepare_oracle.py)


def test_prepare_matches_reference_and_oracle():
    import prepare_oracle as P
    from libcml_b200 import CaptureImageGenerator
    w = cmlw.load(os.path.join(GOLDEN, "prepare_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "prepare_golden.cmlw"))
    Wi, Hi = w["size_in"]; Wo, Ho = w["size_out"]
    gen = CaptureImageGenerator(Wi, Hi, Wo, Ho)
    gen.setLut(w["lut"]); gen.setInverseVignette(w["inv_vignette"]); gen.setUndistortMap(g["prep_map"])
    cap = gen.generate(w["raw"])
    L = cap.getPyramidLevels()
    assert [list(s) for s in gen.sizes] == g["prep_levels_wh"].reshape(-1, 2).tolist()
    ora = P.prepare(w["raw"], w["lut"], w["inv_vignette"], g["prep_map"], L)
    for l in range(L):
        gray, grad, wgn = cap.getGrayImage(l), cap.getDerivativeImage(l), cap.getWeightedGradientNorm(l)
        assert np.abs(gray - g[f"prep_gray{l}"]).max() <= ATOL and np.abs(grad - g[f"prep_grad{l}"]).max() <= ATOL
        np.testing.assert_allclose(wgn, g[f"prep_wgn{l}"], rtol=1e-4, atol=1e-3)
        # against the plain-fp32 restatement the device is bit-exact (same operation order, no contraction)
        assert np.array_equal(gray, ora[l][0]) and np.array_equal(grad, ora[l][1]) and np.array_equal(wgn, ora[l][2])
    # the zero-copy input path gives the same bits
    gen.inputBuffer()[...] = w["raw"]
    cap2 = gen.generate(gen.inputBuffer())
    assert np.array_equal(cap2.getGrayImage(0), ora[0][0])


def test_prepare_variants_and_errors():
    """No LUT / no vignette / no map (identity sampling), odd sizes, 6 levels; bad arguments."""
    import prepare_oracle as P
    from libcml_b200 import CaptureImageGenerator, CmlbaError
    rng = np.random.default_rng(4)
    W, H = 203, 131
    raw = rng.uniform(0, 254, (H, W)).astype(np.float32)
    gen = CaptureImageGenerator(W, H, levels=6)
    cap = gen.generate(raw)
    ora = P.prepare(raw, None, None, None, 6)
    for l in range(6):
        assert np.array_equal(cap.getGrayImage(l), ora[l][0]) and np.array_equal(cap.getDerivativeImage(l), ora[l][1])
        assert np.array_equal(cap.getWeightedGradientNorm(l), ora[l][2])
    vig = rng.uniform(0.8, 1.3, (H, W)).astype(np.float32)
    gen.setInverseVignette(vig)
    ora = P.prepare(raw, None, vig, None, 6)
    assert np.array_equal(gen.generate(raw).getGrayImage(0), ora[0][0])
    # 8-bit sensor image: identical to preparing the same values as floats
    raw8 = rng.integers(0, 256, (H, W)).astype(np.uint8)
    a = gen.generate(raw8).getGrayImage(0); b = gen.generate(raw8.astype(np.float32)).getGrayImage(0)
    assert np.array_equal(a, b) and np.array_equal(a, P.prepare(raw8.astype(np.float32), None, vig, None, 6)[0][0])
    with pytest.raises(ValueError):
        gen.generate(raw[:, :-1])
    with pytest.raises(CmlbaError):
        CaptureImageGenerator(W, H, 160, 120).generate(raw)                  # sizes differ and no map was set
    bad = np.full((120, 160, 2), 500.0, np.float32)
    with pytest.raises(CmlbaError):
        CaptureImageGenerator(W, H, 160, 120).setUndistortMap(bad)           # map points outside the sensor image
