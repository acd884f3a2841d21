"""CPU restatement of the reference's image preparation -- TEST INFRASTRUCTURE ONLY (NEXT #3).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import or execute this; the product (libcml_b200/) and tools/ never do.  Parity is pinned:
tests/test_prepare_oracle.py checks it against tests/golden/prepare_golden.cmlw, produced by the unmodified reference
(oracle/ref_driver.cpp --mode prepare, oracle/make_golden.py prepare).

Reference (under /root/reference/src/cml):
  lut_apply / lut_inverse   image/LookupTable.h:99-131 (operator()(float), computeInverse)
  prepare                   capture/CaptureImage.cpp:137-262 (generate: LUT, inverse vignette, removeDistortion, pyramid, derivative images,
                            weighted gradient norm), map/InternalCalibration.h:404-437 (removeDistortion), image/Array2DProxy.h:198-226
"""
import numpy as np

import tracker_oracle as T

F32 = np.float32


def lut_apply(values, x):
    i0 = x.astype(np.uint8)
    i1 = (i0 + np.uint8(1)).astype(np.uint8)          # wraps at 255 like the reference's uint8_t
    f = (x - i0.astype(F32)).astype(F32)
    return (values[i0] * (F32(1) - f) + values[i1] * f).astype(F32)


def lut_inverse(values):
    inv = np.arange(256, dtype=F32)
    for i in range(1, 255):
        for s in range(1, 255):
            if values[s] <= i and values[s + 1] >= i:
                inv[i] = F32(s) + F32(F32(i) - values[s]) / F32(values[s + 1] - values[s])
                break
    inv[0] = 0; inv[255] = 255
    return inv


def interpolate(img, x, y):
    """Array2D<float>::interpolate, vectorised (fp32, m00 w00 + m10 w10 + m01 w01 + m11 w11)."""
    ix = x.astype(np.int32); iy = y.astype(np.int32)
    dx = (x - ix.astype(F32)).astype(F32); dy = (y - iy.astype(F32)).astype(F32); dxdy = (dx * dy).astype(F32)
    w00 = (F32(1) - dx - dy + dxdy).astype(F32)
    return (((img[iy, ix] * w00).astype(F32) + (img[iy, ix + 1] * (dx - dxdy).astype(F32)).astype(F32)).astype(F32)
            + (img[iy + 1, ix] * (dy - dxdy).astype(F32)).astype(F32)).astype(F32) + (img[iy + 1, ix + 1] * dxdy).astype(F32)


def prepare(raw, lut, inv_vignette, undistort_map, levels):
    """Returns [(gray_l, grad_l [h][w][3], weighted_gradient_norm_l)] for every level."""
    img = np.ascontiguousarray(raw, dtype=F32)
    if lut is not None:
        img = lut_apply(np.asarray(lut, F32), img)
    if inv_vignette is not None:
        img = (img * inv_vignette).astype(F32)
    if undistort_map is not None:
        out = np.zeros(undistort_map.shape[:2], F32)
        fin = np.isfinite(undistort_map[..., 0])
        out[fin] = interpolate(img, undistort_map[..., 0][fin], undistort_map[..., 1][fin])
        img = out
    inv = lut_inverse(np.asarray(lut, F32)) if lut is not None else np.arange(256, dtype=F32)
    res = []
    for gray, grad in T.build_pyramid(img, levels):
        c = np.clip(np.floor(grad[..., 0] + F32(0.5)).astype(np.int32), 5, 250)       # lroundf for the non-negative intensities of this path
        gw = (inv[c + 1] - inv[c]).astype(F32)
        n2 = ((grad[..., 1] * grad[..., 1]).astype(F32) + (grad[..., 2] * grad[..., 2]).astype(F32)).astype(F32)
        res.append((gray, grad, ((n2 * gw).astype(F32) * gw).astype(F32)))
    return res
