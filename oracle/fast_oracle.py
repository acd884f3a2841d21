"""CPU restatement of the FAST-9 corner detector -- TEST INFRASTRUCTURE ONLY (NEXT #4: first unit of the ORB extractor).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import or execute this; the product (libcml_b200/) and tools/ never do.  Parity is pinned:
tests/test_fast_oracle.py checks it against tests/golden/fast_golden.cmlw, produced by the unmodified reference
(oracle/ref_driver.cpp --mode fast, oracle/make_golden.py fast).

Reference: /root/reference/src/cml/features/corner/FAST.cpp
  is_corner   the generated decision tree of fast9_detect (:2985-5913) / fast9_corner_score (:15-2944): 9 contiguous circle pixels all
              brighter than p + b or all darker than p - b
  score       fast9_corner_score: bisection on b in [threshold, 255] -> the largest b for which the pixel is still a corner
  nonmax      nonmax_suppression (:5921-6033): a corner survives iff no 8-neighbour corner has a score >= its own; raster order
"""
import numpy as np

# make_offsets (FAST.cpp:2946-2964): (dx, dy) of the 16 circle pixels
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def _ring(img):
    """[16][h-6][w-6] circle pixels of every interior pixel, and the centres."""
    h, w = img.shape
    c = img[3:h - 3, 3:w - 3].astype(np.int32)
    ring = np.stack([img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int32) for dx, dy in CIRCLE])
    return ring, c


def _has_arc9(mask):
    """mask [16][...] bool: True where 9 contiguous (circular) entries are set."""
    m = np.concatenate([mask, mask[:8]])
    run = np.ones(mask.shape[1:], bool)
    out = np.zeros(mask.shape[1:], bool)
    for s in range(16):
        run = np.ones(mask.shape[1:], bool)
        for k in range(9):
            run &= m[s + k]
        out |= run
    return out


def is_corner(ring, c, b):
    return _has_arc9(ring > (c + b)[None]) | _has_arc9(ring < (c - b)[None])


def compute(img, threshold):
    """Features::FAST::compute: returns (xy [n][2] int32 in raster order, scores [n] int32)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    ring, c = _ring(img)
    corner = is_corner(ring, c, np.full(c.shape, threshold, np.int32))
    ys, xs = np.nonzero(corner)
    r = ring[:, ys, xs]; cc = c[ys, xs]
    bmin = np.full(ys.size, threshold, np.int32); bmax = np.full(ys.size, 255, np.int32)
    b = (bmax + bmin) // 2
    done = np.zeros(ys.size, bool)
    while not done.all():
        ok = is_corner(r, cc, b)
        bmin = np.where(~done & ok, b, bmin); bmax = np.where(~done & ~ok, b, bmax)
        done |= (bmin == bmax - 1) | (bmin == bmax)
        b = (bmin + bmax) // 2
    score = np.full((h, w), -1, np.int32)
    score[ys + 3, xs + 3] = bmin
    keep = np.ones(ys.size, bool)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx == 0 and dy == 0:
                continue
            yy = ys + 3 + dy; xx = xs + 3 + dx
            nb = score[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
            keep &= ~(nb >= bmin)           # neighbours that are not corners hold -1
    return np.stack([xs[keep] + 3, ys[keep] + 3], 1).astype(np.int32), bmin[keep]
