"""CPU restatement of the DSO pixel selector -- TEST INFRASTRUCTURE ONLY (NEXT #4, PixelSelector part).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import or execute this; the product (libcml_b200/) and tools/ never do.  Parity is pinned:
tests/test_select_oracle.py checks it against tests/golden/select_golden.cmlw, produced by the unmodified reference
(oracle/ref_driver.cpp --mode select, oracle/make_golden.py select).

Reference: /root/reference/src/cml/features/corner/PixelSelector.{h,cpp}
  random_pattern      PixelSelector.h:55-59, PixelSelector.cpp:12-13 (LCG, state 777)
  make_hists          PixelSelector.cpp:30-118 (computeHistQuantil, makeHists)
  select              PixelSelector.cpp:217-365
  make_maps           PixelSelector.cpp:121-213
  compute             PixelSelector.cpp:367-384
"""
import math

import numpy as np

F32 = np.float32
DIRECTIONS = np.array([[0, 1.0000], [0.3827, 0.9239], [0.1951, 0.9808], [0.9239, 0.3827], [0.7071, 0.7071], [0.3827, -0.9239], [0.8315, 0.5556], [0.8315, -0.5556],
                       [0.5556, -0.8315], [0.9808, 0.1951], [0.9239, -0.3827], [0.7071, -0.7071], [0.5556, 0.8315], [0.9808, -0.1951], [1.0000, 0.0000], [0.1951, -0.9808]], dtype=F32)


def random_pattern(n):
    out = np.empty(n, np.uint8)
    state = 777
    for i in range(n):
        state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = (state >> 24) & 0xFF
    return out


def hist_quantile(hist, below):
    th = int(math.floor(float(F32(hist[0]) * F32(below)) + 0.5))          # lroundf of a non-negative value
    for i in range(90):
        th -= hist[i + 1]
        if th < 0:
            return i
    return 90


def make_hists(wgn0):
    h, w = wgn0.shape
    w32, h32 = w // 32, h // 32
    ths = np.zeros((h32, w32), F32)
    for y in range(h32):
        for x in range(w32):
            blk = wgn0[32 * y:32 * y + 32, 32 * x:32 * x + 32]
            jt, it = np.mgrid[32 * y:32 * y + 32, 32 * x:32 * x + 32]
            ok = ~((it > w - 2) | (jt > h - 2) | (it < 1) | (jt < 1))
            gf = np.sqrt(blk.astype(F32))
            fin = np.isfinite(gf) & ok
            g = np.minimum(gf[fin].astype(np.int32), 48)
            hist = np.zeros(100, np.int64)
            np.add.at(hist, g + 1, 1)
            hist[0] = g.size
            ths[y, x] = F32(hist_quantile(hist, 0.5)) + F32(7)
    sm = np.zeros_like(ths)
    for y in range(h32):
        for x in range(w32):
            s = F32(0); n = F32(0)
            for dx in (-1, 1):            # the reference's order: left column (up, down, centre), right column, then up, down, centre
                if 0 <= x + dx < w32:
                    if y > 0: n += F32(1); s = F32(s + ths[y - 1, x + dx])
                    if y < h32 - 1: n += F32(1); s = F32(s + ths[y + 1, x + dx])
                    n += F32(1); s = F32(s + ths[y, x + dx])
            if y > 0: n += F32(1); s = F32(s + ths[y - 1, x])
            if y < h32 - 1: n += F32(1); s = F32(s + ths[y + 1, x])
            n += F32(1); s = F32(s + ths[y, x])
            sm[y, x] = F32(F32(s / n) * F32(s / n))
    return ths, sm


def select(levels, ths_smoothed, pattern, pot, th_factor=1.0):
    """levels[l] = (grad_l [h][w][3], wgn_l [h][w]) for l = 0, 1, 2.  Returns (map [h][w] float32, (n2, n3, n4))."""
    grad0, wgn0 = levels[0]
    wgn1, wgn2 = levels[1][1], levels[2][1]
    h, w = wgn0.shape
    out = np.zeros((h, w), F32)
    dw1 = F32(0.75); dw2 = F32(dw1 * dw1)
    thf = F32(th_factor)
    step = w // 32
    n2 = n3 = n4 = 0
    gx0 = grad0[..., 1]; gy0 = grad0[..., 2]
    for y4 in range(0, h, 4 * pot):
        for x4 in range(0, w, 4 * pot):
            my3 = min(4 * pot, h - y4); mx3 = min(4 * pot, w - x4)
            best4 = -1; val4 = F32(0)
            dir4 = DIRECTIONS[pattern[n2] & 0xF]
            for y3 in range(0, my3, 2 * pot):
                for x3 in range(0, mx3, 2 * pot):
                    x34 = x3 + x4; y34 = y3 + y4
                    my2 = min(2 * pot, h - y34); mx2 = min(2 * pot, w - x34)
                    best3 = -1; val3 = F32(0)
                    dir3 = DIRECTIONS[pattern[n2] & 0xF]
                    for y2 in range(0, my2, pot):
                        for x2 in range(0, mx2, pot):
                            x234 = x2 + x34; y234 = y2 + y34
                            my1 = min(pot, h - y234); mx1 = min(pot, w - x234)
                            best2 = -1; val2 = F32(0)
                            dir2 = DIRECTIONS[pattern[n2] & 0xF]
                            for y1 in range(my1):
                                for x1 in range(mx1):
                                    xf = x1 + x234; yf = y1 + y234
                                    idx = xf + w * yf
                                    if xf < 4 or xf >= w - 5 or yf < 4 or yf > h - 4:
                                        continue
                                    th0 = ths_smoothed[yf >> 5, xf >> 5] if (yf >> 5) < ths_smoothed.shape[0] and (xf >> 5) < ths_smoothed.shape[1] else _flat(ths_smoothed, (xf >> 5) + (yf >> 5) * step)
                                    th1 = F32(th0 * dw1); th2 = F32(th1 * dw2)
                                    ag0 = wgn0[yf, xf]
                                    if ag0 > F32(th0 * thf):
                                        dn = abs(F32(F32(gx0[yf, xf] * dir2[0]) + F32(gy0[yf, xf] * dir2[1])))
                                        if dn > val2:
                                            val2 = dn; best2 = idx; best3 = -2; best4 = -2
                                    if best3 == -2:
                                        continue
                                    ag1 = wgn1[int(F32(F32(yf) * F32(0.5)) + F32(0.25)), int(F32(F32(xf) * F32(0.5)) + F32(0.25))]
                                    if ag1 > F32(th1 * thf):
                                        dn = abs(F32(F32(gx0[yf, xf] * dir3[0]) + F32(gy0[yf, xf] * dir3[1])))
                                        if dn > val3:
                                            val3 = dn; best3 = idx; best4 = -2
                                    if best4 == -2:
                                        continue
                                    ag2 = wgn2[int(float(F32(yf) * F32(0.25)) + 0.125), int(float(F32(xf) * F32(0.25)) + 0.125)]
                                    if ag2 > F32(th2 * thf):
                                        dn = abs(F32(F32(gx0[yf, xf] * dir4[0]) + F32(gy0[yf, xf] * dir4[1])))
                                        if dn > val4:
                                            val4 = dn; best4 = idx
                            if best2 > 0:
                                out.flat[best2] = 1; val3 = F32(1e10); n2 += 1
                    if best3 > 0:
                        out.flat[best3] = 2; val4 = F32(1e10); n3 += 1
            if best4 > 0:
                out.flat[best4] = 4; n4 += 1
    return out, (n2, n3, n4)


def _flat(a, i):
    """thsSmoothed is a flat array with 100 spare zero entries: blocks beyond the last full 32-column/row read whatever index they hit."""
    f = a.ravel()
    return f[i] if i < f.size else F32(0)


class PixelSelector:
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.pattern = random_pattern(w * h)
        self.pot = 3

    def make_maps(self, levels, density, recursions=1, th_factor=1.0):
        _, sm = make_hists(levels[0][1])
        out, n = select(levels, sm, self.pattern, self.pot, th_factor)
        have = F32(n[0] + n[1] + n[2])
        with np.errstate(divide="ignore"):
            quotia = F32(F32(density) / have)
        K = F32(have * F32((self.pot + 1) * (self.pot + 1)))
        ideal = int(np.sqrt(F32(K / F32(density)))) - 1
        if ideal < 1:
            ideal = 1
        if recursions > 0 and quotia > 1.25 and self.pot > 1:
            if ideal >= self.pot:
                ideal = self.pot - 1
            self.pot = ideal
            return self.make_maps(levels, density, recursions - 1, th_factor)
        if recursions > 0 and quotia < 0.25:
            if ideal <= self.pot:
                ideal = self.pot + 1
            self.pot = ideal
            return self.make_maps(levels, density, recursions - 1, th_factor)
        if quotia < 0.95:
            char_th = int(F32(255.0) * quotia) & 0xFF
            nz = np.nonzero(out.ravel() != 0)[0]
            kill = self.pattern[:nz.size] > char_th
            out.flat[nz[kill]] = 0
        self.pot = ideal
        return out

    def compute(self, levels, density):
        """Returns (corners [n][2] in the reference's column-major emission order, types [n])."""
        out = self.make_maps(levels, density)
        grad0 = levels[0][0]
        sub = out[32:self.h - 32, 32:self.w - 32]
        fin = np.isfinite(grad0[32:self.h - 32, 32:self.w - 32]).all(axis=2)
        ys, xs = np.nonzero((sub != 0) & fin)
        order = np.lexsort((ys, xs))                 # i (x) outer, j (y) inner
        xs, ys = xs[order] + 32, ys[order] + 32
        return np.stack([xs, ys], 1).astype(F32), out[ys, xs]
