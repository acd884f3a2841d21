"""CPU restatement of DSOTracker's coarse-to-fine direct image alignment -- TEST INFRASTRUCTURE ONLY (NEXT #1).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import or execute this; the product (libcml_b200/) and tools/ never do.  Parity is pinned:
tests/test_tracker_oracle.py checks it against tests/golden/track_golden.cmlw, produced by the unmodified reference
(oracle/ref_driver.cpp --mode track, oracle/make_golden.py tracker).

Reference (under /root/reference/src/cml):
  build_pyramid        capture/CaptureImage.cpp:209-262, image/Array2D.h:388-401 (reduceByTwo), :288-331 (gradientImage)
  make_coarse_depth    optimization/dso/DSOTracker.cpp:494-725 (makeCoarseDepthL0)
  compute_residual     optimization/dso/DSOTracker.cpp:248-419
  compute_hessian      optimization/dso/DSOTracker.cpp:421-492, MatrixAccumulators.h:1135-1230 (Accumulator9::updateSSE_eighted)
  optimize             optimization/dso/DSOTracker.cpp:15-246
"""
import numpy as np

F32 = np.float32
DEFAULTS = dict(huber=9.0, cutoff=20.0, scale_rot=1.0, scale_trans=0.5, scale_a=10.0, scale_b=1000.0, optimize_a=True, optimize_b=True,
                saturated_th=0.33)
MAX_ITERATIONS = [10, 20, 50, 50, 50]


def gradient_image(gray):
    g = np.zeros(gray.shape + (3,), dtype=F32)
    g[1:-1, 1:-1, 0] = gray[1:-1, 1:-1]
    g[1:-1, 1:-1, 1] = (gray[1:-1, 2:] - gray[1:-1, :-2]) * F32(0.5)
    g[1:-1, 1:-1, 2] = (gray[2:, 1:-1] - gray[:-2, 1:-1]) * F32(0.5)
    return g


def reduce_by_two(gray):
    """Array2D::reduceByTwo: ((a + b) + c) + d) / 4 in fp32, floor(size / 2)."""
    h, w = gray.shape[0] // 2, gray.shape[1] // 2
    a = gray[0:2 * h:2, 0:2 * w:2]; b = gray[0:2 * h:2, 1:2 * w:2]; c = gray[1:2 * h:2, 0:2 * w:2]; d = gray[1:2 * h:2, 1:2 * w:2]
    return (((a + b).astype(F32) + c).astype(F32) + d).astype(F32) / F32(4)


def build_pyramid(gray, levels):
    """[(gray_l, grad_l)] for l = 0..levels-1."""
    out = []
    g = np.ascontiguousarray(gray, dtype=F32)
    for l in range(levels):
        if l:
            g = reduce_by_two(g)
        out.append((g, gradient_image(g)))
    return out


def level_K(K0, level):
    """getK(level) of the reference's PinholeUndistorter pyramid (map/InternalCalibration): fx / 2^l, (cx + 0.5) / 2^l - 0.5."""
    s = 2.0 ** level
    return np.array([K0[0] / s, K0[1] / s, (K0[2] + 0.5) / s - 0.5, (K0[3] + 0.5) / s - 0.5])


def se3_exp(xi):
    """Sophus SE3::exp, tangent = (translation, rotation)."""
    u = np.asarray(xi[:3], np.float64); w = np.asarray(xi[3:], np.float64)
    th2 = w @ w; th = np.sqrt(th2)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-10:
        R = np.eye(3) + W + 0.5 * W @ W
        V = np.eye(3) + 0.5 * W + W @ W / 6.0
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th2 * W @ W
        V = np.eye(3) + (1 - np.cos(th)) / th2 * W + (th - np.sin(th)) / (th2 * th) * W @ W
    return R, V @ u


def bilinear(grad, x, y):
    ix = x.astype(np.int32); iy = y.astype(np.int32)
    dx = (x - ix.astype(F32)).astype(F32); dy = (y - iy.astype(F32)).astype(F32)
    dxdy = (dx * dy).astype(F32)
    w00 = (F32(1) - dx - dy + dxdy).astype(F32); w10 = (dx - dxdy).astype(F32); w01 = (dy - dxdy).astype(F32); w11 = dxdy
    return (grad[iy, ix] * w00[:, None] + grad[iy, ix + 1] * w10[:, None] + grad[iy + 1, ix] * w01[:, None] + grad[iy + 1, ix + 1] * w11[:, None]).astype(F32)


def exposure_to(a0, b0, t0, a1, b1, t1):
    a = np.exp(a1 - a0) * t1 / t0
    return a, b1 - a * b0


def compute_residual(pc, grad, K, R, t, ref_exp, new_exp, level, cutoff, p):
    """Returns (stats dict, warped dict).  pc = [n,4] (u, v, idepth, color); R,t = refToNew (double); exposures = (tau, a, b)."""
    fx, fy, cx, cy = [F32(v) for v in K]
    hl, wl = grad.shape[:2]
    Ki = np.array([[1 / fx, 0, -cx / fx], [0, 1 / fy, -cy / fy], [0, 0, 1]], dtype=F32)
    RKi = (R.astype(F32) @ Ki).astype(F32); tf = t.astype(F32)
    a_ll, b_ll = exposure_to(ref_exp[1], ref_exp[2], ref_exp[0], new_exp[1], new_exp[2], new_exp[0])
    a_ll, b_ll = F32(a_ll), F32(b_ll)
    huber = F32(p["huber"]); cut = F32(cutoff)
    max_energy = F32(2) * huber * cut - huber * huber
    x, y, idp, col = pc[:, 0].astype(F32), pc[:, 1].astype(F32), pc[:, 2].astype(F32), pc[:, 3].astype(F32)
    ok = np.isfinite(col)
    ones = np.ones_like(x)
    xy1 = np.stack([x, y, ones], axis=1)
    pt = (xy1 @ RKi.T).astype(F32) + tf[None, :] * idp[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = (pt[:, 0] / pt[:, 2]).astype(F32); v = (pt[:, 1] / pt[:, 2]).astype(F32)
        Ku = (fx * u + cx).astype(F32); Kv = (fy * v + cy).astype(F32)
        new_id = (idp / pt[:, 2]).astype(F32)
    flow = np.zeros(3)
    if level == 0:
        sel = ok & (np.arange(x.size) % 32 == 0)
        if sel.any():
            def proj(M, sign):
                q = (xy1[sel] @ M.T).astype(F32) + F32(sign) * tf[None, :] * idp[sel, None]
                return (fx * (q[:, 0] / q[:, 2]) + cx).astype(F32), (fy * (q[:, 1] / q[:, 2]) + cy).astype(F32)
            KuT, KvT = proj(Ki, 1); KuT2, KvT2 = proj(Ki, -1); Ku3, Kv3 = proj(RKi, -1)
            xs, ys = x[sel], y[sel]
            sT = float(((KuT - xs) ** 2 + (KvT - ys) ** 2 + (KuT2 - xs) ** 2 + (KvT2 - ys) ** 2).astype(F32).sum(dtype=np.float64))
            sRT = float(((Ku[sel] - xs) ** 2 + (Kv[sel] - ys) ** 2 + (Ku3 - xs) ** 2 + (Kv3 - ys) ** 2).astype(F32).sum(dtype=np.float64))
            num = 2.0 * sel.sum()
            flow = np.array([sT / (num + 0.1), 0.0, sRT / (num + 0.1)])
    with np.errstate(invalid="ignore"):
        inb = ok & (Ku > 2) & (Kv > 2) & (Ku < wl - 3) & (Kv < hl - 3) & (new_id > 0)
    idx = np.nonzero(inb)[0]
    hit = bilinear(grad, Ku[idx], Kv[idx])
    fin = np.isfinite(hit).all(axis=1)
    idx = idx[fin]; hit = hit[fin]
    res = (hit[:, 0] - (a_ll * col[idx] + b_ll).astype(F32)).astype(F32)
    ar = np.abs(res)
    with np.errstate(divide="ignore"):
        hw = np.where(ar < huber, F32(1), huber / ar).astype(F32)
    sat = ar > cut
    e_in = (hw * res * res * (F32(2) - hw)).astype(F32)
    E = float(np.where(sat, max_energy, e_in).astype(F32).sum(dtype=np.float64))
    keep = ~sat
    wi = idx[keep]
    warped = dict(idepth=new_id[wi], u=u[wi], v=v[wi], dx=hit[keep, 1], dy=hit[keep, 2], residual=res[keep], weight=hw[keep], refcolor=col[wi])
    n = wi.size
    pad = (-n) % 4
    if pad:
        warped = {k: np.concatenate([a, np.zeros(pad, F32)]) for k, a in warped.items()}
    stats = dict(E=E, numTermsInE=int(idx.size), numSaturated=int(sat.sum()), numRobust=int((ar <= F32(p["cutoff"])).sum()), flow=flow)
    return stats, warped


def compute_hessian(wp, K, ref_exp, new_exp, p):
    fx, fy = F32(K[0]), F32(K[1])
    a, _ = exposure_to(ref_exp[1], ref_exp[2], ref_exp[0], new_exp[1], new_exp[2], new_exp[0])
    a = F32(a); b0 = F32(ref_exp[2])
    dx = (wp["dx"] * fx).astype(F32); dy = (wp["dy"] * fy).astype(F32)
    u, v, idp = wp["u"], wp["v"], wp["idepth"]
    one = F32(1)
    J = np.stack([idp * dx, idp * dy, -(idp * (u * dx + v * dy)), -((u * v * dx) + dy * (one + v * v)), (u * v * dy) + (dx * (one + u * u)), u * dy - v * dx,
                  a * (b0 - wp["refcolor"]), -np.ones_like(u), wp["residual"]], axis=1).astype(F32)
    n = J.shape[0]
    H9 = np.einsum("ni,n,nj->ij", J.astype(np.float64), wp["weight"].astype(np.float64), J.astype(np.float64))
    H = H9[:8, :8] / n; b = H9[:8, 8] / n
    s = np.array([p["scale_rot"]] * 3 + [p["scale_trans"]] * 3 + [p["scale_a"], p["scale_b"]], dtype=np.float64)
    s = s.astype(F32).astype(np.float64)
    return H * s[:, None] * s[None, :], b * s


def optimize(pcs, grads, Ks, ref_to_new, ref_exp, cur_exp, params=None, last_rmse=None):
    """DSOTracker::optimize.  pcs[l] [n,4], grads[l] [h,w,3] of the NEW frame, Ks[l] (fx,fy,cx,cy), ref_to_new = (R,t) initial,
    ref_exp / cur_exp = (tau, a, b).  Returns dict with the final ref_to_new, exposure and the Residual fields."""
    p = dict(DEFAULTS); p.update(params or {})
    L = len(grads)
    max_level = min(L - 1, 4)
    R, t = np.asarray(ref_to_new[0], np.float64), np.asarray(ref_to_new[1], np.float64)
    cur = list(cur_exp)
    old = dict(E=[0.0] * (max_level + 1), numTermsInE=[0] * (max_level + 1), numSaturated=[0] * (max_level + 1), numRobust=[0] * (max_level + 1), flow=np.zeros(3))
    new = dict(E=[0.0] * (max_level + 1), numTermsInE=[0] * (max_level + 1), numSaturated=[0] * (max_level + 1), numRobust=[0] * (max_level + 1), flow=np.zeros(3))
    rep = [0.0] * (max_level + 1)
    have_repeated = False
    H = np.eye(8); b = np.zeros(8)
    s = np.array([p["scale_rot"]] * 3 + [p["scale_trans"]] * 3 + [p["scale_a"], p["scale_b"]], dtype=np.float64).astype(F32).astype(np.float64)
    out = dict(isCorrect=False, iterations=0)

    def residual(level, Rr, tt, ex, cutoff):
        return compute_residual(pcs[level], grads[level], Ks[level], Rr, tt, ref_exp, ex, level, cutoff, p)

    def store(dst, st, level):
        for k in ("E", "numTermsInE", "numSaturated", "numRobust"):
            dst[k][level] = st[k]
        dst["flow"] = st["flow"]

    level = max_level
    while level >= 0:
        rep[level] = 1.0
        st, wp = residual(level, R, t, cur, p["cutoff"] * rep[level]); store(old, st, level)
        if old["numTermsInE"][level] < 20:
            return dict(out, **old, R=R, t=t, exposure=cur, levelCutoffRepeat=rep)
        while old["numSaturated"][level] / old["numTermsInE"][level] > 0.6 and rep[level] < 50:
            rep[level] *= 2
            st, wp = residual(level, R, t, cur, p["cutoff"] * rep[level]); store(old, st, level)
        if old["numTermsInE"][level] - old["numSaturated"][level] < 10:
            return dict(out, **old, R=R, t=t, exposure=cur, levelCutoffRepeat=rep)
        H, b = compute_hessian(wp, Ks[level], ref_exp, cur, p)
        lam = 0.01
        for it in range(MAX_ITERATIONS[level]):
            Hd = H.copy(); Hd[np.diag_indices(8)] *= (1 + lam)
            inc = np.zeros(8)
            if p["optimize_a"] and p["optimize_b"]:
                inc = np.linalg.solve(Hd, -b)
            elif p["optimize_a"]:
                inc[:7] = np.linalg.solve(Hd[:7, :7], -b[:7])
            elif p["optimize_b"]:
                Hs = Hd.copy(); bs = b.copy()
                Hs[:, 6] = Hs[:, 7]; Hs[6, :] = Hs[7, :]; bs[6] = bs[7]
                i7 = np.linalg.solve(Hs[:7, :7], -bs[:7])
                inc[:6] = i7[:6]; inc[7] = i7[6]
            else:
                inc[:6] = np.linalg.solve(Hd[:6, :6], -b[:6])
            if not np.isfinite(inc).all():
                return dict(out, **old, R=R, t=t, exposure=cur, levelCutoffRepeat=rep)
            if lam < 0.001:
                inc = inc * np.sqrt(np.sqrt(0.001 / lam))
            incs = inc * s
            dR, dt = se3_exp(incs[:6])
            Rn = dR @ R; tn = dR @ t + dt
            new_exp = [cur[0], cur[1] + incs[6], cur[2] + incs[7]]
            nst, nwp = residual(level, Rn, tn, new_exp, p["cutoff"] * rep[level]); store(new, nst, level)
            out["iterations"] += 1
            if p.get("trace") is not None:
                p["trace"].append((level, it, lam, nst["E"] / max(nst["numTermsInE"], 1), old["E"][level] / old["numTermsInE"][level], float(np.linalg.norm(inc))))
            accept = (nst["E"] / nst["numTermsInE"]) < (old["E"][level] / old["numTermsInE"][level]) if nst["numTermsInE"] > 0 else False
            if accept:
                H, b = compute_hessian(nwp, Ks[level], ref_exp, new_exp, p)
                # `oldResidual = newResidual` copies EVERY level: coarser levels inherit the last *tried* step there, accepted or not (DSOTracker.cpp:166)
                old = dict(E=list(new["E"]), numTermsInE=list(new["numTermsInE"]), numSaturated=list(new["numSaturated"]), numRobust=list(new["numRobust"]), flow=new["flow"])
                R, t = Rn, tn; cur = new_exp
                lam *= 0.5
            else:
                lam *= 4
            if np.linalg.norm(inc) < 1e-3:
                break
        if last_rmse is not None and old["E"][level] / old["numTermsInE"][level] > 1.5 * last_rmse[level]:      # mLastResidual.isCorrect branch, DSOTracker.cpp:190-196
            return dict(out, **old, R=R, t=t, exposure=cur, levelCutoffRepeat=rep)
        if rep[level] > 1 and not have_repeated:
            level += 1
            have_repeated = True
        level -= 1
    rel_a, rel_b = exposure_to(ref_exp[1], ref_exp[2], ref_exp[0], cur[1], cur[2], cur[0])
    good = True
    if p["optimize_a"]:
        good &= abs(cur[1]) <= 1.2
    else:
        good &= abs(np.log(F32(rel_a))) <= 1.5
    if p["optimize_b"]:
        good &= abs(cur[2]) <= 200
    else:
        good &= abs(F32(rel_b)) <= 200
    out.update(old)
    out.update(R=R, t=t, exposure=cur, levelCutoffRepeat=rep, isCorrect=bool(good), relAff=(rel_a, rel_b),
               tooManySaturated=not (old["numSaturated"][0] / old["numTermsInE"][0] > p["saturated_th"]), covariance=np.diag(np.linalg.inv(H))[:6])
    return out


def project_to_reference(K0, host_cams, ref_cam, pt_host, pt_xy, pt_idepth, pt_unc):
    """First loop of makeCoarseDepthL0 (DSOTracker.cpp:520-553): rows (Ku, Kv, new_idepth, uncertainty), all double.
    cams = world-to-camera [R(9) | t(3)] rows."""
    fx, fy, cx, cy = K0
    Rr, tr = ref_cam[:9].reshape(3, 3), ref_cam[9:]
    rows = []
    for h, xy, idp, unc in zip(pt_host, pt_xy, pt_idepth, pt_unc):
        Rh, th = host_cams[h][:9].reshape(3, 3), host_cams[h][9:]
        R = Rr @ Rh.T; t = tr - R @ th
        q = R @ np.array([(float(xy[0]) - cx) / fx, (float(xy[1]) - cy) / fy, 1.0]) + t * idp
        rows.append((fx * q[0] / q[2] + cx, fy * q[1] / q[2] + cy, idp / q[2], unc))
    return rows


def make_coarse_depth(points_uv_idepth_unc, gray_pyr):
    """makeCoarseDepthL0 given the points already projected into the reference frame: rows (Ku, Kv, new_idepth, uncertainty).
    gray_pyr[l] = reference gray image of level l.  Returns pcs[l] [n,4] (u, v, idepth, color) in raster order."""
    L = len(gray_pyr)
    h0, w0 = gray_pyr[0].shape
    idepth = [np.zeros(g.shape, F32) for g in gray_pyr]; wsum = [np.zeros(g.shape, F32) for g in gray_pyr]
    for Ku, Kv, nid, unc in points_uv_idepth_unc:
        u = int(Ku + 0.5); v = int(Kv + 0.5)
        wgt = F32(np.sqrt(F32(1e-3 / (unc + 1e-12))))
        if u < 0 or u >= w0 or v < 0 or v >= h0:
            continue
        idepth[0][v, u] += F32(float(nid) * float(wgt)); wsum[0][v, u] += wgt
    for l in range(1, L):
        hl, wl = gray_pyr[l].shape
        for arr in (idepth, wsum):
            a = arr[l - 1]
            arr[l] = (((a[0:2 * hl:2, 0:2 * wl:2] + a[0:2 * hl:2, 1:2 * wl:2]).astype(F32) + a[1:2 * hl:2, 0:2 * wl:2]).astype(F32) + a[1:2 * hl:2, 1:2 * wl:2]).astype(F32)
    for l in range(L):
        hl, wl = gray_pyr[l].shape
        offs = [wl + 1, -wl - 1, wl - 1, -wl + 1] if l < 2 else [1, -1, wl, -wl]
        idf = idepth[l].ravel(); wf = wsum[l].ravel(); bak = wf.copy()      # depth is only read where bak > 0 and written where bak <= 0
        size = wl * hl
        src_id = idf.copy()
        for i in range(wl, size - wl):
            if bak[i] <= 0:
                sm = F32(0); nm = F32(0); nn = 0
                for o in offs:
                    j = i + o
                    if 0 <= j < size and bak[j] > 0:
                        sm += src_id[j]; nm += bak[j]; nn += 1
                if nn:
                    idf[i] = sm / F32(nn); wf[i] = nm / F32(nn)
        idepth[l] = idf.reshape(hl, wl); wsum[l] = wf.reshape(hl, wl)
    pcs = []
    for l in range(L):
        hl, wl = gray_pyr[l].shape
        rows = []
        for y in range(2, hl - 2):
            for x in range(2, wl - 2):
                if wsum[l][y, x] > 0:
                    idp = F32(idepth[l][y, x] / wsum[l][y, x]); col = gray_pyr[l][y, x]
                    if np.isfinite(col) and idp > 0:
                        rows.append((x, y, idp, col))
        pcs.append(np.array(rows, dtype=F32).reshape(-1, 4))
    return pcs
