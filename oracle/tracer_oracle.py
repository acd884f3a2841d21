"""CPU restatement of DSOTracer (immature-point tracing and activation) -- TEST INFRASTRUCTURE ONLY (NEXT #2).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import or execute this; the product (libcml_b200/) and tools/ never do.  Parity is pinned:
tests/test_tracer_oracle.py checks it against tests/golden/trace_golden.cmlw, produced by the unmodified reference
(oracle/ref_driver.cpp --mode trace, oracle/make_golden.py tracer).

Reference (under /root/reference/src/cml/optimization/dso):
  init_point              DSOTracer.cpp:538-583 (makeNewTracesFrom), DSOTracer.h:14-32 (DSOTracerPointPrivate defaults)
  trace                   DSOTracer.cpp:585-832
  linearize_residual      DSOTracer.cpp:413-494
  optimize_immature_point DSOTracer.cpp:280-411
"""
import math

import numpy as np

F32 = np.float32
PATTERN = [(0, -2), (-1, -1), (1, -1), (-2, 0), (0, 0), (2, 0), (-1, 1), (0, 2)]
IPS_GOOD, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED = range(6)
RES_IN, RES_OOB, RES_OUTLIER = 0, 1, 2
DEFAULTS = dict(huber=9.0, outlier_th=144.0, outlier_th_sum=2500.0, max_pix_search=float(F32(0.027)), max_slack_interval=1.5, step_size=1.0,
                min_improvement=2.0, test_radius=2.0, extra_slack=float(F32(1.2)), min_idepth_h_act=100.0, gn_iterations=3)


def interp(img, x, y):
    """Array2D::interpolate (image/Array2D.h:242-286): fp32, m00 w00 + m10 w10 + m01 w01 + m11 w11; img [h][w] or [h][w][c]."""
    x = F32(x); y = F32(y)
    ix = int(x); iy = int(y)
    dx = F32(x - F32(ix)); dy = F32(y - F32(iy)); dxdy = F32(dx * dy)
    return (img[iy, ix] * F32(F32(F32(1) - dx) - dy + dxdy) + img[iy, ix + 1] * F32(dx - dxdy) + img[iy + 1, ix] * F32(dy - dxdy) + img[iy + 1, ix + 1] * dxdy).astype(F32) \
        if img.ndim == 3 else F32(F32(F32(img[iy, ix] * F32(F32(F32(1) - dx) - dy + dxdy)) + F32(img[iy, ix + 1] * F32(dx - dxdy))) + F32(img[iy + 1, ix] * F32(dy - dxdy))) + F32(img[iy + 1, ix + 1] * dxdy)


def exposure_to(e0, e1):
    a = math.exp(e1[1] - e0[1]) * e1[0] / e0[0]
    return a, e1[2] - a * e0[2]


def rel_pose(c0, c1):
    R0, t0 = c0[:9].reshape(3, 3), c0[9:]
    R1, t1 = c1[:9].reshape(3, 3), c1[9:]
    R = R1 @ R0.T
    return R, t1 - R @ t0


class ImmaturePoint:
    def __init__(self, host, xy, grad_host, p=DEFAULTS):
        self.host = host
        self.xy = (float(xy[0]), float(xy[1]))
        self.status = IPS_UNINITIALIZED
        self.idmin = 1.0 / 1000.0
        self.idmax = float("nan")
        self.uv = (-1.0, -1.0)
        self.interval = -1.0
        self.quality = 10000.0
        gh = np.zeros((2, 2))
        self.weights = []
        c = F32(p["outlier_th_sum"])
        for sx, sy in PATTERN:
            g = interp(grad_host, self.xy[0] + sx, self.xy[1] + sy)[1:].astype(np.float64)
            gh += np.outer(g, g)
            self.weights.append(math.sqrt(float(c) / (float(c) + float(g @ g))))
        self.gradH = gh
        self.energyTH = 8 * float(F32(p["outlier_th"]))


def trace(pt, K, cam_host, cam_target, exp_host, exp_target, gray_host, gray_target, p=DEFAULTS):
    """DSOTracer::trace: updates pt in place, returns the status."""
    H, W = gray_target.shape
    if pt.status == IPS_OOB:
        return IPS_OOB
    fx, fy, cx, cy = K
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    R, t = rel_pose(cam_host, cam_target)
    KRKi = Km @ R @ np.linalg.inv(Km)
    Kt = Km @ t
    pr = KRKi @ np.array([pt.xy[0], pt.xy[1], 1.0])
    max_pix = float(W + H) * p["max_pix_search"]

    def inside(q, pad):
        return q[0] >= pad and q[1] >= pad and q[0] < W - pad and q[1] < H - pad

    def oob():
        pt.uv = (-1.0, -1.0); pt.interval = 0.0; pt.status = IPS_OOB
        return IPS_OOB

    ptp_min = pr + Kt * pt.idmin
    with np.errstate(all="ignore"):
        pmin = ptp_min[:2] / ptp_min[2]
    if not inside(pmin, 4):
        return oob()
    if math.isfinite(pt.idmax):
        ptp_max = pr + Kt * pt.idmax
        with np.errstate(all="ignore"):
            pmax = ptp_max[:2] / ptp_max[2]
        if not inside(pmax, 5):
            return oob()
        interval = float(np.linalg.norm(pmax - pmin))
        if interval < p["max_slack_interval"]:
            pt.uv = tuple((pmax + pmin) / 2.0); pt.interval = interval; pt.status = IPS_SKIPPED
            return IPS_SKIPPED
    else:
        interval = max_pix
        ptp_max = pr + Kt * 0.01
        pmax = ptp_max[:2] / ptp_max[2]
        d = pmax - pmin
        inv = 1.0 / float(np.linalg.norm(d))
        pmax = np.array([pmin[0] + interval * d[0] * inv, pmin[1] + interval * d[1] * inv])
        if not inside(pmax, 5):
            return oob()
    if not (pt.idmin < 0 or (0.75 < ptp_min[2] < 1.5)):
        return oob()
    dx = p["step_size"] * (pmax[0] - pmin[0]); dy = p["step_size"] * (pmax[1] - pmin[1])
    v1 = np.array([dx, dy]); v2 = np.array([dy, -dx])
    a = float(v1 @ (pt.gradH @ v1)); b = float(v2 @ (pt.gradH @ v2))
    with np.errstate(all="ignore"):
        err_px = float(F32(0.2)) + float(F32(0.2)) * (a + b) / a if a != 0 else float("nan") if (a + b) == 0 else math.copysign(float("inf"), (a + b))
    if err_px * p["min_improvement"] > interval and math.isfinite(pt.idmax):
        pt.uv = tuple((pmax + pmin) / 2.0); pt.interval = interval; pt.status = IPS_BADCONDITION
        return IPS_BADCONDITION
    if err_px > 10:
        err_px = 10.0
    dx /= interval; dy /= interval
    if interval > max_pix:
        interval = max_pix
    num_steps = int(float(F32(1.9999)) + interval / p["step_size"])
    Rplane = KRKi[:2, :2]
    rand_shift = pmin[0] * 1000 - math.floor(pmin[0] * 1000)
    ptx = F32(pmin[0] - rand_shift * dx); pty = F32(pmin[1] - rand_shift * dy)
    rot = [Rplane @ np.array(s, dtype=np.float64) for s in PATTERN]
    if not (math.isfinite(dx) and math.isfinite(dy)):
        pt.interval = 0.0; pt.uv = (-1.0, -1.0); pt.status = IPS_OOB
        return IPS_OOB
    a_t, b_t = exposure_to(exp_host, exp_target)
    ixc, iyc = int(pt.xy[0]), int(pt.xy[1])
    ref_col = [a_t * float(gray_host[iyc + sy, ixc + sx]) + b_t for sx, sy in PATTERN]
    huber = float(F32(p["huber"]))
    if num_steps >= 100:
        num_steps = 99
    errors = []
    best_u = best_v = 0.0; best_e = 1e10; best_i = -1
    for i in range(num_steps):
        e = 0.0
        for k in range(8):
            qx = float(ptx) + rot[k][0]; qy = float(pty) + rot[k][1]
            if not inside((qx, qy), 3):
                e += 1e5
                continue
            hit = float(interp(gray_target, qx, qy))
            r = hit - ref_col[k]
            hw = 1.0 if abs(r) < huber else huber / abs(r)
            e += hw * r * r * (2 - hw)
        errors.append(e)
        if e < best_e:
            best_u, best_v, best_e, best_i = float(ptx), float(pty), e, i
        ptx = F32(float(ptx) + dx); pty = F32(float(pty) + dy)
    second = 1e10
    rad = p["test_radius"]
    for i in range(num_steps):
        if (i < best_i - rad or i > best_i + rad) and errors[i] < second:
            second = errors[i]
    new_q = second / best_e if best_e != 0 else float("inf")
    if new_q < pt.quality or num_steps > 10:
        pt.quality = new_q
    if best_e >= pt.energyTH * p["extra_slack"]:
        pt.interval = 0.0; pt.uv = (-1.0, -1.0)
        pt.status = IPS_OOB if pt.status == IPS_OUTLIER else IPS_OUTLIER
        return pt.status
    with np.errstate(all="ignore"):
        if dx * dx > dy * dy:
            lo = (pr[2] * (best_u - err_px * dx) - pr[0]) / (Kt[0] - Kt[2] * (best_u - err_px * dx))
            hi = (pr[2] * (best_u + err_px * dx) - pr[0]) / (Kt[0] - Kt[2] * (best_u + err_px * dx))
        else:
            lo = (pr[2] * (best_v - err_px * dy) - pr[1]) / (Kt[1] - Kt[2] * (best_v - err_px * dy))
            hi = (pr[2] * (best_v + err_px * dy) - pr[1]) / (Kt[1] - Kt[2] * (best_v + err_px * dy))
    pt.idmin, pt.idmax = float(lo), float(hi)
    if pt.idmin > pt.idmax:
        pt.idmin, pt.idmax = pt.idmax, pt.idmin
    pt.interval = 2 * err_px
    pt.uv = (best_u, best_v)
    pt.status = IPS_GOOD
    return IPS_GOOD


def linearize_residual(pt, K, R, t, a_t, b_t, grad_host, grad_target, slack, res, acc, idepth, p=DEFAULTS):
    """One target frame: res = dict(state, energy, new_state, new_energy); acc = [Hdd, bd] (fp32 running sums, updated in place even
    when a later pattern pixel leaves the image -- the reference's early return keeps them).  Returns the energy."""
    if res["state"] == RES_OOB:
        res["new_state"] = RES_OOB
        return res["energy"]
    fx, fy, cx, cy = K
    H, W = grad_target.shape[:2]
    energy = F32(0)
    huber = float(F32(p["huber"])); c = F32(p["outlier_th_sum"])
    ixc, iyc = int(pt.xy[0]), int(pt.xy[1])
    for sx, sy in PATTERN:
        ux = (pt.xy[0] + sx - cx) / fx; uy = (pt.xy[1] + sy - cy) / fy
        q = R @ np.array([ux, uy, 1.0]) + t * float(idepth)
        with np.errstate(all="ignore"):
            px, py = q[0] / q[2], q[1] / q[2]
            proj = (fx * px + cx, fy * py + cy)
            dres = 1.0 / q[2]
        if not (proj[0] >= 1 and proj[1] >= 1 and proj[0] < W - 1 and proj[1] < H - 1) or dres <= 0:
            res["new_state"] = RES_OOB
            return res["energy"]
        gv = interp(grad_target, proj[0], proj[1])
        gt = grad_host[iyc + sy, ixc + sx]
        r = float(gv[0]) - (a_t * float(gt[0]) + b_t)
        hw = 1.0 if abs(r) < huber else huber / abs(r)
        w = float(np.sqrt(F32(c / F32(c + F32(F32(gt[1] * gt[1]) + F32(gt[2] * gt[2]))))))
        energy = F32(float(energy) + w * w * hw * r * r * (2 - hw))
        dxi = float(gv[1]) * fx; dyi = float(gv[2]) * fy
        d_id = dxi * dres * (t[0] - t[2] * px) + dyi * dres * (t[1] - t[2] * py)
        hw *= w * w
        acc[0] = F32(float(acc[0]) + (hw * d_id) * d_id)
        acc[1] = F32(float(acc[1]) + (hw * r) * d_id)
    if float(energy) > pt.energyTH * float(F32(slack)):
        energy = F32(pt.energyTH * float(F32(slack)))
        res["new_state"] = RES_OUTLIER
    else:
        res["new_state"] = RES_IN
    res["new_energy"] = float(energy)
    return float(energy)


def optimize_immature_point(pt, K, cams, exposures, grads, window, min_obs=1, p=DEFAULTS):
    """DSOTracer::optimizeImmaturePoint over the frames `window` (indices; the host is skipped).  Returns (rc, idepth, states)."""
    targets = [f for f in window if f != pt.host]
    res = [dict(state=RES_IN, energy=0.0, new_state=RES_OUTLIER, new_energy=0.0) for _ in targets]
    pre = []
    for f in targets:
        R, t = rel_pose(cams[pt.host], cams[f])
        a_t, b_t = exposure_to(exposures[pt.host], exposures[f])
        pre.append((R, t, a_t, b_t))
    last_e = F32(0); acc = [F32(0), F32(0)]
    cur = F32((pt.idmax + pt.idmin) * float(F32(0.5)))
    for i, f in enumerate(targets):
        last_e = F32(float(last_e) + linearize_residual(pt, K, *pre[i], grads[pt.host], grads[f], 1000, res[i], acc, cur, p))
        res[i]["state"] = res[i]["new_state"]; res[i]["energy"] = res[i]["new_energy"]
    last_h, last_b = acc
    if not math.isfinite(float(last_e)) or float(last_h) < float(F32(p["min_idepth_h_act"])):
        return 0, 0.0, [r["state"] for r in res]
    lam = F32(0.1)
    for _ in range(p["gn_iterations"]):
        Hh = F32(last_h * F32(F32(1) + lam))
        step = F32((1.0 / float(Hh)) * float(last_b))
        new_id = F32(cur - step)
        nacc = [F32(0), F32(0)]; new_e = F32(0)
        for i, f in enumerate(targets):
            new_e = F32(float(new_e) + linearize_residual(pt, K, *pre[i], grads[pt.host], grads[f], 1, res[i], nacc, new_id, p))
        if not math.isfinite(float(last_e)) or float(nacc[0]) < float(F32(p["min_idepth_h_act"])):
            return 0, 0.0, [r["state"] for r in res]
        if float(new_e) < float(last_e):
            cur = new_id; last_h, last_b = nacc; last_e = new_e
            for r in res:
                r["state"] = r["new_state"]; r["energy"] = r["new_energy"]
            lam = F32(lam * F32(0.5))
        else:
            lam = F32(lam * F32(5))
        if abs(float(step)) < 0.0001 * float(cur):
            break
    if not math.isfinite(float(cur)) or float(cur) <= 0:
        return -1, 0.0, [r["state"] for r in res]
    good = sum(1 for r in res if r["state"] == RES_IN)
    if good < min_obs or not math.isfinite(pt.energyTH):
        return -1, 0.0, [r["state"] for r in res]
    return 1, float(cur), [r["state"] for r in res]


def adapt_minimum_distance(cur, num_active, desired):
    """Head of activatePoints (DSOTracer.cpp:62-85): returns (mCurrentMinimumDistance, mUrgentlyNeedNewPoints)."""
    n = num_active
    if n < desired * 0.66: cur -= 0.8
    if n < desired * 0.8: cur -= 0.5
    elif n < desired * 0.9: cur -= 0.2
    elif n < desired: cur -= 0.1
    if n > desired * 1.5: cur += 0.8
    if n > desired * 1.3: cur += 0.5
    if n > desired * 1.15: cur += 0.2
    if n > desired: cur += 0.1
    urgent = cur < 1
    return min(max(cur, 0.0), 4.0), urgent


def activate_points(pts, order, types, active_xy, K, cams, exposures, grads, window, last_frame, min_distance, desired_density, min_quality=3.0, min_obs=1, p=DEFAULTS):
    """DSOTracer::activatePoints (DSOTracer.cpp:62-278) for the immature points `pts` (dict id -> ImmaturePoint) visited in `order`; `active_xy` = the active
    points projected into the last frame.  The distance map (utils/DistanceMap.h: 8-neighbour BFS capped at maxDist) is the Chebyshev distance to the
    nearest added integer pixel.  Returns (mapped {id: idepth}, removed [ids], min_distance, urgent)."""
    H, W = grads[last_frame].shape[:2]
    fx, fy, cx, cy = K
    min_distance, urgent = adapt_minimum_distance(min_distance, len(active_xy), desired_density)
    max_dist = int(min_distance * float(F32(10)))
    added = []

    def add(x, y):
        x, y = int(x), int(y)
        if 0 <= x < W and 0 <= y < H:
            added.append((x, y))

    def get(x, y):
        x, y = int(x), int(y)
        if not added:
            return max_dist
        a = np.asarray(added)
        return min(max_dist, int(np.maximum(np.abs(a[:, 0] - x), np.abs(a[:, 1] - y)).min()))

    for q in active_xy:
        add(q[0], q[1])
    to_opt, removed = [], []
    for i in order:
        pt = pts[i]
        if pt.host == last_frame:
            continue
        if not math.isfinite(pt.idmax) or pt.status == IPS_OUTLIER:
            removed.append(i); continue
        can = pt.status in (IPS_GOOD, IPS_SKIPPED, IPS_BADCONDITION, IPS_OOB) and pt.interval < 8 and pt.quality > float(F32(min_quality)) and (pt.idmax + pt.idmin) > 0
        if not can:
            if pt.status == IPS_OOB:
                removed.append(i)
            continue
        idepth = (pt.idmin + pt.idmax) / 2.0
        R, t = rel_pose(cams[pt.host], cams[last_frame])
        X = R @ (np.array([(pt.xy[0] - cx) / fx, (pt.xy[1] - cy) / fy, 1.0]) / idepth) + t
        px, py = fx * X[0] / X[2] + cx, fy * X[1] / X[2] + cy
        if px >= 0 and py >= 0 and px < W and py < H:
            dist = get(px, py) + (px - math.floor(px))
            if dist >= min_distance * float(F32(types[i])):
                add(px, py); to_opt.append(i)
        else:
            removed.append(i)
    mapped = {}
    for i in to_opt:
        rc, idp, _ = optimize_immature_point(pts[i], K, cams, exposures, grads, window, min_obs, p)
        if rc == 1:
            mapped[i] = idp
        elif rc == -1 or pts[i].status == IPS_OOB:
            removed.append(i)
    return mapped, removed, min_distance, urgent
