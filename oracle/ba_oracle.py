"""CPU restatement (numpy) of libCML's DSO photometric bundle adjustment -- TEST INFRASTRUCTURE ONLY.

This file is the *oracle* for the hot path named in BASELINE.json: it restates
`CML::Optimization::DSOBundleAdjustment::run` (reference file
src/cml/optimization/dso/DSOBundleAdjustment.cpp, "BA" below) and its callees in plain numpy.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it; the product
(libcml_b200/, include/) never does.

Parity is PINNED: tests/test_oracle_golden.py checks every stage of this restatement against golden
vectors produced by the unmodified reference compiled from /root/reference (oracle/Makefile ->
oracle/_ref/cmlba_ref, driver oracle/ref_driver.cpp) and committed under tests/golden/.

Reference anchors (file:line under /root/reference/src/cml):
  linearize                optimization/dso/DSOBundleAdjustment.cpp:62-316
  applyRes                 ...:2051-2093
  computeAdjoints/Delta    ...:1030-1194
  addToHessianTop          ...:1648-1779   + MatrixAccumulators.h:776-937 (AccumulatorApprox)
  stitchDoubleTop          ...:1781-1878
  addToHessianSC           ...:1880-1937
  stitchDoubleSC           ...:1939-2043
  solveLevenbergMarquardt  ...:1284-1337 ; orthogonalize ...:1196-1261 ; computeNullspaces ...:2365-2417
  solveSystem tail         ...:1427-1487
  doStepFromBackup         ...:948-1028
  setNewFrameEnergyTH      ...:2419-2464
  linearizeAll(true) tail  ...:1568-1642
  DSOFrame state algebra   optimization/dso/DSOFrame.h:88-199, DSOFramePrecomputed :259-273
  bilinear interpolate     image/Array2D.h:265-286 ; Exposure::to map/Exposure.h:119-123
  SE3 exp/log/Adj          thirdparty/Sophus/sophus/se3.hpp, so3.hpp (published closed forms)
"""
import numpy as np

F32 = np.float32
STAR8 = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.float64)  # types.h:1395-1407
IN, OOB, OUTLIER = 0, 1, 2  # DSOResidual.h:14-16

DEFAULTS = dict(  # DSOBundleAdjustment.h:235-288
    huber=9.0, outlier_th_sum=2500.0, scale_rot=1.0, scale_trans=0.5, scale_a=10.0, scale_b=1000.0,
    scale_f=50.0, scale_c=50.0, fixed_lambda=1e-5, fix_lambda=True, force_accept=True,
    idepth_fix_prior=2500, solver_mode_delta=1e-5, th_opt_iterations=1.2, optimize_a=True, optimize_b=True,
    disable_marginalization=True,
)


# ----------------------------------------------------------------------------- SE3 (Sophus conventions)
def hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)


def so3_exp(w):
    th2 = float(w @ w)
    th = np.sqrt(th2)
    W = hat(w)
    if th < 1e-10:
        return np.eye(3) + W + 0.5 * W @ W, th
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th2 * (W @ W), th


def se3_exp(xi):
    """xi = (upsilon, omega); returns (R, t) with t = V(omega) * upsilon (sophus/se3.hpp exp)."""
    ups, w = np.asarray(xi[:3], float), np.asarray(xi[3:6], float)
    R, th = so3_exp(w)
    W = hat(w)
    if th < 1e-10:
        V = np.eye(3) + 0.5 * W + (1.0 / 6.0) * (W @ W)
    else:
        V = np.eye(3) + (1 - np.cos(th)) / (th * th) * W + (th - np.sin(th)) / (th ** 3) * (W @ W)
    return R, V @ ups


def so3_log(R):
    # via quaternion, like Sophus (robust near zero)
    tr = np.trace(R)
    qw = np.sqrt(max(0.0, 1 + tr)) / 2
    if qw > 1e-6:
        qv = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (4 * qw)
    else:  # 180 deg, not reached in sliding-window BA
        i = int(np.argmax(np.diag(R)))
        v = np.zeros(3); v[i] = np.sqrt(max(0.0, (R[i, i] + 1) / 2))
        for j in range(3):
            if j != i: v[j] = (R[i, j] + R[j, i]) / (4 * v[i])
        qv = v
    n2 = float(qv @ qv)
    n = np.sqrt(n2)
    if n < 1e-10:
        two_atan = 2.0 / qw - (2.0 / 3.0) * n2 / (qw ** 3)
    else:
        two_atan = 2 * np.arctan2(n, qw) / n if abs(qw) > 1e-10 else (np.pi / n if qw >= 0 else -np.pi / n)
    return two_atan * qv


def se3_log(R, t):
    w = so3_log(R)
    th = np.linalg.norm(w)
    W = hat(w)
    if th < 1e-10:
        Vinv = np.eye(3) - 0.5 * W + (1.0 / 12.0) * (W @ W)
    else:
        half = 0.5 * th
        Vinv = np.eye(3) - 0.5 * W + (1 - th * np.cos(half) / (2 * np.sin(half))) / (th * th) * (W @ W)
    return np.concatenate([Vinv @ t, w])


def se3_mul(A, B):
    return A[0] @ B[0], A[0] @ B[1] + A[1]


def se3_inv(A):
    return A[0].T, -A[0].T @ A[1]


def se3_adj(A):
    R, t = A
    M = np.zeros((6, 6))
    M[:3, :3] = R; M[:3, 3:] = hat(t) @ R; M[3:, 3:] = R
    return M


# ----------------------------------------------------------------------------- window state
class Window:
    """Mirror of the BA's bookkeeping (DSOContext/DSOFrame/DSOPoint/DSOResidual) as flat arrays."""

    def __init__(self, win, grad=None, **params):
        p = dict(DEFAULTS); p.update(params)
        for k in ("optimize_a", "optimize_b", "force_accept"):
            if k in win: p[k] = bool(np.asarray(win[k]).ravel()[0])
        if "fixed_lambda" in win: p["fixed_lambda"] = float(F32(np.asarray(win["fixed_lambda"]).ravel()[0]))
        self.p = p
        self.W, self.H = int(win["size"][0]), int(win["size"][1])
        self.fx, self.fy, self.cx, self.cy = [float(v) for v in win["calib"]]
        self.iterations = int(np.asarray(win.get("iterations", [4])).ravel()[0])
        self.update_points_only = bool(np.asarray(win.get("update_points_only", [0])).ravel()[0])
        self.grad = np.ascontiguousarray(grad if grad is not None else win["grad"], dtype=F32)  # [N,H,W,3]
        N = self.N = win["frame_evalpt"].shape[0]
        sc = self.scales = np.array([p["scale_trans"]] * 3 + [p["scale_rot"]] * 3 + [p["scale_a"], p["scale_b"], p["scale_a"], p["scale_b"]], dtype=np.float64)
        # --- addNewFrame (BA:417-462) -> setEvalPT_scaled (DSOFrame.h:98-106)
        self.evalpt = [(win["frame_evalpt"][i, :9].reshape(3, 3).copy(), win["frame_evalpt"][i, 9:].copy()) for i in range(N)]
        self.exposure = np.asarray(win["frame_exposure"], dtype=np.float64).copy()
        self.state_scaled = np.zeros((N, 10)); self.state_scaled[:, 6:8] = win["frame_affine"]
        self.state = self.state_scaled / sc
        self.state_zero = self.state.copy()
        self.prior_zero = np.zeros((N, 10))
        self.keyid = np.arange(N)
        self.frame_energy_th = np.full(N, 512.0)  # DSOFrame.h:35
        self.step = np.zeros((N, 10))
        self.pre_w2c = [None] * N
        for i in range(N): self._set_state(i, self.state[i])
        self.ns_pose = [None] * N; self.ns_scale = [None] * N; self.ns_affine = [None] * N
        for i in range(N): self._set_state_zero(i, self.state[i])
        # --- run() prologue: updateCamera -> setStateFromCamera (DSOFrame.h:143-151)
        for i in range(N):
            cam = (win["frame_cam"][i, :9].reshape(3, 3), win["frame_cam"][i, 9:])
            r2c = se3_mul(cam, se3_inv(self.evalpt[i]))
            s = self.state_scaled[i].copy(); s[:6] = se3_log(*r2c)
            self._set_state_scaled(i, s)
        # --- points: addPoints (BA:382-415), DSOContext::addPoint (DSOContext.h:76-91)
        self.pt_host = np.asarray(win["pt_host"], dtype=np.int64)
        self.pt_xy = np.asarray(win["pt_xy"], dtype=F32)
        self.idepth = np.asarray(win["pt_idepth"], dtype=np.float64).copy()
        P = self.P = self.pt_host.size
        self.idepth_zero = self.idepth.astype(F32)
        self.has_prior = np.asarray(win["frame_init"], dtype=bool)[self.pt_host] if "frame_init" in win else np.zeros(P, bool)
        ix = self.pt_xy[:, 0].astype(np.int64); iy = self.pt_xy[:, 1].astype(np.int64)
        self.colors = np.zeros((P, 8)); self.weights = np.zeros((P, 8))
        c = p["outlier_th_sum"]
        for k in range(8):
            sx, sy = int(STAR8[k, 0]), int(STAR8[k, 1])
            self.colors[:, k] = self.grad[self.pt_host, iy + sy, ix + sx, 0]  # getGrayPatch: integer pixel (MapObject.h:398-399)
            g = bilinear(self.grad, self.pt_host, self.pt_xy[:, 0] + F32(sx), self.pt_xy[:, 1] + F32(sy))
            g2 = g[:, 1].astype(np.float64) ** 2 + g[:, 2].astype(np.float64) ** 2
            self.weights[:, k] = np.sqrt(F32(c) / (F32(c) + g2))
        # --- residuals: one per (point, frame != host) (createResidual BA:336-380); explicit list allowed
        if "res_point" in win:
            self.res_point = np.asarray(win["res_point"], dtype=np.int64); self.res_target = np.asarray(win["res_target"], dtype=np.int64)
        else:
            pp, tt = np.meshgrid(np.arange(P), np.arange(N), indexing="ij")
            m = tt != self.pt_host[:, None]
            self.res_point, self.res_target = pp[m], tt[m]
        R = self.R = self.res_point.size
        self.res_host = self.pt_host[self.res_point]
        # resetOOB (DSOResidual.h:81-86)
        self.res_state = np.full(R, IN); self.res_new_state = np.full(R, OUTLIER)
        self.res_energy = np.zeros(R); self.res_new_energy = np.zeros(R); self.res_new_energy_wo = np.zeros(R)
        self.res_good = np.zeros(R, bool)
        self.res_alive = np.ones(R, bool)
        self.JpJdF = np.zeros((R, 8), dtype=F32)
        self.efsJ = None; self.rJ = None
        self.num_good_res = np.zeros(P, dtype=np.int64); self.max_rel_baseline = np.zeros(P, dtype=F32)
        self.idepth_hessian = np.zeros(P, dtype=F32)
        self.pt_step = np.zeros(P)
        n = 8 * N + 4
        self.HM = np.zeros((n, n)); self.bM = np.zeros(n)

    # DSOFrame::setState / setStateScaled (DSOFrame.h:110-141)
    def _set_state(self, i, state):
        self.state[i] = state
        self.state_scaled[i] = self.scales * state
        self.pre_w2c[i] = se3_mul(se3_exp(self.state_scaled[i, :6]), self.evalpt[i])

    def _set_state_scaled(self, i, ss):
        self.state_scaled[i] = ss
        self.state[i] = ss / self.scales
        self.pre_w2c[i] = se3_mul(se3_exp(ss[:6]), self.evalpt[i])

    # DSOFrame::setStateZero (DSOFrame.h:153-186)
    def _set_state_zero(self, i, state_zero):
        self.state_zero[i] = state_zero
        E = self.evalpt[i]; Ei = se3_inv(E)
        nsp = np.zeros((6, 6))
        for k in range(6):
            eps = np.zeros(6); eps[k] = 1e-3
            Pp = se3_mul(se3_mul(E, se3_exp(eps)), Ei); Pm = se3_mul(se3_mul(E, se3_exp(-eps)), Ei)
            nsp[:, k] = (se3_log(*Pp) - se3_log(*Pm)) / 2e-3
        self.ns_pose[i] = nsp
        Pp = se3_mul((E[0], E[1] * 1.00001), Ei); Pm = se3_mul((E[0], E[1] / 1.00001), Ei)
        self.ns_scale[i] = (se3_log(*Pp) - se3_log(*Pm)) / 2e-3
        a0 = self.state_zero[i, 6] * self.p["scale_a"]
        na = np.zeros((4, 2)); na[0, 0] = 1; na[1, 0] = 0; na[0, 1] = 0; na[1, 1] = float(np.exp(F32(a0))) * self.exposure[i]
        self.ns_affine[i] = na

    def aff(self, i):       # aff_g2l (DSOFrame.h:189-191)
        return self.state_scaled[i, 6], self.state_scaled[i, 7]

    def aff0(self, i):      # aff_g2l_0 (DSOFrame.h:193-195)
        return self.state_zero[i, 6] * self.p["scale_a"], self.state_zero[i, 7] * self.p["scale_b"]


def exposure_to(a_h, b_h, tau_h, a_t, b_t, tau_t):  # map/Exposure.h:119-123
    a = np.exp(a_t - a_h) * tau_t / tau_h
    return a, b_t - a * b_h


def bilinear(grad, frame, x, y):
    """image/Array2D.h:265-286 in fp32; grad [N,H,W,3]; frame int array; x,y float32 arrays."""
    x = x.astype(F32); y = y.astype(F32)
    ix = x.astype(np.int64); iy = y.astype(np.int64)
    dx = (x - ix.astype(F32)).astype(F32); dy = (y - iy.astype(F32)).astype(F32)
    dxdy = (dx * dy).astype(F32)
    w00 = (F32(1) - dx - dy + dxdy)[:, None]; w10 = (dx - dxdy)[:, None]; w01 = (dy - dxdy)[:, None]; w11 = dxdy[:, None]
    return (grad[frame, iy, ix] * w00 + grad[frame, iy, ix + 1] * w10 + grad[frame, iy + 1, ix] * w01 + grad[frame, iy + 1, ix + 1] * w11).astype(F32)


# ----------------------------------------------------------------------------- per-run precomputation
def compute_adjoints(w):
    """BA:1030-1101. Returns AH, AT indexed [h + N*t]."""
    N = w.N
    AH = np.zeros((N * N, 8, 8)); AT = np.zeros((N * N, 8, 8))
    sc = w.scales[:8]
    for h in range(N):
        for t in range(N):
            T0 = se3_mul(w.evalpt[t], se3_inv(w.evalpt[h]))
            a0, _ = exposure_to(*w.aff0(h), w.exposure[h], *w.aff0(t), w.exposure[t])
            ah = np.eye(8); at = np.eye(8)
            ah[:6, :6] = -se3_adj(T0).T
            at[6, 6] = -a0; ah[6, 6] = a0; at[7, 7] = -1; ah[7, 7] = a0
            AH[h + N * t] = ah * sc[:, None]; AT[h + N * t] = at * sc[:, None]
    w.AH, w.AT = AH, AT


def compute_delta(w):
    """BA:1103-1194."""
    N = w.N
    dz = (w.state - w.state_zero)[:, :8]
    w.ad_ht_delta = np.zeros((N * N, 8))
    for h in range(N):
        for t in range(N):
            i = h + N * t
            w.ad_ht_delta[i] = dz[h] @ w.AH[i] + dz[t] @ w.AT[i]
    prior = np.zeros((N, 8))
    pa = 1e12 if w.p["optimize_a"] else 1e14
    pb = 1e8 if w.p["optimize_b"] else 1e14
    for i in range(N):
        if w.keyid[i] == 0:
            prior[i] = [1e10] * 3 + [1e11] * 3 + [1e14, 1e14]
        else:
            prior[i, 6] = pa; prior[i, 7] = pb
    w.prior = prior.astype(F32).astype(np.float64)  # the settings are float in the reference (BA:1129-1135)
    w.delta = dz.copy()
    w.delta_prior = (w.state - w.prior_zero)[:, :8]
    w.priorF = np.where(w.has_prior, F32(w.p["idepth_fix_prior"]), F32(0)).astype(F32)
    w.deltaF = (w.idepth - w.idepth_zero.astype(np.float64)).astype(F32)


def precompute_pairs(w):
    """DSOFramePrecomputed::precompute (DSOFrame.h:259-273), for all (h,t)."""
    N = w.N
    pc = dict(R=np.zeros((N, N, 3, 3)), t=np.zeros((N, N, 3)), R0=np.zeros((N, N, 3, 3)), t0=np.zeros((N, N, 3)), a=np.zeros((N, N)), b=np.zeros((N, N)))
    for h in range(N):
        for t in range(N):
            T = se3_mul(w.pre_w2c[t], se3_inv(w.pre_w2c[h])); T0 = se3_mul(w.evalpt[t], se3_inv(w.evalpt[h]))
            pc["R"][h, t], pc["t"][h, t] = T; pc["R0"][h, t], pc["t0"][h, t] = T0
            pc["a"][h, t], pc["b"][h, t] = exposure_to(*w.aff(h), w.exposure[h], *w.aff(t), w.exposure[t])
    return pc


# ----------------------------------------------------------------------------- HOT LOOP 1: linearize
def linearize_all(w, fix_linearization=False):
    """BA:1497-1646 + LinearizationContext::linearize BA:62-316. Returns summed energy."""
    pc = precompute_pairs(w)
    act = np.nonzero(w.res_alive)[0]
    h = w.res_host[act]; t = w.res_target[act]; p = w.res_point[act]
    n = act.size
    fx, fy, cx, cy = w.fx, w.fy, w.cx, w.cy
    fxf, fyf = F32(fx), F32(fy)
    finvx, finvy = 1.0 / fx, 1.0 / fy
    Rm, tv, R0, t0 = pc["R"][h, t], pc["t"][h, t], pc["R0"][h, t], pc["t0"][h, t]
    rho = w.idepth[p]
    x = w.pt_xy[p, 0].astype(np.float64); y = w.pt_xy[p, 1].astype(np.float64)

    w.res_new_energy_wo[act] = -1
    was_oob = w.res_state[act] == OOB
    ret = w.res_energy[act].copy()          # value returned on every early exit (BA:71,117,211,222,299)
    new_state = w.res_new_state[act].copy()
    new_energy = w.res_new_energy[act].copy()
    new_energy_wo = np.full(n, -1.0)

    def project(xx, yy):
        k = np.stack([(xx - cx) * finvx, (yy - cy) * finvy, np.ones_like(xx)], axis=1)
        Pp = np.einsum("rij,rj->ri", Rm, k) + tv * rho[:, None]
        return k, Pp, Pp[:, 0] / Pp[:, 2] * fx + cx, Pp[:, 1] / Pp[:, 2] * fy + cy

    def inside(Ku, Kv):
        with np.errstate(invalid="ignore"):
            return (Ku >= 2) & (Kv >= 2) & (Ku < F32(w.W) - 2) & (Kv < F32(w.H) - 2)

    KliP, Pc, Ku, Kv = project(x, y)
    ok = ~was_oob
    oob = ok & ~inside(Ku, Kv)
    ok &= ~oob
    drescale = (1.0 / Pc[:, 2]).astype(F32)
    new_idepth = (drescale.astype(np.float64) * rho).astype(F32)
    u = Pc[:, 0].astype(F32); v = Pc[:, 1].astype(F32)   # sic: un-normalised (BA:121-122)
    center = np.stack([Ku.astype(F32).astype(np.float64), Kv.astype(F32).astype(np.float64), new_idepth.astype(np.float64)], axis=1)
    ud, vd, dr = u.astype(np.float64), v.astype(np.float64), drescale.astype(np.float64)
    Jpdd = np.stack([dr * (t0[:, 0] - t0[:, 2] * ud) * fx_(fxf), dr * (t0[:, 1] - t0[:, 2] * vd) * fx_(fyf)], axis=1).astype(F32)
    sF, sC = float(F32(w.p["scale_f"])), float(F32(w.p["scale_c"]))
    fxdr = (fxf * drescale).astype(np.float64); fydr = (fyf * drescale).astype(np.float64)
    dCx = np.zeros((n, 4)); dCy = np.zeros((n, 4))
    dCx[:, 2] = dr * (R0[:, 2, 0] * ud - R0[:, 0, 0]); dCx[:, 3] = fxdr * (R0[:, 2, 1] * ud - R0[:, 0, 1]) / fx_(fyf)
    dCx[:, 0] = KliP[:, 0] * dCx[:, 2]; dCx[:, 1] = KliP[:, 1] * dCx[:, 3]
    dCy[:, 2] = fydr * (R0[:, 2, 0] * vd - R0[:, 1, 0]) / fx_(fxf); dCy[:, 3] = dr * (R0[:, 2, 1] * vd - R0[:, 1, 1])
    dCy[:, 0] = KliP[:, 0] * dCy[:, 2]; dCy[:, 1] = KliP[:, 1] * dCy[:, 3]
    dCx[:, 0] = (dCx[:, 0] + ud) * sF; dCx[:, 1] *= sF; dCx[:, 2] = (dCx[:, 2] + 1) * sC; dCx[:, 3] *= sC
    dCy[:, 0] *= sF; dCy[:, 1] = (dCy[:, 1] + vd) * sF; dCy[:, 2] *= sC; dCy[:, 3] = (dCy[:, 3] + 1) * sC
    one = F32(1)
    dxi_x = np.stack([new_idepth * fxf, np.zeros(n, F32), -new_idepth * u * fxf, -u * v * fxf, (one + u * u) * fxf, -v * fxf], axis=1).astype(F32)
    dxi_y = np.stack([np.zeros(n, F32), new_idepth * fyf, -new_idepth * v * fyf, -(one + v * v) * fyf, u * v * fyf, u * fyf], axis=1).astype(F32)
    rJ = dict(Jpdxi=np.stack([dxi_x, dxi_y], axis=1), Jpdc=np.stack([dCx, dCy], axis=1).astype(F32), Jpdd=Jpdd,
              resF=np.zeros((n, 8), F32), JIdx=np.zeros((n, 2, 8), F32), JabF=np.zeros((n, 2, 8), F32))
    a_ht, b_ht = pc["a"][h, t], pc["b"][h, t]
    b0 = (w.state_zero[h, 7] * float(F32(w.p["scale_b"]))).astype(F32)
    huber = F32(w.p["huber"]); cth = F32(w.p["outlier_th_sum"])
    J00 = np.zeros(n, F32); J11 = np.zeros(n, F32); J10 = np.zeros(n, F32)
    A00 = np.zeros(n, F32); A01 = np.zeros(n, F32); A10 = np.zeros(n, F32); A11 = np.zeros(n, F32)
    B00 = np.zeros(n, F32); B01 = np.zeros(n, F32); B11 = np.zeros(n, F32)
    wJI2 = np.zeros(n, F32); E = np.zeros(n, F32)
    nonfinite = np.zeros(n, bool)
    for k in range(8):
        _, _, qx, qy = project(x + STAR8[k, 0], y + STAR8[k, 1])
        o = ok & ~inside(qx, qy)
        oob |= o; ok &= ~o
        qxs = np.where(ok, qx, 2.0).astype(F32); qys = np.where(ok, qy, 2.0).astype(F32)
        hit = bilinear(w.grad, t, qxs, qys)
        nf = ok & ~np.isfinite(hit).all(axis=1)
        nonfinite |= nf; ok &= ~nf
        I, gx, gy = hit[:, 0], hit[:, 1], hit[:, 2]
        ref_real = (a_ht * w.colors[p, k] + b_ht).astype(F32)
        r = (I - ref_real).astype(F32)
        ar = np.abs(r)
        with np.errstate(divide="ignore", invalid="ignore"):
            hw = np.where(ar < huber, one, huber / ar).astype(F32)
        ww = np.sqrt(cth / (cth + (gx * gx + gy * gy))).astype(F32)
        ww = (0.5 * (ww.astype(np.float64) + w.weights[p, k])).astype(F32)
        E = (E.astype(np.float64) + (ww * ww * hw * r * r).astype(np.float64) * (2.0 - hw.astype(np.float64))).astype(F32)
        hw = np.where(hw < 1, np.sqrt(hw), hw).astype(F32) * ww
        h1 = (gx * hw).astype(F32); h2 = (gy * hw).astype(F32)
        drdA = (I - b0).astype(F32)
        rJ["resF"][:, k] = r * hw
        rJ["JIdx"][:, 0, k] = h1; rJ["JIdx"][:, 1, k] = h2
        rJ["JabF"][:, 0, k] = drdA * hw if w.p["optimize_a"] else 0
        rJ["JabF"][:, 1, k] = hw if w.p["optimize_b"] else 0
        J00 += h1 * h1; J11 += h2 * h2; J10 += h1 * h2
        A00 += drdA * hw * h1; A01 += drdA * hw * h2; A10 += hw * h1; A11 += hw * h2
        B00 += drdA * drdA * hw * hw; B01 += drdA * hw * hw; B11 += hw * hw
        wJI2 += hw * hw * (h1 * h1 + h2 * h2)
    rJ["JIdx2"] = np.stack([J00, J10, J10, J11], axis=1).reshape(n, 2, 2)
    rJ["JabJIdx"] = np.stack([A00, A01, A10, A11], axis=1).reshape(n, 2, 2)
    rJ["Jab2"] = np.stack([B00, B01, B01, B11], axis=1).reshape(n, 2, 2)
    o = ok & ~np.isfinite(E)
    oob |= o; ok &= ~o
    th = np.maximum(w.frame_energy_th[h].astype(F32), w.frame_energy_th[t].astype(F32))
    Ed = E.astype(np.float64)
    outl = ok & ((E > th) | (wJI2 < 2))
    new_energy_wo = np.where(ok, Ed, new_energy_wo)
    fin = np.where(outl, th.astype(np.float64), Ed)
    new_state = np.where(oob, OOB, new_state)
    new_state = np.where(ok, np.where(outl, OUTLIER, IN), new_state)
    new_energy = np.where(ok, fin, new_energy)
    ret = np.where(ok, fin, ret)
    # quirk BA:220-223: non-finite sample sets the *committed* state to OOB
    st = w.res_state[act].copy(); st[nonfinite] = OOB
    w.res_state[act] = st
    w.res_new_state[act] = new_state; w.res_new_energy[act] = new_energy; w.res_new_energy_wo[act] = new_energy_wo
    center_valid = ~was_oob & inside(Ku, Kv)
    if not hasattr(w, "res_center"): w.res_center = np.zeros((w.R, 3))
    cc = w.res_center[act]; cc[center_valid] = center[center_valid]; w.res_center[act] = cc
    w.rJ = dict(idx=act, **rJ)
    w.lin_pc = pc
    energy = float(ret.sum())
    set_new_frame_energy_th(w, act)
    if fix_linearization:
        _fix_linearization_tail(w, act, pc)
    return energy


def fx_(f32v):
    return float(f32v)


def set_new_frame_energy_th(w, act):
    """BA:2419-2464: 0.7-quantile (nth_element) of new-frame energies."""
    newest = w.N - 1
    m = (w.res_new_energy_wo[act] >= 0) & (w.res_target[act] == newest)
    vals = w.res_new_energy_wo[act][m].astype(F32)
    if vals.size == 0:
        w.frame_energy_th[newest] = 12 * 12 * 8
        return
    nth = int(F32(0.7) * F32(vals.size))
    e = np.partition(vals, nth)[nth]
    th = F32(np.sqrt(e)) * F32(1.5)
    th = F32(26.0) * F32(0.5) + th * F32(0.5)
    w.frame_energy_th[newest] = float(F32(th * th))


def apply_active_res(w):
    """applyRes(r, copyJacobians=true) for every active residual (BA:2045-2093)."""
    act = w.rJ["idx"]
    st = w.res_state[act]; ns = w.res_new_state[act]
    not_oob = st != OOB
    is_in = not_oob & (ns == IN)
    good = w.res_good[act].copy()
    good[not_oob] = is_in[not_oob]
    w.res_good[act] = good
    if w.efsJ is None:
        w.efsJ = {k: np.zeros((w.R,) + v.shape[1:], v.dtype) for k, v in w.rJ.items() if k != "idx"}
    sel = act[is_in]
    for k in w.efsJ: w.efsJ[k][sel] = w.rJ[k][is_in]
    J = w.efsJ
    v = np.einsum("rab,rb->ra", J["JIdx2"][sel], J["Jpdd"][sel]).astype(F32)
    jp = (J["Jpdxi"][sel, 0] * v[:, :1] + J["Jpdxi"][sel, 1] * v[:, 1:2]).astype(F32)
    ab = np.einsum("rab,rb->ra", J["JabJIdx"][sel], J["Jpdd"][sel]).astype(F32)
    w.JpJdF[sel] = np.concatenate([jp, ab], axis=1)
    st2 = st.copy(); st2[not_oob] = ns[not_oob]
    en = w.res_energy[act].copy(); en[not_oob] = w.res_new_energy[act][not_oob]
    w.res_state[act] = st2; w.res_energy[act] = en


def _fix_linearization_tail(w, act, pc):
    """linearizeAll(true) body BA:1568-1642: applyRes, relBS/numGoodResiduals, drop non-good residuals."""
    apply_active_res(w)
    good = w.res_good[act]
    g = act[good]
    h, t, p = w.res_host[g], w.res_target[g], w.res_point[g]
    K = np.array([[w.fx, 0, w.cx], [0, w.fy, w.cy], [0, 0, 1.0]]); Ki = np.linalg.inv(K)
    KRKi = np.einsum("ij,rjk,kl->ril", K, pc["R"][h, t], Ki); Kt = np.einsum("ij,rj->ri", K, pc["t"][h, t])
    xy1 = np.concatenate([w.pt_xy[p].astype(np.float64), np.ones((g.size, 1))], axis=1)
    inf = np.einsum("rij,rj->ri", KRKi, xy1); real = inf + Kt * w.idepth[p][:, None]
    rel = (0.01 * np.linalg.norm(inf[:, :2] / inf[:, 2:3] - real[:, :2] / real[:, 2:3], axis=1)).astype(F32)
    np.maximum.at(w.max_rel_baseline, p, rel)
    np.add.at(w.num_good_res, p, 1)
    w.res_alive[act[~good]] = False
    cnt = np.bincount(w.res_point[w.res_alive], minlength=w.P)
    w.pt_outlier = (cnt == 0) & getattr(w, "pt_alive", np.ones(w.P, bool))
    w.pt_alive = getattr(w, "pt_alive", np.ones(w.P, bool)) & (cnt > 0)


# ----------------------------------------------------------------------------- HOT LOOP 2/3: accumulate + Schur
def accumulate_top(w):
    """addToHessianTop(ACTIVE) for all points (BA:1648-1779, MatrixAccumulators.h:776-937).
    Returns acc[N*N,13,13] (order [C4|xi6|a|b|r]) and sets per-point Hdd/bd/Hcd."""
    N = w.N
    g = np.nonzero(w.res_good & w.res_alive)[0]
    J = {k: v[g].astype(np.float64) for k, v in w.efsJ.items()}
    res = J["resF"]
    JI_r = np.einsum("rak,rk->ra", J["JIdx"], res); Jab_r = np.einsum("rak,rk->ra", J["JabF"], res); rr = (res * res).sum(1)
    Jp = np.concatenate([J["Jpdc"], J["Jpdxi"]], axis=2)                      # [n,2,10]
    n = g.size
    M = np.zeros((n, 13, 13))
    M[:, :10, :10] = np.einsum("rai,rab,rbj->rij", Jp, J["JIdx2"], Jp)
    TR = np.stack([J["JabJIdx"][:, 0, :], J["JabJIdx"][:, 1, :], JI_r], axis=2)   # [n,2,3]: columns a, b, r
    M[:, :10, 10:] = np.einsum("rai,rac->ric", Jp, TR)
    M[:, 10:, :10] = M[:, :10, 10:].transpose(0, 2, 1)
    M[:, 10, 10] = J["Jab2"][:, 0, 0]; M[:, 10, 11] = M[:, 11, 10] = J["Jab2"][:, 0, 1]; M[:, 11, 11] = J["Jab2"][:, 1, 1]
    M[:, 10, 12] = M[:, 12, 10] = Jab_r[:, 0]; M[:, 11, 12] = M[:, 12, 11] = Jab_r[:, 1]; M[:, 12, 12] = rr
    bins = w.res_host[g] + N * w.res_target[g]
    acc = np.zeros((N * N, 13, 13))
    np.add.at(acc, bins, M)
    w.acc_num = np.bincount(bins, minlength=N * N)
    v = np.einsum("rab,rb->ra", J["JIdx2"], J["Jpdd"])
    p = w.res_point[g]
    w.bd = np.bincount(p, weights=(JI_r * J["Jpdd"]).sum(1), minlength=w.P).astype(F32)
    w.Hdd = np.bincount(p, weights=(v * J["Jpdd"]).sum(1), minlength=w.P).astype(F32)
    Hcd = np.zeros((w.P, 4))
    np.add.at(Hcd, p, J["Jpdc"][:, 0] * v[:, :1] + J["Jpdc"][:, 1] * v[:, 1:2])
    w.Hcd = Hcd.astype(F32)
    w.acc = acc
    return acc


def stitch_top(w, acc, use_prior):
    """stitchDoubleTop BA:1781-1878."""
    N = w.N; n = 8 * N + 4
    H = np.zeros((n, n)); b = np.zeros(n)
    for h in range(N):
        for t in range(N):
            i = h + N * t
            if acc is None or w.acc_num[i] == 0: continue
            A = acc[i]; hI, tI = 4 + 8 * h, 4 + 8 * t
            AH, AT = w.AH[i], w.AT[i]
            H[hI:hI + 8, hI:hI + 8] += AH @ A[4:12, 4:12] @ AH.T
            H[tI:tI + 8, tI:tI + 8] += AT @ A[4:12, 4:12] @ AT.T
            H[hI:hI + 8, tI:tI + 8] += AH @ A[4:12, 4:12] @ AT.T
            H[hI:hI + 8, 0:4] += AH @ A[4:12, 0:4]; H[tI:tI + 8, 0:4] += AT @ A[4:12, 0:4]
            H[0:4, 0:4] += A[0:4, 0:4]
            b[hI:hI + 8] += AH @ A[4:12, 12]; b[tI:tI + 8] += AT @ A[4:12, 12]; b[0:4] += A[0:4, 12]
    if use_prior:
        # calibration prior mCPrior is uninitialised in the reference when forceAccept (BA:2123,2137): rows 0-3 are not compared
        for h in range(N):
            hI = 4 + 8 * h
            H[np.arange(hI, hI + 8), np.arange(hI, hI + 8)] += w.prior[h]
            b[hI:hI + 8] += w.prior[h] * w.delta_prior[h]
    for h in range(N):
        hI = 4 + 8 * h
        H[0:4, hI:hI + 8] = H[hI:hI + 8, 0:4].T
        for t in range(h + 1, N):
            tI = 4 + 8 * t
            H[hI:hI + 8, tI:tI + 8] += H[tI:tI + 8, hI:hI + 8].T
            H[tI:tI + 8, hI:hI + 8] = H[hI:hI + 8, tI:tI + 8].T
    return H, b


def accumulate_sc(w, shift_prior_to_zero=True):
    """addToHessianSC for all points (BA:1880-1937)."""
    N, P = w.N, w.P
    good = w.res_good & w.res_alive
    ngood = np.bincount(w.res_point[good], minlength=P)
    has = ngood > 0
    Hp = (w.Hdd + w.priorF).astype(F32)                  # Hdd_accLF == 0 inside run() (no linearized residuals)
    Hp = np.maximum(Hp, F32(1e-10))
    w.idepth_hessian = np.where(has, Hp, F32(0)).astype(F32)
    w.max_rel_baseline = np.where(has, w.max_rel_baseline, F32(0)).astype(F32)
    w.HdiF = np.where(has, (1.0 / Hp.astype(np.float64)), 0).astype(F32)
    bdS = w.bd.copy()
    if shift_prior_to_zero: bdS = (bdS + w.priorF * w.deltaF).astype(F32)
    w.bdSumF = np.where(has, bdS, F32(0)).astype(F32)
    JT = np.zeros((P, N, 8))
    g = np.nonzero(good)[0]
    JT[w.res_point[g], w.res_target[g]] = w.JpJdF[g]
    Hdi = w.HdiF.astype(np.float64) * has; Hcd = w.Hcd.astype(np.float64); bds = w.bdSumF.astype(np.float64)
    accE = np.zeros((N * N, 8, 4)); accEB = np.zeros((N * N, 8)); accD = np.zeros((N * N * N, 8, 8))
    for h in range(N):
        m = (w.pt_host == h) & has
        if not m.any(): continue
        D = np.einsum("p,pai,pbj->abij", Hdi[m], JT[m], JT[m])
        E = np.einsum("p,pai,pc->aic", Hdi[m], JT[m], Hcd[m])
        EB = np.einsum("p,pai->ai", Hdi[m] * bds[m], JT[m])
        for t1 in range(N):
            accE[h + N * t1] = E[t1]; accEB[h + N * t1] = EB[t1]
            for t2 in range(N):
                accD[h + N * t1 + N * N * t2] = D[t1, t2]
    w.accHcc = np.einsum("p,pi,pj->ij", Hdi, Hcd, Hcd); w.accbc = np.einsum("p,pi->i", Hdi * bds, Hcd)
    w.accE, w.accEB, w.accD = accE, accEB, accD
    w.sc_num_D = np.zeros(N * N * N, dtype=np.int64)
    tgt_good = np.zeros((P, N), bool); tgt_good[w.res_point[g], w.res_target[g]] = True
    for h in range(N):
        m = w.pt_host == h
        c = np.einsum("pa,pb->ab", tgt_good[m].astype(np.int64), tgt_good[m].astype(np.int64))
        for t1 in range(N):
            for t2 in range(N):
                w.sc_num_D[h + N * t1 + N * N * t2] = c[t1, t2]


def stitch_sc(w):
    """stitchDoubleSC BA:1939-2043."""
    N = w.N; n = 8 * N + 4; N2 = N * N
    H = np.zeros((n, n)); b = np.zeros(n)
    for i in range(N):
        for j in range(N):
            iI, jI, ij = 4 + 8 * i, 4 + 8 * j, i + N * j
            H[iI:iI + 8, 0:4] += w.AH[ij] @ w.accE[ij]; H[jI:jI + 8, 0:4] += w.AT[ij] @ w.accE[ij]
            b[iI:iI + 8] += w.AH[ij] @ w.accEB[ij]; b[jI:jI + 8] += w.AT[ij] @ w.accEB[ij]
            for k in range(N):
                kI, ijk, ik = 4 + 8 * k, ij + k * N2, i + N * k
                if w.sc_num_D[ijk] == 0: continue
                D = w.accD[ijk]
                H[iI:iI + 8, iI:iI + 8] += w.AH[ij] @ D @ w.AH[ik].T
                H[jI:jI + 8, kI:kI + 8] += w.AT[ij] @ D @ w.AT[ik].T
                H[jI:jI + 8, iI:iI + 8] += w.AT[ij] @ D @ w.AH[ik].T
                H[iI:iI + 8, kI:kI + 8] += w.AH[ij] @ D @ w.AT[ik].T
    H[0:4, 0:4] = w.accHcc; b[0:4] = w.accbc
    for h in range(N):
        hI = 4 + 8 * h
        H[0:4, hI:hI + 8] = H[hI:hI + 8, 0:4].T
    return H, b


def nullspaces(w):
    """computeNullspaces BA:2365-2417 -> the 7 vectors orthogonalize() uses (6 pose + 1 scale)."""
    N = w.N; n = 8 * N + 4
    ns = np.zeros((7, n))
    st, sr = w.p["scale_trans"], w.p["scale_rot"]
    for i in range(N):
        o = 4 + 8 * i
        for k in range(6):
            ns[k, o:o + 6] = w.ns_pose[i][:, k]
            ns[k, o:o + 3] *= 1.0 / st; ns[k, o + 3:o + 6] *= 1.0 / sr
        ns[6, o:o + 6] = w.ns_scale[i]
        ns[6, o:o + 3] *= 1.0 / st; ns[6, o + 3:o + 6] *= 1.0 / sr
    return ns


def orthogonalize(w, x):
    """BA:1196-1261."""
    ns = nullspaces(w)
    Nm = (ns / np.linalg.norm(ns, axis=1, keepdims=True)).T
    U, S, Vt = np.linalg.svd(Nm, full_matrices=False)
    Si = np.where(S > w.p["solver_mode_delta"] * S.max(), 1.0 / S, 0.0)
    Npi = U @ np.diag(Si) @ Vt
    NNpiT = Nm @ Npi.T
    return x - 0.5 * (NNpiT + NNpiT.T) @ x


def solve_system(w, iteration, lam):
    """solveSystem BA:1339-1495 (+ solveLevenbergMarquardt BA:1284-1337). Sets frame steps and point steps."""
    N = w.N
    if w.p["fix_lambda"]: lam = w.p["fixed_lambda"]
    acc = accumulate_top(w)
    HA, bA = stitch_top(w, acc, False)
    HL, bL = stitch_top(w, None, True)
    accumulate_sc(w, True)
    Hsc, bsc = stitch_sc(w)
    d = np.concatenate([np.zeros(4), w.delta.reshape(-1)])
    if w.p["disable_marginalization"]:
        w.HM[:] = 0; w.bM[:] = 0
    bMt = w.bM + w.HM @ d
    H = HL + w.HM + HA
    b = bL + bMt + bA - bsc
    H[np.diag_indices_from(H)] *= (1 + lam)
    H = H - Hsc * (1.0 / (1 + lam))
    s = 1.0 / np.sqrt(np.diag(H) + 10)
    Hs = H * s[:, None] * s[None, :]
    x = np.zeros(H.shape[0])
    x[4:] = s[4:] * np.linalg.solve(Hs[4:, 4:], s[4:] * b[4:])
    if iteration >= 2: x = orthogonalize(w, x)
    w.sys = dict(HA=HA, bA=bA, HL=HL, bL=bL, Hsc=Hsc, bsc=bsc, bM=bMt, x=x)
    # back-substitution BA:1427-1487
    w.step[:] = 0
    w.step[:, :8] = -x[4:].reshape(N, 8)
    xAd = np.zeros((N * N, 8))
    for h in range(N):
        for t in range(N):
            xAd[N * h + t] = x[4 + 8 * h:12 + 8 * h] @ w.AH[h + N * t] + x[4 + 8 * t:12 + 8 * t] @ w.AT[h + N * t]
    good = np.nonzero(w.res_good & w.res_alive)[0]
    ngood = np.bincount(w.res_point[good], minlength=w.P)
    bb = w.bdSumF.astype(np.float64) - (-x[:4]) @ w.Hcd.astype(np.float64).T
    contrib = (xAd[w.res_host[good] * N + w.res_target[good]] * w.JpJdF[good].astype(np.float64)).sum(1)
    bb -= np.bincount(w.res_point[good], weights=contrib, minlength=w.P)
    w.pt_step = np.where(ngood > 0, -bb * w.HdiF.astype(np.float64), 0.0)
    return bool(np.isfinite(w.pt_step).all())


def backup_state(w):
    w.state_backup = w.state.copy()
    w.idepth_backup = w.idepth.astype(F32)


def do_step_from_backup(w, fix_camera=False):
    """BA:948-1028."""
    N = w.N
    if fix_camera: w.step[:, :6] = 0
    for i in range(N): w._set_state(i, w.state_backup[i] + w.step[i])
    st = w.step.astype(F32)
    sumA = F32((st[:, 6] ** 2).sum()); sumB = F32((st[:, 7] ** 2).sum()); sumT = F32((st[:, :3] ** 2).sum()); sumR = F32((st[:, 3:6] ** 2).sum())
    newid = w.idepth_backup.astype(np.float64) + w.pt_step
    okp = np.isfinite(newid) & (newid > 0) & getattr(w, "pt_alive", np.ones(w.P, bool))
    w.idepth = np.where(okp, newid, w.idepth)
    w.idepth_zero = np.where(okp, w.idepth.astype(F32), w.idepth_zero)
    numID = okp.sum(); sumNID = np.abs(w.idepth_backup[okp]).astype(np.float64).sum()
    sumA /= N; sumB /= N; sumR /= N; sumT /= N
    sumNID /= max(numID, 1)
    compute_delta(w)
    th = w.p["th_opt_iterations"]
    return bool(np.sqrt(sumA) < 0.0005 * th and np.sqrt(sumB) < 0.00005 * th and np.sqrt(sumR) < 0.00005 * th and np.sqrt(sumT) * sumNID < 0.00005 * th)


def calc_l_energy(w):
    """calcLEnergy BA:2118-2208 for a window without linearized residuals: frame priors + point priors; 0 when forceAccept (BA:2123).
    calcMEnergy (BA:2095-2116) is 0 as long as the marginalisation prior H_M, b_M is zero."""
    if w.p["force_accept"]:
        return 0.0
    F = float(np.sum(w.delta_prior * w.prior * w.delta_prior))           # BA:2130-2133 (compute_delta keeps delta_prior current)
    return F + float(np.sum(np.asarray(w.deltaF, np.float64) ** 2 * np.asarray(w.priorF, np.float64)))   # BA:2200


def run(w, hook=None):
    """DSOBundleAdjustment::run BA:744-910 (forceAccept path; the reject branch re-linearizes at the backup)."""
    hook = hook or (lambda *a: None)
    compute_adjoints(w); compute_delta(w)
    hook("pre", w)
    last = linearize_all(w, False)
    hook("lin0", w)
    apply_active_res(w)
    hook("app0", w)
    lam = w.p["fixed_lambda"]
    w.iterations_done = 0
    w.accepted = []
    lastL = calc_l_energy(w)
    for it in range(w.iterations):
        backup_state(w)
        if not solve_system(w, it, lam): return False
        hook(f"sol{it}", w)
        canbreak = do_step_from_backup(w, w.update_points_only)
        w.canbreak = canbreak
        hook(f"step{it}", w)
        new = linearize_all(w, False)
        newL = calc_l_energy(w)
        hook(f"lin{it + 1}", w)
        if not np.isfinite(new + newL): return False
        if new + newL < last + lastL or w.p["force_accept"]:
            apply_active_res(w); last = new; lastL = newL; lam *= 0.25
            w.accepted.append(1)
        else:
            for i in range(w.N): w._set_state(i, w.state_backup[i])
            w.idepth = w.idepth_backup.astype(np.float64); w.idepth_zero = w.idepth_backup.copy()
            compute_delta(w)
            last = linearize_all(w, False); lastL = calc_l_energy(w); lam *= 1e2
            w.accepted.append(0)
        w.iterations_done = it + 1
        if canbreak and it >= 1: break
    # epilogue BA:885-896
    nb = w.N - 1
    w.evalpt[nb] = (w.pre_w2c[nb][0].copy(), w.pre_w2c[nb][1].copy())
    nz = np.zeros(10); nz[6:8] = w.state[nb, 6:8]
    w._set_state(nb, nz); w._set_state_zero(nb, nz)
    compute_adjoints(w); compute_delta(w)
    w.fin_energy = linearize_all(w, True)
    hook("fin", w)
    return bool(np.isfinite(w.fin_energy))


def marginalize_points_prior(w, marg):
    """Contribution of the points `marg` (bool[P]) to the marginalisation prior: tryMarginalize's re-linearization of their
    residuals (resetOOB + linearize + applyRes + fixLinearization, BA:2289-2300, 2210-2238) followed by marginalizePointsF
    (addToHessianTop(MARGINALIZED) + addToHessianSC(false) + stitch, BA:2466-2513).  Returns (dH_M, db_M) = 1/4 (M - M_sc, b - b_sc).
    The window is left with only those residuals alive (the caller removes the points afterwards anyway)."""
    compute_adjoints(w); compute_delta(w)
    th_keep = w.frame_energy_th.copy()
    w.res_alive = w.res_alive & marg[w.res_point]
    act = np.nonzero(w.res_alive)[0]
    w.res_state[act] = IN; w.res_energy[act] = 0                          # resetOOB (DSOResidual.h:81-86)
    linearize_all(w, False)
    w.frame_energy_th = th_keep                                           # the per-residual linearize of tryMarginalize never touches frameEnergyTH
    apply_active_res(w)
    g = np.nonzero(w.res_good & w.res_alive)[0]
    J = w.efsJ
    dp = w.ad_ht_delta[w.res_host[g] + w.N * w.res_target[g]].astype(F32)
    dF = w.deltaF[w.res_point[g]].astype(F32)
    jx = ((J["Jpdxi"][g, 0] * dp[:, :6]).sum(1, dtype=F32) + J["Jpdd"][g, 0] * dF).astype(F32)     # Jpdc . dc = 0 (calibration fixed)
    jy = ((J["Jpdxi"][g, 1] * dp[:, :6]).sum(1, dtype=F32) + J["Jpdd"][g, 1] * dF).astype(F32)
    rtz = (J["resF"][g] - J["JIdx"][g, 0] * jx[:, None] - J["JIdx"][g, 1] * jy[:, None] - J["JabF"][g, 0] * dp[:, 6:7] - J["JabF"][g, 1] * dp[:, 7:8]).astype(F32)
    keep = J["resF"][g].copy()
    J["resF"][g] = rtz                                                    # MARGINALIZED mode accumulates res_toZeroF (BA:1686-1690)
    acc = accumulate_top(w)
    M, Mb = stitch_top(w, acc, False)
    accumulate_sc(w, False)
    Msc, Mbsc = stitch_sc(w)
    J["resF"][g] = keep
    return 0.25 * (M - Msc), 0.25 * (Mb - Mbsc)
