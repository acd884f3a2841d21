"""Deterministic synthetic sliding windows for the photometric-BA hot path.

A tilted textured plane is rendered analytically into every keyframe (exact photoconsistency up to
the affine brightness model I_i = exp(a_i) * tau_i * T + b_i), points are random integer pixels with
an 8-px margin, inverse depths and poses are perturbed around the truth.  The output dict is the
*input* side of a window snapshot (Appendix C): both this repo's CUDA path and the reference driver
(oracle/ref_driver.cpp) are fed from it.  Pure numpy, no oracle dependency.
"""
import numpy as np

STAR8 = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], dtype=np.float64)  # types.h:1395-1407

CONFIGS = {
    # name: (W, H, N, pts_per_kf, iterations, affine)
    "tiny": (160, 120, 3, 60, 3, False),
    "tiny_affine": (160, 120, 4, 50, 4, True),
    "c1": (640, 480, 2, 100, 1, False),       # BASELINE.json configs[0]: 2 KF, 200 active points, 1 GN iteration
    "c2": (640, 480, 8, 2000, 6, False),      # configs[1]: the headline config
    "c3": (1241, 376, 8, 2500, 6, True),      # configs[2]: KITTI shape, affine brightness on
    "c4": (640, 480, 16, 3000, 6, False),     # configs[3]
    "c5": (1920, 1080, 8, 2000, 6, False),    # configs[4] window shape (stream handled by caller)
}


def _rodrigues(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def gradient_image(gray):
    """(I, dI/dx, dI/dy) AoS image as CaptureImageGenerator builds it (image/Array2D.h:288-294, 314-331):
    central differences * 0.5 in fp32, all three channels zero on the 1-px border."""
    g = np.zeros(gray.shape + (3,), dtype=np.float32)
    g[1:-1, 1:-1, 0] = gray[1:-1, 1:-1]
    g[1:-1, 1:-1, 1] = (gray[1:-1, 2:] - gray[1:-1, :-2]) * np.float32(0.5)
    g[1:-1, 1:-1, 2] = (gray[2:, 1:-1] - gray[:-2, 1:-1]) * np.float32(0.5)
    return g


def make_window(W=640, H=480, N=8, pts_per_kf=2000, iterations=6, affine=False, seed=1234,
                idepth_noise=0.005, pose_noise=5e-4, fej_offset=True, with_gradients=True, low_freq=False, with_depth=False):
    rng = np.random.default_rng(seed)
    fx = fy = 0.78 * W
    cx, cy = (W - 1) / 2.0, (H - 1) / 2.0
    # plane n.X = d in world coordinates (mean depth 2 m, slightly tilted)
    n = np.array([0.10, -0.06, 1.0]); n /= np.linalg.norm(n)
    d = 2.0
    # band-limited texture: products of sinusoids, periods >= 12 px at 2 m
    m_per_px = 2.0 / fx
    wmax = 2 * np.pi / (12 * m_per_px)
    nterm = 6
    wx = rng.uniform(0.35 * wmax, wmax, nterm); wy = rng.uniform(0.35 * wmax, wmax, nterm)
    phx = rng.uniform(0, 2 * np.pi, nterm); phy = rng.uniform(0, 2 * np.pi, nterm)
    amp = rng.uniform(12.0, 17.5, nterm)

    if low_freq:
        # coarse-to-fine consumers (the tracker, SURVEY 8f NEXT #1) need structure that survives four halvings: periods 60-400 px
        rl = np.random.default_rng(seed + 7919)
        lwx = 2 * np.pi / (rl.uniform(60, 400, 5) * m_per_px); lwy = 2 * np.pi / (rl.uniform(60, 400, 5) * m_per_px)
        lphx = rl.uniform(0, 2 * np.pi, 5); lphy = rl.uniform(0, 2 * np.pi, 5); lamp = rl.uniform(8.0, 14.0, 5)

    def texture(X, Y):
        T = np.full(X.shape, 127.5)
        for j in range(nterm):
            T = T + amp[j] * np.sin(wx[j] * X + phx[j]) * np.sin(wy[j] * Y + phy[j])
        if low_freq:
            T = 127.5 + 0.45 * (T - 127.5)
            for j in range(5):
                T = T + lamp[j] * np.sin(lwx[j] * X + lphx[j]) * np.sin(lwy[j] * Y + lphy[j])
        return T

    # truth trajectory (world -> camera)
    R_true, t_true = [], []
    for i in range(N):
        Rwc = _rodrigues([0.1, 1.0, 0.05], 0.01 * i)          # camera-to-world rotation
        cwc = np.array([0.04 * i, 0.01 * np.sin(i), 0.01 * i])  # camera centre in world
        R = Rwc.T
        t = -R @ cwc
        R_true.append(R); t_true.append(t)
    a_true = rng.uniform(-0.05, 0.05, N) if affine else np.zeros(N)
    b_true = rng.uniform(-5.0, 5.0, N) if affine else np.zeros(N)
    a_true[0] = 0.0; b_true[0] = 0.0
    exposure = np.ones(N)

    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    kx = (xs - cx) / fx; ky = (ys - cy) / fy
    gray = np.zeros((N, H, W), dtype=np.float32)
    depth = np.zeros((N, H, W), dtype=np.float64)
    for i in range(N):
        R, t = R_true[i], t_true[i]
        # X_w = R^T (lam*k - t);  n.X_w = d  ->  lam = (d + n.R^T t) / (n.R^T k)
        nR = R @ n  # (R^T)^T n
        lam = (d + nR @ t) / (nR[0] * kx + nR[1] * ky + nR[2])
        Xc = np.stack([lam * kx, lam * ky, lam], axis=-1) - t
        Xw = Xc @ R  # == (R^T Xc)
        T = texture(Xw[..., 0], Xw[..., 1])
        gray[i] = (np.exp(a_true[i]) * exposure[i] * T + b_true[i]).astype(np.float32)
        depth[i] = lam

    # points: random integer pixels, 8-px margin
    pt_host = np.repeat(np.arange(N, dtype=np.int32), pts_per_kf)
    P = pt_host.size
    px = rng.integers(8, W - 8, P); py = rng.integers(8, H - 8, P)
    pt_xy = np.stack([px, py], axis=1).astype(np.float32)
    true_id = 1.0 / depth[pt_host, py, px]
    pt_idepth = true_id * (1.0 + idepth_noise * rng.standard_normal(P))

    def pack(Rs, ts):
        return np.stack([np.concatenate([R.reshape(9), t]) for R, t in zip(Rs, ts)]).astype(np.float64)

    # current estimate = truth + translation noise (KF0 exact); FEJ eval point = an older, slightly different estimate
    t_cam = [t_true[i] + (pose_noise * rng.standard_normal(3) if i > 0 else 0) for i in range(N)]
    R_cam = list(R_true)
    if fej_offset:
        t_eval = [t_cam[i] + (0.5 * pose_noise * rng.standard_normal(3) if 0 < i < N - 1 else 0) for i in range(N)]
        R_eval = [(_rodrigues(rng.standard_normal(3), 2e-4) @ R_cam[i]) if 0 < i < N - 1 else R_cam[i] for i in range(N)]
    else:
        t_eval, R_eval = t_cam, R_cam

    win = {
        "size": np.array([W, H], dtype=np.int32),
        "calib": np.array([fx, fy, cx, cy], dtype=np.float64),
        "iterations": np.array([iterations], dtype=np.int32),
        "update_points_only": np.array([0], dtype=np.int32),
        "frame_evalpt": pack(R_eval, t_eval),
        "frame_cam": pack(R_cam, t_cam),
        "frame_affine": np.zeros((N, 2), dtype=np.float64),   # initial a,b estimate
        "frame_exposure": exposure.astype(np.float64),
        "frame_init": np.zeros(N, dtype=np.uint8),
        "gray": gray,
        "pt_host": pt_host,
        "pt_xy": pt_xy,
        "pt_idepth": pt_idepth.astype(np.float64),
        "truth_frame": pack(R_true, t_true),
        "truth_affine": np.stack([a_true, b_true], axis=1),
        "truth_idepth": true_id,
    }
    if with_gradients:
        win["grad"] = np.stack([gradient_image(gray[i]) for i in range(N)])
    if with_depth:
        win["truth_depth"] = depth          # [N][H][W] depth of the plane along the optical axis (pipeline sanity checks)
    return win


def make_config(name, seed=1234, **kw):
    W, H, N, ppk, iters, affine = CONFIGS[name]
    return make_window(W, H, N, ppk, iters, affine, seed=seed, **kw)


def prepare_scenario(Wi, Hi, Wo, Ho, seed=3):
    """Sensor image with a gamma response, a radial vignette and radial-tangential distortion, to be rectified to Wo x Ho (input of the image
    preparation, SURVEY 8f NEXT #3)."""
    win = make_window(Wi, Hi, 2, 10, 1, False, seed=seed, low_freq=True, with_gradients=False)
    raw = np.clip(win["gray"][0], 0.0, 254.9).astype(np.float32)
    i = np.arange(256, dtype=np.float32)
    lut = (255.0 * (i / 255.0) ** 0.8).astype(np.float32)
    yy, xx = np.mgrid[0:Hi, 0:Wi]
    r2 = ((xx - Wi / 2) ** 2 + (yy - Hi / 2) ** 2) / ((Wi / 2) ** 2 + (Hi / 2) ** 2)
    vig = (1.0 / (1.0 - 0.35 * r2)).astype(np.float32)
    f = 0.9 * Wi
    return dict(size_in=np.array([Wi, Hi], np.int32), size_out=np.array([Wo, Ho], np.int32), calib_in=np.array([f, f, Wi / 2 - 0.5, Hi / 2 - 0.5]),
                calib_out=np.array([0.8 * f * Wo / Wi, 0.8 * f * Wo / Wi, Wo / 2 - 0.5, Ho / 2 - 0.5]), radtan=np.array([-0.28, 0.07, 0.0005, -0.0003]), raw=raw, lut=lut,
                inv_vignette=vig)


def radtan_undistort_map(scn):
    """Source position of every rectified pixel for the scenario's radial-tangential camera (Brown-Conrady k1 k2 p1 p2), NaN outside the sensor:
    the calibration-time input of cmlimg_set_undistort_map when no reference binary is around to provide InternalCalibration's own map."""
    Wi, Hi = scn["size_in"]; Wo, Ho = scn["size_out"]
    fxo, fyo, cxo, cyo = scn["calib_out"]; fxi, fyi, cxi, cyi = scn["calib_in"]
    k1, k2, p1, p2 = scn["radtan"]
    yy, xx = np.mgrid[0:Ho, 0:Wo].astype(np.float32)
    x = (xx - np.float32(cxo)) / np.float32(fxo); y = (yy - np.float32(cyo)) / np.float32(fyo)
    x = x.astype(np.float64); y = y.astype(np.float64)
    r2 = x * x + y * y
    rad = k1 * r2 + k2 * r2 * r2
    xd = x + x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y + y * rad + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y)
    u = (xd * fxi + cxi).astype(np.float32); v = (yd * fyi + cyi).astype(np.float32)
    bad = (u < 0) | (v < 0) | (u >= Wi - 1) | (v >= Hi - 1)
    m = np.stack([u, v], axis=-1)
    m[bad] = np.nan
    return m
