"""Helpers shared by the parity tests: golden loading, decoding of libcmlba's device buffers into the
reference's layouts (DSORawResidualJacobian, AccumulatorApprox 13x13, stitched H/b)."""
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

FRAME_DEV = np.dtype([("evalR", "<f8", 9), ("evalt", "<f8", 3), ("preR", "<f8", 9), ("pret", "<f8", 3), ("state", "<f8", 10), ("state_zero", "<f8", 10),
                      ("state_backup", "<f8", 10), ("state_scaled", "<f8", 10), ("step", "<f8", 10), ("prior", "<f8", 8), ("exposure", "<f8"),
                      ("energy_th", "<f4"), ("keyid", "<i4")])
CTRL = np.dtype([("cur", "<i4"), ("done", "<i4"), ("canbreak", "<i4"), ("failed", "<i4"), ("iteration", "<i4"), ("accepted", "<i4"), ("num_dropped", "<i4"), ("pad0", "<i4"),
                 ("lambda", "<f8"), ("energy_last", "<f8"), ("energy_new", "<f8"), ("energy_first", "<f8"), ("sumA", "<f4"), ("sumB", "<f4"), ("sumT", "<f4"), ("sumR", "<f4"),
                 ("sumNID", "<f8"), ("numID", "<i4"), ("sc_done_count", "<i4"), ("stats", "<f8", 16),
                 ("energyL_last", "<f8"), ("energyL_new", "<f8"), ("prior_energy_pts", "<f8"), ("rejected_at", "<i4"), ("rejected", "<i4"), ("pt_bad", "<i4"), ("asm_done_count", "<i4"), ("acc_done_count", "<i4"), ("pad1", "<i4"), ("final_done", "<i4"), ("pad2", "<i4")])


def load_golden(name):
    from libcml_b200 import cmlw, synth
    win = cmlw.load(os.path.join(GOLDEN, f"{name}_window.cmlw"))
    win["grad"] = np.stack([synth.gradient_image(win["gray"][i]) for i in range(win["gray"].shape[0])])
    gold = cmlw.load(os.path.join(GOLDEN, f"{name}_stages.cmlw"))
    return win, gold


def rel(a, b, floor=1e-30):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor))


def record(test, **vals):
    """Append the MEASURED errors of a parity check to gpurun_out/parity_report.jsonl (copied to profiles/ per round),
    so the asserted gates can be read next to what the kernels actually achieve."""
    import json
    try:
        d = os.path.join(ROOT, "gpurun_out"); os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, **{k: (float(v) if np.ndim(v) == 0 else np.asarray(v).tolist()) for k, v in vals.items()}}) + "\n")
    except OSError:
        pass


def pose_errors(w2c, w2c_ref):
    """(rotation error [absolute, max entry of R - R_ref], translation error RELATIVE to the largest ||t_ref|| of the window, the same after
    removing the common scale factor of the translations).  A pose row is 9 rotation entries (row major) followed by the translation.
    Monocular BA cannot observe the scale of the window (DSO only damps it through the nullspace projection, BA:1348-1420), so the part of
    the solver's conditioning noise that falls into that direction shows up as ONE factor on all translations and inverse depths."""
    a = np.asarray(w2c, np.float64).reshape(-1, 12); b = np.asarray(w2c_ref, np.float64).reshape(-1, 12)
    ta, tb = a[:, 9:], b[:, 9:]
    tn = max(float(np.linalg.norm(tb, axis=1).max()), 1e-30)
    scale = float((ta * tb).sum() / max((tb * tb).sum(), 1e-300))
    return (float(np.abs(a[:, :9] - b[:, :9]).max()), float(np.linalg.norm(ta - tb, axis=1).max() / tn),
            float(np.linalg.norm(ta - scale * tb, axis=1).max() / tn))


def x_noise_floor(gold, pre, N, eps=float(np.finfo(np.float32).eps), trials=16, seed=0):
    """Forward-error floor of the reference's own Gauss-Newton step: its H is accumulated in fp32 (AccumulatorApprox), so
    every entry carries >= 1 ulp of noise.  Returns (rel. change of x when the reference solves the upper instead of the lower
    triangle of ITS matrix, median rel. change of x under an fp32-ulp relative perturbation of H, cond(H))."""
    H, b = solve_reference_system(gold, pre, N)
    Hs = H[4:, 4:]; bs = b[4:]
    lo = np.tril(Hs) + np.tril(Hs, -1).T; up = np.triu(Hs) + np.triu(Hs, 1).T
    xl = np.linalg.solve(lo, bs)
    rng = np.random.default_rng(seed)
    fl = []
    for _ in range(trials):
        P = np.tril(lo * (1 + eps * rng.standard_normal(lo.shape)))
        fl.append(rel(np.linalg.solve(P + np.tril(P, -1).T, bs), xl))
    return rel(np.linalg.solve(up, bs), xl), float(np.median(fl)), float(np.linalg.cond(lo))


def res_key(point, target):
    return np.asarray(point, np.int64) * 64 + np.asarray(target, np.int64)


def map_residuals(dev_point, dev_target, gold_point, gold_target):
    """index array m with dev[m[i]] == gold[i] (matching on (point, target))."""
    dk = res_key(dev_point, dev_target); gk = res_key(gold_point, gold_target)
    order = np.argsort(dk)
    pos = np.searchsorted(dk[order], gk)
    m = order[pos]
    assert np.array_equal(dk[m], gk), "residual sets differ"
    return m


def acc_index(r, c):
    if r > c:
        r, c = c, r
    if c < 10:
        return r * 10 - (r * (r - 1)) // 2 + (c - r)
    if r < 10:
        return 55 + r * 3 + (c - 10)
    rr, cc = r - 10, c - 10
    return 85 + (cc if rr == 0 else (2 + cc if rr == 1 else 5))


_ACC_IDX = np.array([[acc_index(r, c) for c in range(13)] for r in range(13)])


def unpack_acc(acc96, N):
    """[N*N (bin=t*N+h), 96] packed -> reference layout [h + N*t][13][13] (mAccumulatorActive[i].H)."""
    a = np.asarray(acc96, np.float64).reshape(N * N, 96)
    full = a[:, _ACC_IDX]            # [bin,13,13]
    out = np.zeros_like(full)
    for t in range(N):
        for h in range(N):
            out[h + N * t] = full[t * N + h]
    return out


def decode_rj(rj, dbg):
    """device records -> dict with the reference's rJ field names (DSOResidual.h:22-69)."""
    rj = np.asarray(rj, np.float32).reshape(-1, 36); dbg = np.asarray(dbg, np.float32).reshape(-1, 52)
    n = rj.shape[0]
    J = {}
    J["Jpdc"] = np.stack([rj[:, 0:4], rj[:, 10:14]], axis=1)
    J["Jpdxi"] = np.stack([rj[:, 4:10], rj[:, 14:20]], axis=1)
    J["resF"] = dbg[:, 0:8]
    J["JIdx"] = np.stack([dbg[:, 8:16], dbg[:, 16:24]], axis=1)
    J["JabF"] = np.stack([dbg[:, 24:32], dbg[:, 32:40]], axis=1)
    J["Jpdd"] = dbg[:, 40:42]
    J["JIdx2"] = np.stack([dbg[:, 42], dbg[:, 43], dbg[:, 43], dbg[:, 44]], axis=1).reshape(n, 2, 2)
    J["JabJIdx"] = dbg[:, 45:49].reshape(n, 2, 2)
    J["Jab2"] = np.stack([dbg[:, 49], dbg[:, 50], dbg[:, 50], dbg[:, 51]], axis=1).reshape(n, 2, 2)
    return J


def split_sys(sys, n):
    nn = n * n
    return dict(HA=sys[:nn].reshape(n, n), bA=sys[nn:nn + n], HS=sys[nn + n:2 * nn + n].reshape(n, n), bS=sys[2 * nn + n:2 * nn + 2 * n])


def solve_reference_system(gold, pre, N, lam=None):
    """x from the golden H/b exactly as solveLevenbergMarquardt does it (lower triangle, BA:1299-1320)."""
    lam = float(np.float32(1e-5)) if lam is None else lam
    H = gold[pre + "HL_top"] + gold[pre + "HA_top"]
    b = (gold[pre + "bL_top"] + gold[pre + "bM_top"] + gold[pre + "bA_top"] - gold[pre + "b_sc"])[:, 0]
    H = H.copy(); H[np.diag_indices_from(H)] *= (1 + lam); H -= gold[pre + "H_sc"] * (1.0 / (1 + lam))
    return H, b


class DeviceView:
    """Reads the buffers of a DSOBundleAdjustment handle and re-indexes them like the golden dumps."""

    def __init__(self, ba, win):
        self.ba = ba
        self.N = win["frame_evalpt"].shape[0]
        self.P = win["pt_host"].size
        self.n = 8 * self.N + 4
        self.pt_order = ba.read("pt_order", np.int32)          # device point i -> caller index
        self.res_point = ba.read("res_point", np.int32)        # caller point index per device residual
        self.res_target = ba.read("res_target", np.uint8).astype(np.int64)
        self.R = self.res_point.size

    def point_array(self, name, dtype, width=1):
        v = self.ba.read(name, dtype).reshape(self.P, width) if width > 1 else self.ba.read(name, dtype)
        out = np.zeros_like(v)
        out[self.pt_order] = v
        return out

    def frames(self):
        return self.ba.read("frames", FRAME_DEV)

    def ctrl(self):
        return self.ba.read("ctrl", CTRL)[0]

    def map_to(self, gold, pre):
        return map_residuals(self.res_point, self.res_target, gold[pre + "res_point"], gold[pre + "res_target"])
