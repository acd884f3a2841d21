"""The direct-odometry pipeline end to end on the device, on a synthetic sequence with known truth (SURVEY.md 8f rows chained together):
image preparation -> pixel selection -> immature-point tracing -> activation -> photometric bundle adjustment -> coarse tracking of the next
frame.  Every frame crosses PCIe once (cmlimg); all consumers read the device-resident levels.  Prints one JSON line with the accuracy of every
stage against the truth and the per-stage call times.  `run()` is also what tests/test_gpu_pipeline.py asserts on."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import CaptureImageGenerator, DSOBundleAdjustment, DSOTracer, DSOTracker, PixelSelector, synth  # noqa: E402


def run(W=320, H=240, N=6, density=400, seed=5):
    win = synth.make_window(W, H, N, 10, 4, False, seed=seed, low_freq=True, with_gradients=False, with_depth=True)
    truth = win["truth_frame"]; K = win["calib"]
    ex = [(1.0, 0.0, 0.0)] * N
    T = {}
    def lap(name, t0):
        T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    gens = [CaptureImageGenerator(W, H) for _ in range(N)]           # one generator per frame keeps every frame's levels alive on the device
    caps = []
    for f in range(N):
        t0 = time.perf_counter(); caps.append(gens[f].generate(win["gray"][f])); lap("prepare", t0)
    KF = N - 1                                                        # frames 0..N-2 are keyframes, frame N-1 is tracked at the end
    sel = PixelSelector(W, H)
    trc = DSOTracer(W, H, K)
    ids, corners, types = {}, {}, {}
    for f in range(KF):
        t0 = time.perf_counter(); trc.addFrameDevice(f, caps[f], truth[f], ex[f]); lap("tracer.addFrame", t0)
        if f > 0:
            t0 = time.perf_counter(); trc.traceNewCoarse(f); lap("tracer.trace", t0)
        if f < KF - 1:
            t0 = time.perf_counter(); xy, ty = sel.compute(caps[f], density); lap("selector", t0)
            t0 = time.perf_counter(); ids[f] = trc.makeNewTracesFrom(f, xy); corners[f] = xy; types[f] = ty; lap("tracer.makeNewTraces", t0)
    all_ids = np.concatenate([ids[f] for f in ids]); all_xy = np.concatenate([corners[f] for f in corners]); all_host = np.concatenate([np.full(ids[f].size, f) for f in ids])
    all_ty = np.concatenate([types[f] for f in types])
    # activatePoints: no active points yet (first window), so the distance map only spaces the new points among themselves
    t0 = time.perf_counter(); act_ids, act, rem_ids, st = trc.activatePoints(KF - 1, np.zeros((0, 2)), all_ids, desiredPointDensity=2 * density, types=all_ty); lap("tracer.activatePoints", t0)
    where = {int(i): k for k, i in enumerate(all_ids)}
    sel_idx = np.array([where[int(i)] for i in act_ids])
    a_xy, a_host, a_id = all_xy[sel_idx], all_host[sel_idx], act["idepth"].astype(np.float64)
    ok = np.ones(a_id.size, bool)
    true_id = 1.0 / win["truth_depth"][a_host, a_xy[:, 1].astype(int), a_xy[:, 0].astype(int)]
    rel = np.abs(a_id - true_id) / true_id
    out = {"frames": N, "selected_per_keyframe": [int(corners[f].shape[0]) for f in corners], "traced_points": int(all_ids.size), "activated": int(ok.sum()), "removed": int(rem_ids.size),
           "minimum_distance": float(st.current_minimum_distance),
           "activation_idepth_median_rel_err": float(np.median(rel)), "activation_idepth_p90_rel_err": float(np.quantile(rel, 0.9))}
    # bundle adjustment over the keyframes with the activated points, poses perturbed
    rng = np.random.default_rng(seed)
    start = truth[:KF].copy(); start[1:, 9:] += 2e-3 * rng.standard_normal((KF - 1, 3))
    ba = DSOBundleAdjustment(device=0, iterations=6)
    ba.setCalibration(*[float(v) for v in K], W, H)
    t0 = time.perf_counter()
    for f in range(KF):
        ba.addNewFrameDevice(f, start[f], 0.0, 0.0, 1.0, caps[f].devicePtr("texel0"), f == 0)
    ba.addPoints(np.arange(a_id.size), a_host, a_xy, a_id)
    lap("ba.add", t0)
    t0 = time.perf_counter(); ok_run = ba.run(start, iterations=6); lap("ba.run", t0)
    fr = ba.getFrames(); bp = ba.getPoints()
    def reproj_px(cams, host, xy, idepth, t_cams, t_idepth, targets):
        """Mean pixel distance between where the estimate (poses + inverse depths) and the truth put every point in every target frame: free of the
        gauge a monocular BA leaves (global scale: inverse depths and translations scale together)."""
        fx, fy, cx, cy = K
        def proj(c, hs, px, idp, tg):
            Rh, th = c[hs][:, :9].reshape(-1, 3, 3), c[hs][:, 9:]
            Rt, tt = c[tg][:9].reshape(3, 3), c[tg][9:]
            ray = np.stack([(px[:, 0] - cx) / fx, (px[:, 1] - cy) / fy, np.ones(len(px))], 1) / idp[:, None]
            Xw = np.einsum("nji,nj->ni", Rh, ray - th)
            Xt = Xw @ Rt.T + tt
            return np.stack([fx * Xt[:, 0] / Xt[:, 2] + cx, fy * Xt[:, 1] / Xt[:, 2] + cy], 1)
        errs = []
        for tg in targets:
            m = host != tg
            errs.append(np.linalg.norm(proj(cams, host[m], xy[m], idepth[m], tg) - proj(t_cams, host[m], xy[m], t_idepth[m], tg), axis=1) if m.any() else np.zeros(0))
        return float(np.concatenate(errs).mean())
    kept = bp["id"]
    out.update({"ba_ok": bool(ok_run), "ba_iterations": int(ba.last_result.iterations_done),
                "ba_reproj_px_before": reproj_px(start, a_host, a_xy, a_id, truth, true_id, range(KF)),
                "ba_reproj_px_after": reproj_px(fr["world_to_cam"], a_host[kept], a_xy[kept], bp["idepth"], truth, true_id[kept], range(KF)),
                "ba_energy_first": float(ba.last_result.energy_first), "ba_energy_last": float(ba.last_result.energy_last)})
    # coarse tracking of the last frame against the newest keyframe, depth map from the BA's points
    ref = KF - 1
    trk = DSOTracker(W, H, K)
    keep = bp["id"]
    unc = np.full(keep.size, 1e-3)
    t0 = time.perf_counter()
    trk.makeCoarseDepthL0Device(caps[ref], fr["world_to_cam"][ref], ex[ref], fr["world_to_cam"], a_host[keep], a_xy[keep], bp["idepth"], unc)
    lap("tracker.makeCoarseDepth", t0)
    guess = truth[N - 1].copy(); guess[9:] += 4e-3 * rng.standard_normal(3)
    t0 = time.perf_counter(); trk.setFrameDevice(caps[N - 1], 1.0); r = trk.optimize(guess, (0.0, 0.0)); lap("tracker.track", t0)
    cams_before = np.concatenate([fr["world_to_cam"], guess[None]]); cams_after = np.concatenate([fr["world_to_cam"], r.camera[None]])
    t_cams = np.concatenate([truth[:KF], truth[N - 1:N]])
    args = (a_host[keep], a_xy[keep], bp["idepth"], t_cams, true_id[keep], [KF])        # every BA point projected into the tracked frame
    out.update({"track_ok": bool(r.isCorrect), "track_iterations": int(r.iterations), "track_reproj_px_before": reproj_px(cams_before, *args),
                "track_reproj_px_after": reproj_px(cams_after, *args), "call_ms": {k: round(v, 3) for k, v in T.items()}})
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
