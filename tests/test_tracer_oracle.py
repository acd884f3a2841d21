"""The numpy restatement of DSOTracer (oracle/tracer_oracle.py) against the reference's golden vectors (tests/golden/trace_*.cmlw, made by
oracle/make_golden.py tracer from the unmodified reference).  CPU only; a stride of the points keeps it to a few seconds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw, synth  # noqa: E402
import tracer_oracle as T  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
STRIDE = 3


def load():
    win = cmlw.load(os.path.join(GOLDEN, "trace_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "trace_golden.cmlw"))
    N = win["frame_cam"].shape[0]
    grads = [synth.gradient_image(win["gray"][i]) for i in range(N)]
    exps = [(win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1]) for i in range(N)]
    return win, g, N, grads, exps


def traced_points(win, g, N, grads, exps, sel):
    pts = {p: T.ImmaturePoint(int(win["im_host"][p]), win["im_xy"][p], grads[win["im_host"][p]]) for p in sel}
    for f in range(1, N):
        for p in sel:
            pt = pts[p]
            if pt.host < f:
                T.trace(pt, win["calib"], win["frame_cam"][pt.host], win["frame_cam"][f], exps[pt.host], exps[f], win["gray"][pt.host], win["gray"][f])
            assert pt.status == g[f"trc_status_f{f}"][p], (f, p)                      # status: exact
            want = g[f"trc_state_f{f}"][p]
            mine = np.array([pt.idmin, pt.idmax, pt.uv[0], pt.uv[1], pt.interval, pt.quality])
            fin = np.isfinite(want)
            assert np.array_equal(np.isfinite(mine), fin)
            np.testing.assert_allclose(mine[fin], want[fin], rtol=1e-4, atol=1e-6)
    return pts


def test_point_init_and_trace_match_reference():
    win, g, N, grads, exps = load()
    sel = range(0, win["im_host"].size, STRIDE)
    assert g["trc_created"].all()
    pts = traced_points(win, g, N, grads, exps, sel)
    for p in sel:
        np.testing.assert_allclose(pts[p].gradH.ravel(), g["trc_gradH"][p], rtol=1e-6)
        np.testing.assert_allclose(pts[p].weights, g["trc_weights"][p], rtol=1e-6)
        assert pts[p].energyTH == g["trc_energyTH"][p]


def test_activation_matches_reference():
    win, g, N, grads, exps = load()
    sel = range(1, win["im_host"].size, STRIDE)
    pts = traced_points(win, g, N, grads, exps, sel)
    for key, params in (("trc", T.DEFAULTS), ("hact", dict(T.DEFAULTS, min_idepth_h_act=float(g["hact_min_idepth_h_act"][0])))):
        seen = set()
        for p in sel:
            want = int(g[f"{key}_opt_rc"][p])
            if want == -2:
                assert not (np.isfinite(pts[p].idmax) and np.isfinite(pts[p].idmin))
                continue
            rc, idp, _ = T.optimize_immature_point(pts[p], win["calib"], win["frame_cam"], exps, grads, range(N), p=params)
            assert rc == want, (key, p)
            seen.add(rc)
            if rc == 1:
                assert abs(idp - g[f"{key}_opt_idepth"][p]) <= 1e-4 * g[f"{key}_opt_idepth"][p]
        assert seen == ({-1, 1} if key == "trc" else {-1, 0, 1})


def test_activate_points_matches_reference():
    """activatePoints (minimum-distance adaptation, distance-map gating in the reference's iteration order, activation, removal)."""
    win, g, N, grads, exps = load()
    a = cmlw.load(os.path.join(GOLDEN, "activate_golden.cmlw"))
    P = win["im_host"].size
    pts = {p: T.ImmaturePoint(int(win["im_host"][p]), win["im_xy"][p], grads[win["im_host"][p]]) for p in range(P)}
    for f in range(1, N):
        for p in range(P):
            if pts[p].host < f:
                T.trace(pts[p], win["calib"], win["frame_cam"][pts[p].host], win["frame_cam"][f], exps[pts[p].host], exps[f], win["gray"][pts[p].host], win["gray"][f])
    fx, fy, cx, cy = win["calib"]
    axy = []
    for h, xy, idp in zip(a["act_host"], a["act_xy"], a["act_idepth"]):
        R, t = T.rel_pose(win["frame_cam"][h], win["frame_cam"][N - 1])
        X = R @ (np.array([(xy[0] - cx) / fx, (xy[1] - cy) / fy, 1.0]) / idp) + t
        axy.append((fx * X[0] / X[2] + cx, fy * X[1] / X[2] + cy))
    mapped, removed, md, urgent = T.activate_points(pts, list(a["act_order"]), np.ones(P), axy, win["calib"], win["frame_cam"], exps, grads, range(N - 1, -1, -1), N - 1, 2.0,
                                                    int(a["desired_density"][0]))
    assert set(mapped) == set(np.nonzero(a["act_mapped"])[0])
    assert set(range(P)) - set(mapped) - set(removed) == set(np.nonzero(a["act_still_immature"])[0])
    assert md == a["act_min_distance"][0] and urgent == bool(a["act_urgent"][0])
    for i, v in mapped.items():
        assert abs(v - a["act_idepth_out"][i]) <= 1e-4 * v
