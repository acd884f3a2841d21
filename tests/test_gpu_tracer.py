"""GPU parity of the immature-point tracer (cmltrc_*, SURVEY.md 8f NEXT #2) against the reference's golden vectors (tests/golden/trace_*.cmlw),
through the C ABI.  Statuses and return codes are integer outputs (compared exactly, see the note on near-ties); values 1e-4 relative."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load():
    return cmlw.load(os.path.join(GOLDEN, "trace_window.cmlw")), cmlw.load(os.path.join(GOLDEN, "trace_golden.cmlw"))


def exposure(win, i):
    return (win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1])


def run_flow(win, g, **params):
    """The SLAM order: add frame f, trace every older point into it, then create the points hosted in f."""
    from libcml_b200 import DSOTracer
    H, W = win["gray"].shape[1:]
    N = win["frame_cam"].shape[0]
    trc = DSOTracer(W, H, win["calib"], **params)
    ids = np.full(win["im_host"].size, -1, np.int64)
    states = {}
    for f in range(N):
        trc.addFrame(f, win["gray"][f], win["frame_cam"][f], exposure(win, f))
        if f > 0:
            hist = trc.traceNewCoarse(f)
            st = trc.getPoints()
            states[f] = st
            want = g[f"trc_status_f{f}"]
            mine = np.full(want.size, -1, np.int32); live = ids >= 0
            mine[live] = st["status"][ids[live]]
            assert np.array_equal(mine[live], want[live]), f"trace statuses differ in frame {f}: {np.nonzero(mine != want)[0][:10]}"
            assert np.array_equal(hist, np.bincount(want[live], minlength=6)[:6])
        mine_pts = np.nonzero(win["im_host"] == f)[0]
        if mine_pts.size:
            ids[mine_pts] = trc.makeNewTracesFrom(f, win["im_xy"][mine_pts])
    return trc, ids, states


def test_point_init_and_trace_match_reference():
    win, g = load()
    trc, ids, states = run_flow(win, g)
    N = win["frame_cam"].shape[0]
    pts = trc.getPoints()
    np.testing.assert_allclose(pts["grad_h"][ids], g["trc_gradH"], rtol=1e-6)
    assert np.array_equal(pts["energy_th"][ids], g["trc_energyTH"])
    for f in range(1, N):
        st = states[f]
        live = (ids >= 0) & (ids < st.size)
        sel = ids[live]
        mine = np.stack([st["idepth_min"][sel], st["idepth_max"][sel], st["last_trace_uv"][sel, 0], st["last_trace_uv"][sel, 1], st["last_trace_pixel_interval"][sel],
                         st["quality"][sel]], axis=1)
        want = g[f"trc_state_f{f}"][live]
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(mine), fin)
        np.testing.assert_allclose(mine[fin], want[fin], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("key", ["trc", "hact"])
def test_activation_matches_reference(key):
    win, g = load()
    params = {} if key == "trc" else dict(min_idepth_h_act=float(g["hact_min_idepth_h_act"][0]))
    trc, ids, _ = run_flow(win, g, **params)
    want = g[f"{key}_opt_rc"]
    todo = np.nonzero(want != -2)[0]
    pts = trc.getPoints()
    assert np.isfinite(pts["idepth_max"][ids[todo]]).all() and not np.isfinite(pts["idepth_max"][ids[want == -2]]).any()
    res = trc.optimizeImmaturePoint(ids[todo], minObs=1)
    assert np.array_equal(res["rc"], want[todo]), np.nonzero(res["rc"] != want[todo])[0][:10]
    ok = res["rc"] == 1
    np.testing.assert_allclose(res["idepth"][ok], g[f"{key}_opt_idepth"][todo][ok], rtol=1e-4)
    assert (res["in_mask"][ok] != 0).all() and (res["in_mask"][~ok] == 0).all()
    assert set(np.unique(res["rc"])) == ({-1, 1} if key == "trc" else {-1, 0, 1})


def test_oracle_agreement_and_bookkeeping():
    """Window the goldens do not cover (seed, size): CUDA vs the numpy restatement on a stride of the points; removal bookkeeping; determinism."""
    import tracer_oracle as T
    from libcml_b200 import DSOTracer, CmlbaError, synth
    W, H, N, per = 320, 240, 4, 400
    win = synth.make_window(W, H, N, 20, 4, True, seed=23, low_freq=True)
    rng = np.random.default_rng(5)
    cams = win["truth_frame"]
    exps = [(win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1]) for i in range(N)]
    grads = [synth.gradient_image(win["gray"][i]) for i in range(N)]
    xy = {h: np.stack([rng.integers(8, W - 8, per), rng.integers(8, H - 8, per)], 1).astype(np.float32) for h in range(N - 1)}

    def flow():
        trc = DSOTracer(W, H, win["calib"])
        ids = {}
        for f in range(N):
            trc.addFrame(10 + f, win["gray"][f], cams[f], exps[f])
            if f > 0:
                trc.traceNewCoarse(10 + f)
            if f < N - 1:
                ids[f] = trc.makeNewTracesFrom(10 + f, xy[f])
        return trc, ids
    trc, ids = flow()
    pts = trc.getPoints()
    trc2, _ = flow()
    pts2 = trc2.getPoints()
    assert pts.tobytes() == pts2.tobytes()                       # deterministic
    checked = 0
    for h in range(N - 1):
        for k in range(0, per, 7):
            o = T.ImmaturePoint(h, xy[h][k], grads[h])
            for f in range(h + 1, N):
                T.trace(o, win["calib"], cams[h], cams[f], exps[h], exps[f], win["gray"][h], win["gray"][f])
            m = pts[ids[h][k]]
            assert m["status"] == o.status
            for a, b in ((m["idepth_min"], o.idmin), (m["idepth_max"], o.idmax), (m["quality"], o.quality), (m["last_trace_pixel_interval"], o.interval)):
                assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-4 * max(abs(b), 1e-3)
            if np.isfinite(o.idmax):
                rc, idp, states = T.optimize_immature_point(o, win["calib"], cams, exps, grads, range(N - 1, -1, -1))
                r = trc.optimizeImmaturePoint([ids[h][k]])[0]
                assert r["rc"] == rc
                if rc == 1:
                    assert abs(r["idepth"] - idp) <= 1e-4 * idp
                    assert r["in_mask"] == sum(1 << t for t, s in enumerate(states) if s == T.RES_IN)
            checked += 1
    assert checked > 150
    # bookkeeping
    trc.removePoints(ids[0][:10])
    assert (trc.getPoints()["host_frame_slot"][ids[0][:10]] == -1).all()
    with pytest.raises(CmlbaError):
        trc.optimizeImmaturePoint(ids[0][:1])                    # removed point
    trc.removeFrame(11)
    assert (trc.getPoints()["host_frame_slot"][ids[1]] == -1).all()
    with pytest.raises(CmlbaError):
        trc.traceNewCoarse(11)
    with pytest.raises(CmlbaError):
        trc.makeNewTracesFrom(10, [[1.0, 1.0]])                  # pattern would leave the image


def test_activate_points_matches_reference():
    """activatePoints: minimum-distance adaptation, distance-map gating in the reference's iteration order, activation, removals (integer outputs: exact)."""
    win, g = load()
    a = cmlw.load(os.path.join(GOLDEN, "activate_golden.cmlw"))
    trc, ids, _ = run_flow(win, g)
    N = win["frame_cam"].shape[0]; P = win["im_host"].size
    fx, fy, cx, cy = win["calib"]
    axy = []
    for h, xy, idp in zip(a["act_host"], a["act_xy"], a["act_idepth"]):          # the caller projects its active points into the last frame
        c0, c1 = win["frame_cam"][h], win["frame_cam"][N - 1]
        R = c1[:9].reshape(3, 3) @ c0[:9].reshape(3, 3).T; t = c1[9:] - R @ c0[9:]
        X = R @ (np.array([(xy[0] - cx) / fx, (xy[1] - cy) / fy, 1.0]) / idp) + t
        axy.append((fx * X[0] / X[2] + cx, fy * X[1] / X[2] + cy))
    back = {int(ids[p]): p for p in range(P)}
    act_ids, act, rem_ids, st = trc.activatePoints(N - 1, axy, ids[a["act_order"]], desiredPointDensity=int(a["desired_density"][0]))
    mapped = {back[int(i)] for i in act_ids}
    assert mapped == set(np.nonzero(a["act_mapped"])[0])
    assert set(range(P)) - mapped - {back[int(i)] for i in rem_ids} == set(np.nonzero(a["act_still_immature"])[0])
    assert st.current_minimum_distance == a["act_min_distance"][0] and st.urgently_need_new_points == a["act_urgent"][0]
    assert st.num_mapped == len(mapped) and st.num_to_optimize >= st.num_mapped and (act["rc"] == 1).all()
    for i, r in zip(act_ids, act):
        assert abs(r["idepth"] - a["act_idepth_out"][back[int(i)]]) <= 1e-4 * r["idepth"]
    gone = np.concatenate([act_ids, rem_ids])
    assert (trc.getPoints()["host_frame_slot"][gone] == -1).all()                     # they left the immature set


def test_non_default_parameters_against_oracle():
    """Other thresholds / step sizes / search lengths and a 250x187 image: CUDA vs the numpy restatement on every third point."""
    import tracer_oracle as T
    from libcml_b200 import DSOTracer, synth
    W, H, N, per = 250, 187, 4, 240
    win = synth.make_window(W, H, N, 20, 4, True, seed=37, low_freq=True, with_gradients=False)
    rng = np.random.default_rng(8)
    cams = win["truth_frame"]
    exps = [(win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1]) for i in range(N)]
    grads = [synth.gradient_image(win["gray"][i]) for i in range(N)]
    xy = {h: np.stack([rng.integers(8, W - 8, per), rng.integers(8, H - 8, per)], 1).astype(np.float32) for h in range(N - 1)}
    dev = dict(huber_threshold=6.0, outlier_th=100.0, max_pix_search=0.05, max_slack_interval=2.5, trace_step_size=1.0, min_improvement_factor=1.5,
               min_trace_test_radius=3.0, extra_slack_on_th=1.1, min_idepth_h_act=60.0, gn_iterations=2)
    ora = dict(T.DEFAULTS, huber=6.0, outlier_th=100.0, max_pix_search=float(np.float32(0.05)), max_slack_interval=2.5, min_improvement=1.5, test_radius=3.0,
               extra_slack=float(np.float32(1.1)), min_idepth_h_act=60.0, gn_iterations=2)
    trc = DSOTracer(W, H, win["calib"], **dev)
    ids = {}
    for f in range(N):
        trc.addFrame(f, win["gray"][f], cams[f], exps[f])
        if f > 0:
            trc.traceNewCoarse(f)
        if f < N - 1:
            ids[f] = trc.makeNewTracesFrom(f, xy[f])
    pts = trc.getPoints()
    seen = set()
    for h in range(N - 1):
        for k in range(0, per, 3):
            o = T.ImmaturePoint(h, xy[h][k], grads[h], ora)
            for f in range(h + 1, N):
                T.trace(o, win["calib"], cams[h], cams[f], exps[h], exps[f], win["gray"][h], win["gray"][f], ora)
            m = pts[ids[h][k]]
            assert m["status"] == o.status and m["energy_th"] == o.energyTH
            seen.add(o.status)
            for a, b in ((m["idepth_min"], o.idmin), (m["idepth_max"], o.idmax), (m["quality"], o.quality)):
                assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-4 * max(abs(b), 1e-3)
            if np.isfinite(o.idmax):
                rc, idp, _ = T.optimize_immature_point(o, win["calib"], cams, exps, grads, range(N - 1, -1, -1), p=ora)
                r = trc.optimizeImmaturePoint([ids[h][k]])[0]
                assert r["rc"] == rc and (rc != 1 or abs(r["idepth"] - idp) <= 1e-4 * idp)
    assert len(seen) >= 3


def test_dead_slots_are_compacted_and_ids_stay_stable():
    """Removed points leave dead device slots; once more than a quarter of the slots is dead the next makeNewTracesFrom squeezes them out.
    Ids are never reused and keep addressing the same points: the records of the survivors are bit-identical before and after the
    compaction, removed ids read back as host_frame_slot = -1, and a trace pass after the compaction equals the one of an un-compacted handle."""
    from libcml_b200 import DSOTracer
    win, _ = load()
    H, W = win["gray"].shape[1:]
    rng = np.random.default_rng(7)
    xy = np.stack([rng.uniform(8, W - 9, 6000), rng.uniform(8, H - 9, 6000)], axis=1).astype(np.float32)

    def make():
        t = DSOTracer(W, H, win["calib"])
        t.addFrame(0, win["gray"][0], win["frame_cam"][0], exposure(win, 0))
        i = t.makeNewTracesFrom(0, xy)
        t.addFrame(1, win["gray"][1], win["frame_cam"][1], exposure(win, 1))
        t.traceNewCoarse(1)
        return t, i
    trc, ids = make()
    ref, ids_ref = make()
    before = trc.getPoints()
    gone = ids[::2]                                             # half of the points: far above the 25 % threshold
    trc.removePoints(gone); ref.removePoints(gone[:100])        # the reference handle stays below the threshold (no compaction)
    ref.removePoints(gone[100:1000])
    extra = trc.makeNewTracesFrom(1, xy[:10])                   # triggers the compaction
    assert extra[0] == 6000 and trc.numPoints() == 6010         # ids are not reused
    after = trc.getPoints()
    keep = ids[1::2]
    for k in before.dtype.names:
        assert np.array_equal(before[k][keep], after[k][keep], equal_nan=True), k
    assert (after["host_frame_slot"][gone] == -1).all()
    trc.addFrame(2, win["gray"][2], win["frame_cam"][2], exposure(win, 2)); ref.addFrame(2, win["gray"][2], win["frame_cam"][2], exposure(win, 2))
    trc.traceNewCoarse(2); ref.traceNewCoarse(2)
    a, b = trc.getPoints(), ref.getPoints()
    for k in a.dtype.names:
        assert np.array_equal(a[k][keep], b[k][keep], equal_nan=True), k
    with pytest.raises(Exception):
        trc.optimizeImmaturePoint(gone[:1])                      # a removed id stays removed
