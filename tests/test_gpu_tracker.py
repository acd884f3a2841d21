"""GPU parity of the coarse tracker (cmltrk_*, SURVEY.md 8f NEXT #1) against the reference's golden vectors (tests/golden/track_*.cmlw) and
the numpy restatement (oracle/tracker_oracle.py), through the C ABI."""
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
    return cmlw.load(os.path.join(GOLDEN, "track_window.cmlw")), cmlw.load(os.path.join(GOLDEN, "track_golden.cmlw"))


def make_tracker(win, **params):
    from libcml_b200 import DSOTracker
    H, W = win["gray"].shape[1:]
    trk = DSOTracker(W, H, win["calib"], **params)
    ref, new = int(win["track_ref"][0]), int(win["track_new"][0])
    keep = win["pt_host"] != new
    trk.makeCoarseDepthL0(win["gray"][ref], win["frame_cam"][ref], (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1]), win["frame_cam"],
                          win["pt_host"][keep], win["pt_xy"][keep], win["pt_idepth"][keep], win["pt_uncertainty"][keep])
    return trk, ref, new


def test_pyramid_and_coarse_depth_match_reference():
    import tracker_oracle as T
    win, g = load()
    trk, ref, new = make_tracker(win)
    L = g["trk_K"].shape[0]
    assert np.array_equal(trk.read("levels_wh", np.int32), g["trk_levels_wh"])
    assert np.abs(trk.read("K", np.float64).reshape(L, 4) - g["trk_K"]).max() < 1e-12
    assert np.array_equal(trk.read("pc_n", np.int32), g["trk_pc_n"])
    for l in range(L):
        pc = trk.read(f"pc{l}", np.float32).reshape(-1, 4)
        assert np.array_equal(pc[:, [0, 1, 3]], g[f"trk_pc{l}"][:, [0, 1, 3]])                # pixel and colour: bit-exact
        assert np.abs(pc[:, 2] - g[f"trk_pc{l}"][:, 2]).max() <= 2 ** -22                     # inverse depth: fixed-point splat vs fp32 running sums
    # derivative pyramid of the frame to track: bit-exact against the restatement (which is bit-exact against the reference, make_golden.py)
    trk.setFrame(win["gray"][new], win["frame_exposure"][new])
    pyr = T.build_pyramid(win["gray"][new], L)
    for l in range(L):
        h, w = pyr[l][0].shape
        grad = trk.read(f"grad{l}", np.float32).reshape(h, w, 4)
        assert np.array_equal(grad[:, :, :3], pyr[l][1])


@pytest.mark.parametrize("case", ["a", "b", "c"])
@pytest.mark.parametrize("cluster", [16, 8, 1])
def test_optimize_matches_reference(case, cluster):
    win, g = load()
    trk, ref, new = make_tracker(win, cluster_ctas=cluster)
    r = trk.optimize(g[f"{case}_init_cam"], g[f"{case}_new_affine"], gray=win["gray"][new], exposure_time=win["frame_exposure"][new])
    assert r.isCorrect == bool(g[f"{case}_trk_isCorrect"][0])
    assert list(r.numTermsInE) == list(g[f"{case}_trk_numTermsInE"])
    assert list(r.numSaturated) == list(g[f"{case}_trk_numSaturated"])
    assert list(r.numRobust) == list(g[f"{case}_trk_numRobust"])
    np.testing.assert_allclose(r.E, g[f"{case}_trk_E"], rtol=1e-4)                            # fp tolerance of the path: 1e-4 relative
    if not r.isCorrect:
        assert np.array_equal(r.camera, g[f"{case}_init_cam"])                                # camera untouched on failure
        return
    assert np.abs(r.camera - g[f"{case}_trk_cam"]).max() < 1e-5
    np.testing.assert_allclose(r.exposure, g[f"{case}_trk_affine"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r.levelCutoffRepeat, g[f"{case}_trk_levelCutoffRepeat"])
    np.testing.assert_allclose(r.flowVector, g[f"{case}_trk_flow"], rtol=1e-4)
    np.testing.assert_allclose(r.relAff, g[f"{case}_trk_relAff"][:2], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r.covariance, g[f"{case}_trk_covariance"], rtol=1e-3)
    assert r.tooManySaturated == bool(g[f"{case}_trk_tooManySaturated"][0])
    assert r.kernel_launches == 3                                                             # pyramid (2) + the whole optimisation (1)


def test_candidates_batch_equals_single_and_is_deterministic():
    """K start poses in one launch (trackWithMotionModel's candidate loop): every cluster must reproduce the single-pose result bit for bit."""
    win, g = load()
    trk, ref, new = make_tracker(win)
    trk.setFrame(win["gray"][new], win["frame_exposure"][new])
    cams = np.stack([g["a_init_cam"], g["c_init_cam"], g["a_init_cam"], g["b_init_cam"]])
    aff = np.stack([g["a_new_affine"], g["c_new_affine"], g["a_new_affine"], g["b_new_affine"]])
    batch = trk.optimize(cams, aff)
    again = trk.optimize(cams, aff)
    for k, case in enumerate(["a", "c", "a", "b"]):
        single = trk.optimize(g[f"{case}_init_cam"], g[f"{case}_new_affine"])
        for r in (batch[k], again[k]):
            assert np.array_equal(r.camera, single.camera) and np.array_equal(r.exposure, single.exposure) and np.array_equal(r.E, single.E)
            assert r.isCorrect == single.isCorrect and r.iterations == single.iterations


def test_oracle_agreement_on_fresh_window():
    """A window the goldens do not cover (other seed and size, fixed brightness b): CUDA vs the numpy restatement."""
    import tracker_oracle as T
    from libcml_b200 import synth
    N, seed = 3, 7
    win = synth.make_window(320, 240, N, 300, 4, True, seed=seed, low_freq=True)
    rng = np.random.default_rng(seed)
    ref, new = N - 2, N - 1
    cams = win["truth_frame"]
    idepth = win["truth_idepth"] * (1 + 0.01 * rng.standard_normal(win["pt_host"].size))
    unc = 1.0 / (rng.uniform(50, 5000, win["pt_host"].size) + 0.01)
    start = cams[new].copy(); start[9:] += 2e-3 * rng.standard_normal(3)
    from libcml_b200 import DSOTracker
    for opt_a, opt_b in ((1, 1), (1, 0), (0, 1), (0, 0)):
        trk = DSOTracker(320, 240, win["calib"], optimize_a=opt_a, optimize_b=opt_b)
        keep = win["pt_host"] != new
        ref_exp = (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1])
        trk.makeCoarseDepthL0(win["gray"][ref], cams[ref], ref_exp, cams, win["pt_host"][keep], win["pt_xy"][keep], idepth[keep], unc[keep])
        r = trk.optimize(start, (0.0, 0.0), gray=win["gray"][new], exposure_time=win["frame_exposure"][new])
        L = 5
        pcs = [trk.read(f"pc{l}", np.float32).reshape(-1, 4) for l in range(L)]
        pyr = T.build_pyramid(win["gray"][new], L)
        Rr, tr = cams[ref][:9].reshape(3, 3), cams[ref][9:]
        Rn, tn = start[:9].reshape(3, 3), start[9:]
        R0 = Rn @ Rr.T
        o = T.optimize(pcs, [p[1] for p in pyr], [T.level_K(win["calib"], l) for l in range(L)], (R0, tn - R0 @ tr), ref_exp, (win["frame_exposure"][new], 0.0, 0.0),
                       params=dict(optimize_a=bool(opt_a), optimize_b=bool(opt_b)))
        assert r.isCorrect == o["isCorrect"] and r.isCorrect
        cam = np.concatenate([(o["R"] @ Rr).ravel(), o["R"] @ tr + o["t"]])
        assert np.abs(r.camera - cam).max() < 2e-5, (opt_a, opt_b)
        np.testing.assert_allclose(r.exposure, o["exposure"][1:], rtol=1e-3, atol=1e-3)
        assert list(r.numTermsInE) == list(o["numTermsInE"])
        np.testing.assert_allclose(r.E, o["E"], rtol=1e-3)


def test_last_residual_rmse_gate():
    """mLastResidual.isCorrect: a level whose rmse exceeds 1.5 x the last frame's aborts the optimisation (DSOTracker.cpp:190-196)."""
    import tracker_oracle as T
    win, g = load()
    trk, ref, new = make_tracker(win)
    trk.setFrame(win["gray"][new], win["frame_exposure"][new])
    free = trk.optimize(g["a_init_cam"], g["a_new_affine"])
    assert free.isCorrect
    trk.mLastResidual = free                                  # same frame again: rmse equal to the last one -> passes, identical result
    again = trk.optimize(g["a_init_cam"], g["a_new_affine"])
    assert again.isCorrect and np.array_equal(again.camera, free.camera)

    class Fake:                                               # a much better last frame: the coarsest level already fails the gate
        isCorrect = True
        E = free.E
        def rmse(self, l):
            return 0.1 * free.E[l] / free.numTermsInE[l]
    trk.mLastResidual = Fake()
    gated = trk.optimize(g["a_init_cam"], g["a_new_affine"])
    assert not gated.isCorrect and 0 < gated.iterations < free.iterations and np.array_equal(gated.camera, g["a_init_cam"])
    # the restatement stops at the same place
    L = 5
    pcs = [g[f"trk_pc{l}"] for l in range(L)]
    pyr = T.build_pyramid(win["gray"][new], L)
    Rr, tr = win["frame_cam"][ref][:9].reshape(3, 3), win["frame_cam"][ref][9:]
    Rn, tn = g["a_init_cam"][:9].reshape(3, 3), g["a_init_cam"][9:]
    R0 = Rn @ Rr.T
    o = T.optimize(pcs, [p[1] for p in pyr], [g["trk_K"][l] for l in range(L)], (R0, tn - R0 @ tr),
                   (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1]), (win["frame_exposure"][new], 0.0, 0.0),
                   last_rmse=[Fake().rmse(l) for l in range(L)])
    assert not o["isCorrect"] and o["iterations"] == gated.iterations and list(o["numTermsInE"]) == list(gated.numTermsInE)


def test_track_with_motion_model_candidate_loop():
    """The candidate loop of trackWithMotionModel (DSOTracker.h:240-360): a hopeless first pose is skipped, the loop stops at the first pose that
    is good enough, and nothing is returned in failure mode 0 when every pose fails."""
    win, g = load()
    trk, ref, new = make_tracker(win)
    trk.setFrame(win["gray"][new], win["frame_exposure"][new])
    good, far = g["a_init_cam"], g["c_init_cam"]
    single = trk.optimize(good, (0.0, 0.0))
    trk.mLastResidual = None
    ok, cam, ex, res = trk.trackWithMotionModel([far, good, good])
    assert ok and trk.lastTriedCameras == 2 and np.array_equal(cam, single.camera) and np.array_equal(ex, single.exposure) and res.isCorrect
    assert trk.mLastCoarseRMSE == single.rmse() and trk.mFirstRMSE == single.rmse()
    ok, cam, ex, res = trk.trackWithMotionModel([good, far])                 # first pose already good enough: the loop stops after one optimisation
    assert ok and trk.lastTriedCameras == 1
    trk2, _, _ = make_tracker(win)
    trk2.setFrame(win["gray"][new], win["frame_exposure"][new])
    ok, cam, ex, res = trk2.trackWithMotionModel([far, far])
    assert not ok and cam is None and trk2.lastTriedCameras == 2
    # the C entry point (cmltrk_track_with_motion_model) against the Python loop over cmltrk_optimize it was written from: same decisions, same bits
    for poses, mode in (([far, good, good], 0), ([good, far], 0), ([far, far], 1)):
        ta, _, _ = make_tracker(win); tb, _, _ = make_tracker(win)
        ta.setFrame(win["gray"][new], win["frame_exposure"][new]); tb.setFrame(win["gray"][new], win["frame_exposure"][new])
        oa, ca, ea, ra = ta.trackWithMotionModel(poses, failure_mode=mode)
        ob, cb, eb, rb = tb.trackWithMotionModelPy(poses, failure_mode=mode)
        assert oa == ob and ta.lastTriedCameras == tb.lastTriedCameras
        if oa:
            assert np.array_equal(ca, cb) and np.array_equal(ea, eb) and ra.isCorrect == rb.isCorrect


def test_error_paths():
    from libcml_b200 import DSOTracker, CmlbaError
    win, g = load()
    H, W = win["gray"].shape[1:]
    trk = DSOTracker(W, H, win["calib"])
    with pytest.raises(CmlbaError):
        trk.optimize(g["a_init_cam"], g["a_new_affine"])            # no reference / frame yet (CMLTRK_ERR_STATE)
    with pytest.raises(ValueError):
        trk.setFrame(np.zeros((H, W + 1), np.float32))
    with pytest.raises(CmlbaError):
        DSOTracker(4, 4, win["calib"])


@pytest.mark.parametrize("W,H,levels", [(250, 187, 0), (200, 150, 3), (320, 240, 4)])
def test_odd_sizes_and_fewer_levels_against_oracle(W, H, levels):
    """Image sizes that are not multiples of 2^levels (floor halving, clipped pyramid tiles) and pyramids with fewer than five levels
    (maxLevel = levels - 1, DSOTracker.cpp:24), non-default parameters: CUDA vs the numpy restatement."""
    import tracker_oracle as T
    from libcml_b200 import DSOTracker, synth
    N, seed = 3, 13
    win = synth.make_window(W, H, N, 350, 4, True, seed=seed, low_freq=True, with_gradients=False)
    rng = np.random.default_rng(seed)
    ref, new = N - 2, N - 1
    cams = win["truth_frame"]
    keep = win["pt_host"] != new
    idepth = win["truth_idepth"] * (1 + 0.01 * rng.standard_normal(win["pt_host"].size))
    unc = 1.0 / (rng.uniform(50, 5000, win["pt_host"].size) + 0.01)
    start = cams[new].copy(); start[9:] += 2e-3 * rng.standard_normal(3)
    ref_exp = (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1])
    params = dict(huber_threshold=7.0, cutoff_threshold=15.0, scale_translation=0.4, scale_light_b=500.0)
    trk = DSOTracker(W, H, win["calib"], levels=levels, **params)
    trk.makeCoarseDepthL0(win["gray"][ref], cams[ref], ref_exp, cams, win["pt_host"][keep], win["pt_xy"][keep], idepth[keep], unc[keep])
    r = trk.optimize(start, (0.0, 0.0), gray=win["gray"][new], exposure_time=win["frame_exposure"][new])
    L = trk.read("pc_n", np.int32).size
    assert L == (levels or 5)
    wh = trk.read("levels_wh", np.int32).reshape(L, 2)
    pyr = T.build_pyramid(win["gray"][new], L)
    rows = T.project_to_reference(win["calib"], cams, cams[ref], win["pt_host"][keep], win["pt_xy"][keep], idepth[keep], unc[keep])
    pcs_o = T.make_coarse_depth(rows, [p[0] for p in T.build_pyramid(win["gray"][ref], L)])
    for l in range(L):
        assert pyr[l][0].shape == (wh[l, 1], wh[l, 0])
        assert np.array_equal(trk.read(f"grad{l}", np.float32).reshape(wh[l, 1], wh[l, 0], 4)[:, :, :3], pyr[l][1])
        pc = trk.read(f"pc{l}", np.float32).reshape(-1, 4)
        assert pc.shape == pcs_o[l].shape and np.array_equal(pc[:, [0, 1, 3]], pcs_o[l][:, [0, 1, 3]]) and np.abs(pc[:, 2] - pcs_o[l][:, 2]).max() <= 2 ** -22
    Rr, tr = cams[ref][:9].reshape(3, 3), cams[ref][9:]
    R0 = start[:9].reshape(3, 3) @ Rr.T
    pcs = [trk.read(f"pc{l}", np.float32).reshape(-1, 4) for l in range(L)]
    o = T.optimize(pcs, [p[1] for p in pyr], [T.level_K(win["calib"], l) for l in range(L)], (R0, start[9:] - R0 @ tr), ref_exp, (win["frame_exposure"][new], 0.0, 0.0),
                   params=dict(huber=7.0, cutoff=15.0, scale_trans=0.4, scale_b=500.0))
    assert r.isCorrect == o["isCorrect"] and r.iterations == o["iterations"]
    assert list(r.numTermsInE) == list(o["numTermsInE"]) and list(r.numSaturated) == list(o["numSaturated"])
    cam = np.concatenate([(o["R"] @ Rr).ravel(), o["R"] @ tr + o["t"]])
    assert np.abs(r.camera - cam).max() < 2e-5
    np.testing.assert_allclose(r.E, o["E"], rtol=1e-3)
