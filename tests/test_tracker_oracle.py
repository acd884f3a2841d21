"""The numpy restatement of DSOTracker (oracle/tracker_oracle.py) against the reference's golden vectors (tests/golden/track_*.cmlw, made by
oracle/make_golden.py tracker from the unmodified reference).  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from libcml_b200 import cmlw  # noqa: E402
import tracker_oracle as T  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load():
    return cmlw.load(os.path.join(GOLDEN, "track_window.cmlw")), cmlw.load(os.path.join(GOLDEN, "track_golden.cmlw"))


def rel_pose(ref_cam, cam):
    Rr, tr = ref_cam[:9].reshape(3, 3), ref_cam[9:]
    Rn, tn = cam[:9].reshape(3, 3), cam[9:]
    R = Rn @ Rr.T
    return R, tn - R @ tr


def compose(ref_cam, R, t):
    Rr, tr = ref_cam[:9].reshape(3, 3), ref_cam[9:]
    return np.concatenate([(R @ Rr).ravel(), R @ tr + t])


def test_coarse_depth_matches_reference():
    win, g = load()
    ref, new = int(win["track_ref"][0]), int(win["track_new"][0])
    L = g["trk_K"].shape[0]
    keep = win["pt_host"] != new
    rows = T.project_to_reference(win["calib"], win["frame_cam"], win["frame_cam"][ref], win["pt_host"][keep], win["pt_xy"][keep], win["pt_idepth"][keep],
                                  win["pt_uncertainty"][keep])
    pyr = T.build_pyramid(win["gray"][ref], L)
    pcs = T.make_coarse_depth(rows, [p[0] for p in pyr])
    for l in range(L):
        assert pyr[l][0].shape == (g["trk_levels_wh"][2 * l + 1], g["trk_levels_wh"][2 * l])
        assert np.abs(T.level_K(win["calib"], l) - g["trk_K"][l]).max() < 1e-12
        assert pcs[l].shape == g[f"trk_pc{l}"].shape
        assert np.array_equal(pcs[l][:, [0, 1, 3]], g[f"trk_pc{l}"][:, [0, 1, 3]])            # pixel and colour: exact
        assert np.abs(pcs[l][:, 2] - g[f"trk_pc{l}"][:, 2]).max() <= 2 ** -24                 # inverse depth: 1 ulp (summation order inside one pixel)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_optimize_matches_reference(case):
    win, g = load()
    ref, new = int(win["track_ref"][0]), int(win["track_new"][0])
    L = g["trk_K"].shape[0]
    pyr = T.build_pyramid(win["gray"][new], L)
    pcs = [g[f"trk_pc{l}"] for l in range(L)]
    R0, t0 = rel_pose(win["frame_cam"][ref], g[f"{case}_init_cam"])
    aff = g[f"{case}_new_affine"]
    out = T.optimize(pcs, [p[1] for p in pyr], [g["trk_K"][l] for l in range(L)], (R0, t0),
                     (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1]), (win["frame_exposure"][new], aff[0], aff[1]))
    assert out["isCorrect"] == bool(g[f"{case}_trk_isCorrect"][0])
    assert list(out["numTermsInE"]) == list(g[f"{case}_trk_numTermsInE"])
    assert list(out["numSaturated"]) == list(g[f"{case}_trk_numSaturated"])
    assert list(out["numRobust"]) == list(g[f"{case}_trk_numRobust"])
    np.testing.assert_allclose(out["E"], g[f"{case}_trk_E"], rtol=1e-4)
    if not out["isCorrect"]:
        return      # early exit: the reference leaves camera and exposure untouched (DSOTracker.cpp:65-69)
    cam = compose(win["frame_cam"][ref], out["R"], out["t"])
    assert np.abs(cam - g[f"{case}_trk_cam"]).max() < 1e-5
    np.testing.assert_allclose(out["exposure"][1:], g[f"{case}_trk_affine"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["levelCutoffRepeat"], g[f"{case}_trk_levelCutoffRepeat"])
    np.testing.assert_allclose(out["flow"], g[f"{case}_trk_flow"], rtol=1e-4)
    np.testing.assert_allclose(out["relAff"], g[f"{case}_trk_relAff"][:2], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["covariance"], g[f"{case}_trk_covariance"], rtol=1e-3)
    assert out["tooManySaturated"] == bool(g[f"{case}_trk_tooManySaturated"][0])
