"""CPU: the maintenance-decision restatement (oracle/maintenance_oracle.py) against the reference's golden (tests/golden/maint_*)."""
import os
import numpy as np
from parity_util import GOLDEN


def _load():
    from libcml_b200 import cmlw
    return cmlw.load(os.path.join(GOLDEN, "maint_window.cmlw")), cmlw.load(os.path.join(GOLDEN, "maint_golden.cmlw"))


def test_flagging_matches_reference():
    """The 6th addNewFrame of a maxFrames=5 window flags exactly the frame the reference flags (distance score, BA:649-700)."""
    import maintenance_oracle as M
    win, g = _load()
    N = win["frame_evalpt"].shape[0]
    flagged = np.zeros(N - 1, bool)
    # state at the time of the last addNewFrame: N-1 frames in the window, cameras = evaluation points, no points yet
    f = M.flag_frames(win["frame_evalpt"][:N - 1], np.arange(N - 1), win["frame_affine"][:N - 1, 0], win["frame_exposure"][:N - 1],
                      np.zeros(N - 1), np.zeros(N - 1), np.zeros(N - 1), np.zeros(N - 1), flagged, int(win["max_frames"][0]))
    assert np.array_equal(np.concatenate([f, [False]]), g["m0_frame_flagged"].astype(bool))
    # earlier addNewFrame calls (fewer frames than maxFrames) flag nothing
    for n in range(1, N - 1):
        assert not M.flag_frames(win["frame_evalpt"][:n], np.arange(n), win["frame_affine"][:n, 0], win["frame_exposure"][:n],
                                 np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n, bool), int(win["max_frames"][0])).any()


def test_try_marginalize_matches_reference():
    import maintenance_oracle as M
    win, g = _load()
    flagged = g["m0_frame_flagged"].astype(bool)
    alive = g["m0_pt_alive"].astype(bool)
    drop, marg = M.try_marginalize(win["pt_host"], g["m0_pt_idepth"], g["m0_pt_num_good"], g["m0_pt_idepth_hessian"], g["m0_pt_last0_state"], g["m0_pt_last1_state"],
                                   g["m0_res_point"], g["m0_res_target"], g["m0_res_state"], flagged, alive)
    assert np.array_equal(drop, g["m1_pt_outlier"].astype(bool))
    assert np.array_equal(marg, g["m1_pt_to_marginalize"].astype(bool))
    assert drop.sum() > 0 and marg.sum() > 0
    # counters: dropped points after tryMarginalize, marginalised points after marginalizePointsF
    nm, no = M.counters_after_removal(g["m0_frame_num_marginalized"], g["m0_frame_num_residuals_out"], g["m0_res_point"], g["m0_res_target"], drop, np.zeros_like(drop))
    assert np.array_equal(nm, g["m1_frame_num_marginalized"]) and np.array_equal(no, g["m1_frame_num_residuals_out"])
    nm, no = M.counters_after_removal(nm, no, g["m1_res_point"], g["m1_res_target"], marg, marg)
    assert np.array_equal(nm, g["m2_frame_num_marginalized"]) and np.array_equal(no, g["m2_frame_num_residuals_out"])
    assert np.array_equal(g["m2_pt_marginalized"].astype(bool), marg)
    assert list(g["m3_removed_frames"]) == list(np.nonzero(flagged)[0])
