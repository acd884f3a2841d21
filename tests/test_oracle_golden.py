"""CPU tests: the numpy restatement (oracle/ba_oracle.py) against golden vectors produced by the UNMODIFIED
reference (tests/golden/*_stages.cmlw, made by oracle/make_golden.py from oracle/_ref/cmlba_ref).
This is what pins the oracle."""
import numpy as np
import pytest

import ba_oracle as O
from parity_util import load_golden, rel, map_residuals

CASES = ["tiny", "tiny_affine"]


def _w2c(w):
    return np.stack([np.concatenate([R.ravel(), t]) for R, t in w.pre_w2c])


@pytest.mark.parametrize("name", CASES)
def test_prologue(name):
    win, g = load_golden(name)
    w = O.Window(win)
    O.compute_adjoints(w); O.compute_delta(w)
    N = w.N
    assert rel(w.state, g["pre_frame_state"]) < 1e-8
    assert rel(_w2c(w), g["pre_frame_pre_w2c"]) < 1e-12
    assert rel(w.AH, g["pre_ad_host"]) < 1e-12
    assert rel(w.AT, g["pre_ad_target"]) < 1e-12
    assert rel(w.prior, g["pre_frame_prior"]) == 0
    assert rel(np.stack(w.ns_pose), g["pre_frame_ns_pose"]) < 1e-9
    assert rel(np.stack(w.ns_scale), g["pre_frame_ns_scale"]) < 1e-9
    assert rel(w.colors, g["pre_pt_colors"]) == 0        # integer-pixel reads: bit exact
    assert rel(w.weights, g["pre_pt_weights"]) < 1e-7


@pytest.mark.parametrize("name", CASES)
def test_first_linearization_and_system(name):
    """linearize -> applyRes -> accumulators -> stitched H/b -> Schur, all at the reference's initial state."""
    win, g = load_golden(name)
    w = O.Window(win)
    O.compute_adjoints(w); O.compute_delta(w)
    E = O.linearize_all(w)
    m = map_residuals(w.res_point, w.res_target, g["lin0_res_point"], g["lin0_res_target"])
    assert np.array_equal(w.res_new_state[m], g["lin0_res_new_state"])           # state machine: exact
    assert abs(E - g["lin0_energy"][0]) / g["lin0_energy"][0] < 1e-5
    assert rel(w.res_new_energy[m], g["lin0_res_new_energy"]) < 1e-4             # tolerance of north_star: 1e-4 relative, fp32
    inv = np.zeros(w.R, int); inv[w.rJ["idx"]] = np.arange(w.rJ["idx"].size)
    ok = g["lin0_res_new_state"] != 1
    for k in ["resF", "Jpdxi", "Jpdc", "Jpdd", "JIdx", "JabF", "JIdx2", "JabJIdx", "Jab2"]:
        assert rel(w.rJ[k][inv[m]][ok], g["lin0_rJ_" + k][ok]) < 1e-4, k
    assert rel(w.frame_energy_th, g["lin0_frame_energy_th"]) < 1e-5              # 0.7-quantile threshold
    O.apply_active_res(w)
    assert np.array_equal(w.res_good[m], g["app0_res_good"].astype(bool))
    assert rel(w.JpJdF[m], g["app0_res_JpJdF"]) < 1e-5
    O.backup_state(w)
    assert O.solve_system(w, 0, w.p["fixed_lambda"])
    assert rel(w.acc, g["sol0_acc_active"]) < 1e-5
    assert np.array_equal(w.acc_num, g["sol0_acc_active_num"])
    s = w.sys
    assert rel(s["HA"][4:, 4:], g["sol0_HA_top"][4:, 4:]) < 1e-5
    assert rel(s["bA"], g["sol0_bA_top"][:, 0]) < 1e-4
    assert rel(s["Hsc"], g["sol0_H_sc"]) < 1e-5
    assert rel(s["bsc"], g["sol0_b_sc"][:, 0]) < 1e-4
    assert rel(w.accD, g["sol0_acc_D"]) < 1e-5 and rel(w.accE, g["sol0_acc_E"]) < 1e-5 and rel(w.accEB, g["sol0_acc_EB"]) < 1e-4
    for mine, gk in [(w.Hdd, "pt_Hdd"), (w.bd, "pt_bd"), (w.Hcd, "pt_Hcd"), (w.HdiF, "pt_HdiF"), (w.bdSumF, "pt_bdSumF")]:
        assert rel(mine, g["sol0_" + gk]) < 1e-4, gk
    assert rel(O.nullspaces(w), g["sol0_nullspaces"]) < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_solve_backward_error(name):
    """x is ill-conditioned w.r.t. fp32 accumulation noise (cond ~1e5: the reference's own lower- vs upper-triangle
    solve differs by ~4e-4), so the gate on x is the residual of the REFERENCE's system, plus a loose forward bound."""
    win, g = load_golden(name)
    w = O.Window(win)
    O.compute_adjoints(w); O.compute_delta(w)
    O.linearize_all(w); O.apply_active_res(w); O.backup_state(w)
    O.solve_system(w, 0, w.p["fixed_lambda"])
    lam = float(np.float32(1e-5))
    H = g["sol0_HL_top"] + g["sol0_HA_top"]
    b = (g["sol0_bL_top"] + g["sol0_bM_top"] + g["sol0_bA_top"] - g["sol0_b_sc"])[:, 0]
    H = H.copy(); H[np.diag_indices_from(H)] *= 1 + lam; H -= g["sol0_H_sc"] / (1 + lam)
    Hl = np.tril(H[4:, 4:]) + np.tril(H[4:, 4:], -1).T       # Eigen's LDLT reads the lower triangle only
    x = w.sys["x"]
    assert np.abs(Hl @ x[4:] - b[4:]).max() / np.abs(b[4:]).max() < 1e-4
    assert rel(x, g["sol0_x"]) < 2e-2
    # the golden x solves its own lower-triangle system to machine precision (pins the solver semantics)
    assert np.abs(Hl @ g["sol0_x"][4:] - b[4:]).max() / np.abs(b[4:]).max() < 1e-10


@pytest.mark.parametrize("name", CASES)
def test_full_run(name):
    win, g = load_golden(name)
    w = O.Window(win)
    assert O.run(w)
    assert w.iterations_done == int(g["iterations_done"][0])
    assert abs(w.fin_energy - g["fin_energy"][0]) / g["fin_energy"][0] < 1e-4
    assert rel(_w2c(w), g["fin_frame_pre_w2c"]) < 1e-4                           # poses within 1e-4 (north_star)
    assert np.abs(np.array([w.aff(i) for i in range(w.N)]) - g["fin_frame_affine"]).max() < 1e-4 * max(1.0, np.abs(g["fin_frame_affine"]).max())
    assert rel(w.idepth, g["fin_pt_idepth"]) < 1e-3
    alive = set(zip(g["fin_alive_res_point"].tolist(), g["fin_alive_res_target"].tolist()))
    mine = set(zip(w.res_point[w.res_alive].tolist(), w.res_target[w.res_alive].tolist()))
    assert len(alive ^ mine) <= max(1, len(alive) // 1000)                       # >= 99.9 % state agreement
    assert int((w.num_good_res != g["fin_pt_num_good"]).sum()) <= max(1, w.P // 1000)


def test_se3_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(50):
        xi = rng.standard_normal(6) * np.array([1, 1, 1, 0.5, 0.5, 0.5])
        R, t = O.se3_exp(xi)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
        assert np.allclose(O.se3_log(R, t), xi, atol=1e-9)
    assert np.allclose(O.se3_log(*O.se3_exp(np.zeros(6))), 0)


def test_gradient_image_definition():
    """CaptureImageGenerator's level-0 derivative image (image/Array2D.h:288-331): central differences * 0.5, zero border.
    oracle/make_golden.py asserts bit equality with the reference's own image when the fixtures are generated."""
    from libcml_b200 import synth
    g = np.arange(30, dtype=np.float32).reshape(5, 6) ** 2
    d = synth.gradient_image(g)
    assert d.shape == (5, 6, 3) and np.all(d[0] == 0) and np.all(d[:, 0] == 0) and np.all(d[-1] == 0) and np.all(d[:, -1] == 0)
    assert d[2, 2, 0] == g[2, 2] and d[2, 2, 1] == (g[2, 3] - g[2, 1]) * 0.5 and d[2, 2, 2] == (g[3, 2] - g[1, 2]) * 0.5


def test_step_rejection_matches_reference():
    """forceAccept = false: the restatement takes the reference's accept / reject decisions (BA:843-877, calcLEnergy BA:2118-2208)."""
    import os
    from libcml_b200 import cmlw
    from parity_util import GOLDEN
    from libcml_b200 import synth
    win = cmlw.load(os.path.join(GOLDEN, "reject_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "reject_golden.cmlw"))
    win["grad"] = np.stack([synth.gradient_image(win["gray"][i]) for i in range(win["gray"].shape[0])])
    w = O.Window(win)
    assert O.run(w) == bool(g["fin_ok"][0])
    assert sum(w.accepted) == int(g["accepted_count"][0]) and 0 in w.accepted
    mine = np.stack([np.concatenate([R.ravel(), t]) for R, t in w.pre_w2c])
    assert rel(mine, g["fin_frame_pre_w2c"]) < 1e-4
    assert rel(w.idepth, g["fin_pt_idepth"]) < 1e-3


def test_marginalisation_prior_restatement_matches_reference():
    """marginalize_points_prior (tryMarginalize's re-linearization + marginalizePointsF, BA:2289-2300, 2466-2513) against H_M, b_M of the
    reference (disableMarginalization = false) after three run() calls."""
    import os
    from libcml_b200 import cmlw, synth
    from parity_util import GOLDEN
    win = cmlw.load(os.path.join(GOLDEN, "maint_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "maintp_golden.cmlw"))
    win["grad"] = np.stack([synth.gradient_image(win["gray"][i]) for i in range(win["gray"].shape[0])])
    w = O.Window(win)
    for _ in range(int(win["runs"][0])):
        assert O.run(w)
    dH, db = O.marginalize_points_prior(w, g["m2_pt_marginalized"].astype(bool))
    assert np.linalg.norm(dH - g["m2_HM"]) / np.linalg.norm(g["m2_HM"]) < 1e-5
    # b_M cancels heavily near convergence (J^T resF against J^T J delta): 1e-6 state differences after three runs show up at ~1e-3
    assert np.linalg.norm(db.ravel() - g["m2_bM"][:, 0]) / np.linalg.norm(g["m2_bM"]) < 3e-3


@pytest.mark.parametrize("name", ["tiny", "tiny_affine"])
def test_forward_error_floor_of_the_reference_system(name):
    """The experiment behind the x gate of tests/test_gpu_parity.py: the reference's own Gauss-Newton system is conditioned at ~1e11, so
    solving the UPPER instead of the lower triangle of ITS H (they differ by 1e-13 relative), or one fp32 ulp of noise on the entries (the
    reference accumulates H in fp32), moves x by ~1e-3.  A forward gate on x below that floor would test the noise, not the kernels."""
    from parity_util import load_golden, x_noise_floor
    win, g = load_golden(name)
    tri, ulp, cond = x_noise_floor(g, "sol0_", win["frame_evalpt"].shape[0])
    assert cond > 1e10
    assert 2e-4 < tri < 5e-3 and 2e-4 < ulp < 1e-2


def test_oracle_against_reference_summary_c1():
    """The numpy restatement at BASELINE.json's configs[0] (the reference's own CPU-runnable case: 2 KF, 200 points, 1 GN iteration at 640x480) against
    the compact summary of the UNMODIFIED reference's run() on the same seeded window (tests/golden/fullsize_summary.cmlw)."""
    import os
    from libcml_b200 import cmlw, synth
    from parity_util import GOLDEN, pose_errors
    S = cmlw.load(os.path.join(GOLDEN, "fullsize_summary.cmlw"))
    ow = O.Window(synth.make_config("c1"))
    assert O.run(ow) == bool(S["c1_ok"][0])
    w2c = np.stack([np.concatenate([R.ravel(), t]) for R, t in ow.pre_w2c])
    er, et, ets = pose_errors(w2c, S["c1_w2c"])
    assert er < 1e-6 and et < 1e-4 and ets < 1e-5
    alive = S["c1_alive16"].astype(bool)
    assert rel(ow.idepth[::16][alive], S["c1_idepth16"][alive]) < 1e-4
    assert int(ow.res_alive.sum()) == int(S["c1_n_alive_res"][0])
