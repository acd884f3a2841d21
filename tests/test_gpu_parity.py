"""GPU parity tests (pytest -m gpu): the CUDA path through the C ABI against
  (1) the committed golden vectors of the UNMODIFIED reference (tests/golden, small windows),
  (2) the numpy oracle on seeded windows,
  (3) the reference binary itself at BASELINE.json's sizes when oracle/_ref/cmlba_ref runs on this host,
  (4) size-independent properties at the full c2 size (determinism, energy descent, convergence to the truth).
Tolerance: 1e-4 relative (north_star, fp32 path) unless stated; integer / state work must match exactly."""
import os
import subprocess

import numpy as np
import pytest

from parity_util import (GOLDEN, ROOT, CTRL, DeviceView, decode_rj, load_golden, map_residuals, pose_errors, record, rel, solve_reference_system, split_sys, unpack_acc,
                         x_noise_floor)

pytestmark = pytest.mark.gpu
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "cmlba_ref")


X_FLOOR_FACTOR = 4      # x is gated at this multiple of the reference system's own measured noise floor (parity_util.x_noise_floor)


def _ba(**kw):
    from libcml_b200 import DSOBundleAdjustment
    return DSOBundleAdjustment(device=0, **kw)


def _check_lin(ba, dv, g, stage, full, tol=1e-4):
    """tol = 1e-4 where the inputs are identical to the reference's (first linearization).  After a GN step the
    states differ by the conditioning noise floor of x (DESIGN.md "conditioning"), so per-residual values of later
    linearizations are compared at 1e-3 while total energies and the state machine stay at 1e-4 / exact."""
    pre = stage + "_"
    m = dv.map_to(g, pre)
    ns = ba.read("res_new_state", np.uint8)[m]
    assert np.array_equal(ns, g[pre + "res_new_state"]), f"{stage}: residual state machine differs"
    assert rel(ba.read("res_new_energy", np.float32)[m], g[pre + "res_new_energy"]) < tol
    assert rel(ba.read("res_new_energy_wo", np.float32)[m], g[pre + "res_new_energy_wo"]) < tol
    assert rel(dv.frames()["energy_th"], g[pre + "frame_energy_th"]) < tol
    if full and pre + "rJ_resF" in g:
        J = decode_rj(ba.read("rj", np.float32), ba.read("dbg", np.float32))
        ok = g[pre + "res_new_state"] == 0
        for k in ["resF", "Jpdxi", "Jpdc", "Jpdd", "JIdx", "JabF", "JIdx2", "JabJIdx", "Jab2"]:
            assert rel(J[k][m][ok], g[pre + "rJ_" + k][ok]) < tol, (stage, k)
        okc = g[pre + "res_new_state"] != 1
        assert rel(ba.read("res_center", np.float32).reshape(-1, 3)[m][okc], g[pre + "res_center"][okc]) < 1e-6


@pytest.mark.parametrize("staging", ["direct_taps", "tma_tiles"])
@pytest.mark.parametrize("name", ["tiny", "tiny_affine"])
def test_stages_against_reference_golden(name, staging, monkeypatch):
    # small windows are sparse (a few residuals per 64x32 tile): the engine reads their taps directly; CMLBA_FORCE_TMA=1 runs the same window through the
    # TMA-staged tile path that the full-size windows take
    if staging == "tma_tiles":
        monkeypatch.setenv("CMLBA_FORCE_TMA", "1")
    win, g = load_golden(name)
    N = win["frame_evalpt"].shape[0]; n = 8 * N + 4
    ba = _ba()
    ba.enableDebugDump()
    cams = ba.loadWindow(win)
    ba.prepare(cams)
    dv = DeviceView(ba, win)
    fr = dv.frames()
    assert rel(fr["state"], g["pre_frame_state"]) < 1e-8
    assert rel(fr["preR"], g["pre_frame_pre_w2c"][:, :9]) < 1e-12 and rel(fr["pret"], g["pre_frame_pre_w2c"][:, 9:]) < 1e-12
    AH = ba.read("AH", np.float64).reshape(N, N, 8, 8); AT = ba.read("AT", np.float64).reshape(N, N, 8, 8)
    assert rel(AH, g["pre_ad_host"].reshape(N, N, 8, 8).transpose(1, 0, 2, 3)) < 1e-12     # reference index h + N*t
    assert rel(AT, g["pre_ad_target"].reshape(N, N, 8, 8).transpose(1, 0, 2, 3)) < 1e-12
    assert rel(dv.point_array("pt_colors", np.float32, 8), g["pre_pt_colors"]) == 0        # integer pixel reads: exact
    assert rel(dv.point_array("pt_weights", np.float32, 8), g["pre_pt_weights"]) < 1e-6
    # first linearization at the reference's initial state
    E0 = ba.linearizeAll(False)
    assert abs(E0 - g["lin0_energy"][0]) / g["lin0_energy"][0] < 1e-5
    _check_lin(ba, dv, g, "lin0", True)
    ba.applyActiveRes()
    m = dv.map_to(g, "app0_")
    assert np.array_equal(ba.read("res_good", np.uint8)[m], g["app0_res_good"].astype(np.uint8))
    assert np.array_equal(ba.read("res_state", np.uint8)[m], g["app0_res_state"].astype(np.uint8))
    T = ba.read("T", np.float32).reshape(dv.P, N, 16)
    Tc = np.zeros_like(T); Tc[dv.pt_order] = T
    assert rel(Tc[g["app0_res_point"], g["app0_res_target"], :8], g["app0_res_JpJdF"]) < 1e-5
    # first system: accumulators, Schur, stitched matrices at identical inputs
    ba.solveSystem(0)
    assert rel(unpack_acc(ba.read("acc", np.float64), N), g["sol0_acc_active"]) < 1e-5
    s = split_sys(ba.read("sys", np.float64), n)
    assert rel(s["HA"], g["sol0_HA_top"]) < 1e-5 and rel(s["bA"], g["sol0_bA_top"][:, 0]) < 1e-4
    assert rel(s["HS"], g["sol0_H_sc"]) < 1e-5 and rel(s["bS"], g["sol0_b_sc"][:, 0]) < 1e-4
    for k in ["pt_Hdd", "pt_bd", "pt_HdiF", "pt_bdSumF", "pt_idepth_hessian"]:
        assert rel(dv.point_array(k, np.float32), g["sol0_" + k]) < 1e-4, k
    assert rel(dv.point_array("pt_Hcd", np.float32, 4), g["sol0_pt_Hcd"]) < 1e-4
    # x: backward error in the REFERENCE's system (its forward error floor is ~1e-3, see DESIGN.md "conditioning")
    x = ba.read("x", np.float64)
    H, b = solve_reference_system(g, "sol0_", N)
    Hl = np.tril(H[4:, 4:]) + np.tril(H[4:, 4:], -1).T
    assert np.abs(Hl @ x[4:] - b[4:]).max() / np.abs(b[4:]).max() < 1e-4
    # forward error of x against the MEASURED floor of the reference's own system (cond(H) ~ 1e11: solving the upper instead of
    # the lower triangle of the reference's H, or one fp32 ulp of relative noise on its entries, already moves x by ~1e-3)
    tri, ulp, cond = x_noise_floor(g, "sol0_", N)
    ex = rel(x, g["sol0_x"]); es = rel(dv.point_array("pt_step", np.float64), g["sol0_pt_step"])
    record(f"stages[{name},{staging}]", x_rel_err=ex, x_floor_triangle=tri, x_floor_fp32_ulp=ulp, cond_H=cond, pt_step_rel_err=es,
           x_backward_err=np.abs(Hl @ x[4:] - b[4:]).max() / np.abs(b[4:]).max())
    assert ex < X_FLOOR_FACTOR * max(tri, ulp), (ex, tri, ulp)
    assert es < X_FLOOR_FACTOR * max(tri, ulp)
    assert rel(dv.point_array("pt_idepth", np.float64), g["step0_pt_idepth"]) < 1e-4
    assert ba.doStepFromBackup() == bool(g["step0_canbreak"][0])
    # remaining iterations: state machine must stay identical, energies within 1e-4
    it = 1
    E = ba.linearizeAll(False)
    assert abs(E - g["lin1_energy"][0]) / g["lin1_energy"][0] < 1e-4
    _check_lin(ba, dv, g, "lin1", True, tol=1e-3)
    ba.applyActiveRes()
    while f"sol{it}_x" in g:
        ba.solveSystem(it)
        E = ba.linearizeAll(False)
        assert abs(E - g[f"lin{it + 1}_energy"][0]) / g[f"lin{it + 1}_energy"][0] < 1e-4
        _check_lin(ba, dv, g, f"lin{it + 1}", False, tol=1e-3)
        ba.applyActiveRes()
        it += 1
    E = ba.linearizeAll(True)
    assert abs(E - g["fin_energy"][0]) / g["fin_energy"][0] < 1e-4
    assert np.array_equal(dv.point_array("pt_num_good", np.int32), g["fin_pt_num_good"])
    assert rel(dv.point_array("pt_max_rel_baseline", np.float32), g["fin_pt_max_rel_baseline"]) < 1e-4
    ba.close()


@pytest.mark.parametrize("staging", ["direct_taps", "tma_tiles"])
@pytest.mark.parametrize("name", ["tiny", "tiny_affine"])
def test_run_against_reference_golden(name, staging, monkeypatch):
    """The public run(): poses / affine / inverse depths / surviving residuals / outliers vs the reference's run()."""
    if staging == "tma_tiles":
        monkeypatch.setenv("CMLBA_FORCE_TMA", "1")
    win, g = load_golden(name)
    ba = _ba()
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=int(win["iterations"][0]))
    r = ba.last_result
    assert r.iterations_done == int(g["iterations_done"][0])
    assert abs(r.energy_first - g["lin0_energy"][0]) / g["lin0_energy"][0] < 1e-5
    assert abs(r.energy_last - g["fin_energy"][0]) / g["fin_energy"][0] < 1e-4
    fr = ba.getFrames(); pts = ba.getPoints(); rs = ba.getResiduals()
    er, et, ets = pose_errors(fr["world_to_cam"], g["fin_frame_pre_w2c"])
    record(f"run[{name},{staging}]", rot_abs_err=er, trans_rel_err=et, trans_rel_err_scale_removed=ets, affine_abs_err=np.abs(fr["affine"] - g["fin_frame_affine"]).max(),
           idepth_rel_err=rel(pts["idepth"], g["fin_pt_idepth"][pts["id"]]), energy_rel_err=abs(r.energy_last - g["fin_energy"][0]) / g["fin_energy"][0])
    # poses within 1e-4 (north_star): rotations absolutely, translations relative to ||t|| once the unobservable common scale of the
    # window is factored out; with it (raw) the conditioning noise of the scale direction is allowed 3e-4 (measured: profiles/r02_parity_report.txt)
    assert er < 1e-4 and ets < 1e-4 and et < 3e-4
    assert rel(fr["evalpt"], g["fin_frame_evalpt"]) < 1e-4
    assert np.abs(fr["affine"] - g["fin_frame_affine"]).max() < 1e-4 * max(1.0, np.abs(g["fin_frame_affine"]).max())
    assert rel(fr["energy_th"], g["fin_frame_energy_th"]) < 1e-4
    alive = g["fin_pt_alive"].astype(bool)
    assert np.array_equal(np.sort(pts["id"]), np.nonzero(alive)[0])
    assert rel(pts["idepth"], g["fin_pt_idepth"][pts["id"]]) < 1e-3
    assert rel(pts["uncertainty"], g["fin_pt_uncertainty"][pts["id"]]) < 1e-3
    assert np.array_equal(pts["num_good_residuals"], g["fin_pt_num_good"][pts["id"]])
    assert np.array_equal(np.sort(pts["id"][pts["good_for_tracking"] != 0]), g["fin_good_points_for_tracking"])
    assert set(zip(rs["point_id"].tolist(), rs["target_frame_id"].tolist())) == set(zip(g["fin_alive_res_point"].tolist(), g["fin_alive_res_target"].tolist()))
    assert np.array_equal(np.sort(ba.getOutliers()), np.nonzero(g["fin_pt_outlier"])[0])
    ba.close()


def test_against_numpy_oracle_seeded():
    """Different seed / size than the fixtures: CUDA path vs the oracle restatement on the same inputs."""
    import ba_oracle as O
    from libcml_b200 import synth
    win = synth.make_window(W=320, H=240, N=5, pts_per_kf=300, iterations=4, affine=True, seed=77)
    ba = _ba()
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=4)
    ow = O.Window(win)
    assert O.run(ow)
    fr = ba.getFrames(); pts = ba.getPoints()
    ref = np.stack([np.concatenate([R.ravel(), t]) for R, t in ow.pre_w2c])
    assert rel(fr["world_to_cam"], ref) < 1e-4
    assert rel(pts["idepth"], ow.idepth[pts["id"]]) < 1e-3
    assert abs(ba.last_result.energy_last - ow.fin_energy) / ow.fin_energy < 1e-4
    assert abs(ba.last_result.num_dropped - int((~ow.res_alive).sum())) <= 2
    ba.close()


def _ref_runs():
    if not os.path.exists(REF_BIN):
        return False
    try:
        return subprocess.run([REF_BIN], capture_output=True, timeout=20).returncode == 2   # prints usage
    except Exception:
        return False


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_against_reference_binary_full_size(cfg, tmp_path):
    """BASELINE.json configs[0..3] at full size (c4: 16 KF, 720 000 residuals) against the reference binary run HERE on the host CPU."""
    if not _ref_runs():
        pytest.skip("oracle/_ref/cmlba_ref not runnable on this host")
    from libcml_b200 import cmlw, synth
    win = synth.make_config(cfg)
    wp = str(tmp_path / "w.cmlw"); op = str(tmp_path / "o.cmlw")
    cmlw.save(wp, {k: v for k, v in win.items() if k != "grad"})
    subprocess.run([REF_BIN, "--window", wp, "--mode", "run", "--out", op], check=True, capture_output=True, timeout=600)
    g = cmlw.load(op)
    ba = _ba()
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=int(win["iterations"][0])) == bool(g["fin_ok"][0])
    fr = ba.getFrames(); pts = ba.getPoints(); rs = ba.getResiduals()
    er, et, ets = pose_errors(fr["world_to_cam"], g["fin_frame_pre_w2c"])
    assert er < 1e-4 and ets < 1e-4 and et < 3e-4
    assert np.abs(fr["affine"] - g["fin_frame_affine"]).max() < 1e-4 * max(1.0, np.abs(g["fin_frame_affine"]).max())
    assert rel(pts["idepth"], g["fin_pt_idepth"][pts["id"]]) < 1e-3
    mine = set(zip(rs["point_id"].tolist(), rs["target_frame_id"].tolist()))
    theirs = set(zip(g["fin_alive_res_point"].tolist(), g["fin_alive_res_target"].tolist()))
    assert len(mine ^ theirs) <= max(1, len(theirs) // 1000), f"{len(mine ^ theirs)} residual state flips of {len(theirs)}"   # >= 99.9 % agreement
    ba.close()


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3", "c4"])
def test_against_reference_summary_full_size(cfg):
    """BASELINE.json configs[0..3] at FULL size against tests/golden/fullsize_summary.cmlw: a compact summary (final poses, affine,
    summed energy, accepted steps, every 16th inverse depth, the surviving (point, target) set as count / per-target histogram / sha1)
    of the UNMODIFIED reference's run() on the same seeded window (oracle/make_golden.py summaries).  Unlike the test above this one
    does not need the reference binary on the GPU box.  Measured errors go to gpurun_out/parity_report.jsonl."""
    import hashlib
    from libcml_b200 import cmlw, synth
    S = cmlw.load(os.path.join(GOLDEN, "fullsize_summary.cmlw"))
    win = synth.make_config(cfg)
    ba = _ba()
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=int(win["iterations"][0])) == bool(S[f"{cfg}_ok"][0])
    fr = ba.getFrames(); pts = ba.getPoints(); rs = ba.getResiduals()
    er, et, ets = pose_errors(fr["world_to_cam"], S[f"{cfg}_w2c"])
    ea = float(np.abs(fr["affine"] - S[f"{cfg}_affine"]).max())
    P = win["pt_host"].shape[0]
    idepth = np.full(P, np.nan); idepth[pts["id"]] = pts["idepth"]
    alive16 = S[f"{cfg}_alive16"].astype(bool)
    mine_alive16 = ~np.isnan(idepth[::16])
    both = alive16 & mine_alive16
    ed = rel(idepth[::16][both], S[f"{cfg}_idepth16"][both])
    key = np.sort(rs["point_id"].astype(np.int64) * 64 + rs["target_frame_id"].astype(np.int64))
    per_target = np.bincount(rs["target_frame_id"].astype(np.int64), minlength=S[f"{cfg}_res_per_target"].size)
    flips_lb = int(np.abs(per_target - S[f"{cfg}_res_per_target"]).sum())            # lower bound on the symmetric difference
    same_set = bool(np.array_equal(np.frombuffer(hashlib.sha1(key.tobytes()).digest(), dtype=np.uint8), S[f"{cfg}_res_digest"]))
    n_ref = int(S[f"{cfg}_n_alive_res"][0])
    energy = float(np.asarray(rs["energy"], np.float64).sum())
    record(f"fullsize_summary[{cfg}]", rot_abs_err=er, trans_rel_err=et, trans_rel_err_scale_removed=ets, affine_abs_err=ea, idepth16_rel_err=ed, n_alive_res=key.size, n_alive_res_ref=n_ref,
           per_target_count_diff=flips_lb, identical_residual_set=float(same_set), alive16_mismatch=int((alive16 != mine_alive16).sum()),
           iterations_done=ba.last_result.iterations_done, accepted_ref=int(S[f"{cfg}_accepted"][0]),
           energy_sum_rel_err=abs(energy - float(S[f"{cfg}_energy"][0])) / float(S[f"{cfg}_energy"][0]))
    # gates ~4x above the measured errors (profiles/r02_parity_report.txt: rotations <= 7.5e-8, translations <= 2.5e-5 raw / 8e-7 without the
    # common scale, inverse depths <= 2.5e-5, energies <= 8e-6, surviving sets identical on all four configurations)
    assert er < 1e-6 and ets < 1e-5 and et < 1e-4
    assert ea < 1e-5 * max(1.0, np.abs(S[f"{cfg}_affine"]).max())
    assert ed < 1e-4
    assert abs(key.size - n_ref) <= n_ref // 10000 and flips_lb <= n_ref // 10000          # >= 99.99 % of the surviving set (measured: identical, sha1 equal)
    assert (alive16 != mine_alive16).sum() == 0
    assert len(pts["id"]) == int(S[f"{cfg}_n_alive_pts"][0])
    assert abs(energy - float(S[f"{cfg}_energy"][0])) / float(S[f"{cfg}_energy"][0]) < 1e-4
    ba.close()


def test_properties_full_size_c2():
    """Size-independent properties at BASELINE.json's headline size (8 KF x 2000 pts, 112 000 residuals)."""
    from libcml_b200 import synth
    win = synth.make_config("c2")
    ba = _ba()
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=6)
    r1 = ba.last_result
    assert r1.num_residuals == 8 * 2000 * 7
    assert r1.energy_last < r1.energy_first                      # GN descends
    f1 = ba.getFrames(); p1 = ba.getPoints()
    # bitwise reproducible run to run (fixed-order reductions)
    ba2 = _ba()
    cams = ba2.loadWindow(win)
    assert ba2.run(cams, iterations=6)
    f2 = ba2.getFrames(); p2 = ba2.getPoints()
    assert np.array_equal(f1["world_to_cam"], f2["world_to_cam"]) and np.array_equal(p1["idepth"], p2["idepth"])
    assert ba2.last_result.energy_last == r1.energy_last
    # converges towards the generating scene: relative poses to KF0 and inverse depths closer to the truth than the start
    def rel_t(w2c):
        R = w2c[:, :9].reshape(-1, 3, 3); t = w2c[:, 9:]
        R0, t0 = R[0], t[0]
        return np.stack([t[i] - R[i] @ R0.T @ t0 for i in range(len(t))])
    e_start = np.abs(rel_t(win["frame_cam"]) - rel_t(win["truth_frame"])).max()
    e_end = np.abs(rel_t(f1["world_to_cam"]) - rel_t(win["truth_frame"])).max()
    assert e_end < 0.5 * e_start
    tid = win["truth_idepth"][p1["id"]]
    assert np.median(np.abs(p1["idepth"] - tid) / tid) < np.median(np.abs(win["pt_idepth"][p1["id"]] - tid) / tid)
    ba.close(); ba2.close()


def test_empty_and_edge_windows():
    """Edge cases the reference guards: no points (run() returns false, BA:759-762), points that project out of every
    other frame (all residuals OOB -> points become outliers), bad arguments."""
    from libcml_b200 import CmlbaError, synth
    win = synth.make_window(W=160, H=120, N=3, pts_per_kf=20, iterations=2, seed=5)
    ba = _ba()
    W, H = 160, 120
    ba.setCalibration(*[float(v) for v in win["calib"]], W, H)
    for i in range(3):
        ba.addNewFrame(i, win["frame_evalpt"][i], 0.0, 0.0, 1.0, win["grad"][i])
    with pytest.raises(CmlbaError):
        ba.run(win["frame_cam"])                                   # "No points..."
    with pytest.raises(CmlbaError):
        ba.addNewFrame(1, win["frame_evalpt"][1], 0.0, 0.0, 1.0, win["grad"][1])   # ids must increase
    with pytest.raises(CmlbaError):
        ba.addPoints([0], [99], [[50, 50]], [0.5])                 # unknown host frame
    with pytest.raises(CmlbaError):
        ba.addPoints([0], [0], [[1, 1]], [0.5])                    # too close to the border
    # a huge inverse depth throws every projection far outside -> all residuals OOB -> outlier points
    ba.addPoints([0, 1, 2], [0, 0, 0], [[50, 50], [60, 60], [70, 70]], [500.0, 500.0, 500.0])
    ba.addPoints(np.arange(10, 10 + win["pt_host"].size), win["pt_host"], win["pt_xy"], win["pt_idepth"])
    assert ba.run(win["frame_cam"], iterations=2)
    assert set(ba.getOutliers().tolist()) >= {0, 1, 2}
    assert not ({0, 1, 2} & set(ba.getPoints()["id"].tolist()))
    # a second run() on the surviving window keeps working (dropped residuals are gone, like the reference)
    assert ba.run(None, iterations=2)
    ba.close()


def test_step_rejection_against_reference_golden():
    """forceAccept = false (BA:843-877): accepted / rejected step counts and the final state vs the reference's run()."""
    from libcml_b200 import cmlw, synth
    from parity_util import GOLDEN
    win = cmlw.load(os.path.join(GOLDEN, "reject_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "reject_golden.cmlw"))
    win["grad"] = np.stack([synth.gradient_image(win["gray"][i]) for i in range(win["gray"].shape[0])])
    from libcml_b200 import DSOBundleAdjustment
    ba = DSOBundleAdjustment(device=0, force_accept=0)
    cams = ba.loadWindow(win)
    assert ba.run(cams, iterations=int(win["iterations"][0])) == bool(g["fin_ok"][0])
    r = ba.last_result
    assert r.iterations_done - r.num_rejected == int(g["accepted_count"][0]) and r.num_rejected > 0
    fr = ba.getFrames(); pts = ba.getPoints(); rs = ba.getResiduals()
    assert rel(fr["world_to_cam"], g["fin_frame_pre_w2c"]) < 1e-4
    assert np.abs(fr["affine"] - g["fin_frame_affine"]).max() < 1e-4 * max(1.0, np.abs(g["fin_frame_affine"]).max())
    assert rel(fr["energy_th"], g["fin_frame_energy_th"]) < 1e-4
    assert np.array_equal(np.sort(pts["id"]), np.nonzero(g["fin_pt_alive"])[0])
    assert rel(pts["idepth"], g["fin_pt_idepth"][pts["id"]]) < 1e-3
    assert np.array_equal(pts["num_good_residuals"], g["fin_pt_num_good"][pts["id"]])
    assert set(zip(rs["point_id"].tolist(), rs["target_frame_id"].tolist())) == set(zip(g["fin_alive_res_point"].tolist(), g["fin_alive_res_target"].tolist()))
    ba.close()
