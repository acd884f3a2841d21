"""Generates the committed golden fixtures under tests/golden/ -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (oracle/_ref/cmlba_ref, built by oracle/Makefile from /root/reference) on
small synthetic windows and stores window + golden vectors.  Run in the build container (needs
/root/reference only through the prebuilt binary):
    python oracle/make_golden.py
The big arrays that are recomputable (gradient images) are checked for equality and then dropped.
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import cmlw, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "cmlba_ref")
OUT = os.path.join(ROOT, "tests", "golden")


def run_ref(window_path, mode, out_path):
    r = subprocess.run([REF, "--window", window_path, "--mode", mode, "--out", out_path], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"cmlba_ref failed: {r.stderr[-2000:]}")


def main():
    os.makedirs(OUT, exist_ok=True)
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    for name in ("tiny", "tiny_affine"):
        win = synth.make_config(name)
        full = os.path.join(tmp, f"{name}.cmlw")
        cmlw.save(full, win)
        run_ref(full, "stages", os.path.join(tmp, f"{name}_stages.cmlw"))
        run_ref(full, "run", os.path.join(tmp, f"{name}_run.cmlw"))
        st = cmlw.load(os.path.join(tmp, f"{name}_stages.cmlw"))
        rn = cmlw.load(os.path.join(tmp, f"{name}_run.cmlw"))
        # the reference's own level-0 derivative image must equal our restatement bit for bit
        assert np.array_equal(st["grad_images"], win["grad"]), "gradient image restatement differs from the reference"
        del st["grad_images"]
        # "stages" (replayed through the protected members) must equal the public run()
        for k in ("fin_frame_state", "fin_frame_pre_w2c", "fin_pt_idepth", "fin_pt_num_good", "fin_frame_energy_th"):
            assert np.array_equal(st[k], rn[k]), f"stages != run() for {k}"
        slim = {k: v for k, v in win.items() if k not in ("grad", "truth_idepth")}
        cmlw.save(os.path.join(OUT, f"{name}_window.cmlw"), slim)
        cmlw.save(os.path.join(OUT, f"{name}_stages.cmlw"), st)
        print(name, "window", os.path.getsize(os.path.join(OUT, f"{name}_window.cmlw")) // 1024, "KB  stages",
              os.path.getsize(os.path.join(OUT, f"{name}_stages.cmlw")) // 1024, "KB")


def maintenance(prior=False):
    """Window-maintenance flow (flag / tryMarginalize / marginalizePointsF / marginalizeFrames) of the reference on a 6-frame window.
    prior=True: disableMarginalization = false, i.e. the marginalisation prior H_M, b_M is live (golden `maintp`: only what differs)."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    win = synth.make_window(160, 120, 6, 60, 4, False, seed=77)
    win["max_frames"] = np.array([5], np.int32)     # the 6th addNewFrame flags one frame by the distance score (BA:649-700)
    win["runs"] = np.array([3], np.int32)           # numGoodResiduals must pass 14 for the first isOOB rule (BA:2532-2536)
    if prior:
        win["disable_marginalization"] = np.array([0], np.int32)
        full = os.path.join(tmp, "maintp.cmlw")
        cmlw.save(full, win)
        run_ref(full, "maintain", os.path.join(tmp, "maintp_out.cmlw"))
        g = cmlw.load(os.path.join(tmp, "maintp_out.cmlw"))
        keep = {k: g[k] for k in ("m2_HM", "m2_bM", "m3_HM", "m3_bM", "m2_pt_marginalized", "m1_pt_outlier", "m3_removed_frames", "m3_frame_in_window", "m4_ok",
                                  "m4_frame_pre_w2c", "m4_frame_affine", "m4_pt_idepth", "m4_pt_alive", "m4_res_point", "m4_res_target")}
        cmlw.save(os.path.join(OUT, "maintp_golden.cmlw"), keep)     # the window is maint_window.cmlw + disable_marginalization = 0
        print("maintp golden", os.path.getsize(os.path.join(OUT, "maintp_golden.cmlw")) // 1024, "KB")
        return
    full = os.path.join(tmp, "maint.cmlw")
    cmlw.save(full, win)
    run_ref(full, "maintain", os.path.join(tmp, "maint_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "maint_out.cmlw"))
    keep = {k: v for k, v in g.items() if k.startswith(("m1_", "m2_", "m3_", "m4_", "runs_ok")) or k in (
        "m0_frame_in_window", "m0_frame_flagged", "m0_frame_num_marginalized", "m0_frame_num_residuals_out", "m0_frame_num_residuals", "m0_frame_keyid",
        "m0_pt_alive", "m0_pt_outlier", "m0_pt_num_good", "m0_pt_idepth_hessian", "m0_pt_last0_state", "m0_pt_last1_state", "m0_res_point", "m0_res_target",
        "m0_res_state", "m0_frame_pre_w2c", "m0_frame_state", "m0_pt_idepth", "m0_frame_energy_th")}
    slim = {k: v for k, v in win.items() if k not in ("grad", "truth_idepth")}
    cmlw.save(os.path.join(OUT, "maint_window.cmlw"), slim)
    cmlw.save(os.path.join(OUT, "maint_golden.cmlw"), keep)
    print("maint window", os.path.getsize(os.path.join(OUT, "maint_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "maint_golden.cmlw")) // 1024, "KB")


def rejection():
    """forceAccept = false on a noisy window: the reference accepts two steps and rejects the rest (BA:843-877)."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    win = synth.make_window(160, 120, 4, 60, 6, True, seed=5, pose_noise=3e-3, idepth_noise=0.05)
    win["force_accept"] = np.array([0], np.int32)
    full = os.path.join(tmp, "reject.cmlw")
    cmlw.save(full, win)
    run_ref(full, "run", os.path.join(tmp, "reject_run.cmlw"))
    g = cmlw.load(os.path.join(tmp, "reject_run.cmlw"))
    assert 0 < int(g["accepted_count"][0]) < 6, "the scenario must contain accepted AND rejected steps"
    keep = {k: g[k] for k in ("accepted_count", "fin_ok", "fin_frame_pre_w2c", "fin_frame_affine", "fin_frame_state", "fin_frame_energy_th", "fin_pt_idepth", "fin_pt_alive",
                              "fin_pt_num_good", "fin_pt_outlier", "fin_alive_res_point", "fin_alive_res_target", "fin_alive_res_state")}
    slim = {k: v for k, v in win.items() if k not in ("grad", "truth_idepth")}
    cmlw.save(os.path.join(OUT, "reject_window.cmlw"), slim)
    cmlw.save(os.path.join(OUT, "reject_golden.cmlw"), keep)
    print("reject window", os.path.getsize(os.path.join(OUT, "reject_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "reject_golden.cmlw")) // 1024, "KB")


def tracker():
    """DSOTracker::makeCoarseDepthL0 + optimize of the reference (SURVEY 8f NEXT #1) on a 3-frame window: frame 1 is the tracking reference, frame 2 the
    frame to track.  Case a: nominal; b: affine brightness start 60 grey levels off (every residual saturated at the coarsest level -> cutoff doubling and the
    one-time level repeat, DSOTracker.cpp:72-76,198-201); c: start pose far off (fewer than 20 terms -> early exit, DSOTracker.cpp:65-69)."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    N, seed = 3, 31
    win = synth.make_window(256, 192, N, 400, 4, True, seed=seed, low_freq=True)
    rng = np.random.default_rng(seed + 1)
    win["track_ref"] = np.array([N - 2], np.int32); win["track_new"] = np.array([N - 1], np.int32)
    cam = win["truth_frame"][N - 1].copy(); cam[9:] += 3e-3 * rng.standard_normal(3)
    win["pt_uncertainty"] = 1.0 / (rng.uniform(50, 5000, win["pt_host"].size) + 0.01)
    win["frame_cam"] = win["truth_frame"].copy(); win["frame_evalpt"] = win["truth_frame"].copy()
    win["pt_idepth"] = win["truth_idepth"] * (1 + 0.01 * rng.standard_normal(win["pt_host"].size))
    far = cam.copy(); far[9:] += np.array([4.0, 0.0, 0.0])
    cases = {"a": (cam, (0.0, 0.0)), "b": (cam, (0.0, 60.0)), "c": (far, (0.0, 0.0))}
    gold = {}
    for name, (c, aff) in cases.items():
        win["track_init_cam"] = c; win["track_new_affine"] = np.array(aff)
        full = os.path.join(tmp, f"track_{name}.cmlw")
        cmlw.save(full, {k: v for k, v in win.items() if k != "grad"})
        run_ref(full, "track", os.path.join(tmp, f"track_{name}_out.cmlw"))
        g = cmlw.load(os.path.join(tmp, f"track_{name}_out.cmlw"))
        if name == "a":
            import tracker_oracle as T
            L = g["trk_K"].shape[0]
            pyr = T.build_pyramid(win["gray"][N - 1], L)
            for l in range(L):       # the reference's pyramid and derivative images must equal the restatement bit for bit; then they are dropped
                assert np.array_equal(pyr[l][1], g[f"trk_grad{l}"]), f"pyramid level {l} differs from the reference"
            for k in ("trk_K", "trk_levels_wh", "trk_pc_n"):
                gold[k] = g[k]
            for l in range(L):
                gold[f"trk_pc{l}"] = g[f"trk_pc{l}"]
        for k in ("trk_cam", "trk_affine", "trk_E", "trk_numTermsInE", "trk_numSaturated", "trk_numRobust", "trk_levelCutoffRepeat", "trk_flow", "trk_relAff", "trk_covariance",
                  "trk_isCorrect", "trk_tooManySaturated"):
            gold[f"{name}_{k}"] = g[k]
        gold[f"{name}_init_cam"] = c; gold[f"{name}_new_affine"] = np.array(aff)
    assert gold["a_trk_isCorrect"][0] == 1 and gold["b_trk_isCorrect"][0] == 1 and gold["c_trk_isCorrect"][0] == 0
    slim = {k: v for k, v in win.items() if k not in ("grad", "truth_idepth", "track_init_cam", "track_new_affine")}
    cmlw.save(os.path.join(OUT, "track_window.cmlw"), slim)
    cmlw.save(os.path.join(OUT, "track_golden.cmlw"), gold)
    print("track window", os.path.getsize(os.path.join(OUT, "track_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "track_golden.cmlw")) // 1024, "KB")


def tracer():
    """DSOTracer (SURVEY 8f NEXT #2) of the reference on a 5-frame window: 4 x 150 immature points created by makeNewTracesFrom, traced into every
    later frame (trace(), all six statuses occur), then optimizeImmaturePoint on the points with a finite depth interval."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    W, H, N, per, seed = 192, 144, 5, 150, 11
    win = synth.make_window(W, H, N, 20, 4, True, seed=seed, low_freq=True)
    rng = np.random.default_rng(seed + 5)
    hosts, xy = [], []
    for h in range(N - 1):
        x = rng.integers(8, W - 8, per).astype(np.float32); y = rng.integers(8, H - 8, per).astype(np.float32)
        hosts.append(np.full(per, h, np.int32)); xy.append(np.stack([x, y], 1))
    win["im_host"] = np.concatenate(hosts); win["im_xy"] = np.concatenate(xy).astype(np.float32)
    win["frame_cam"] = win["truth_frame"].copy()
    full = os.path.join(tmp, "trace.cmlw")
    keep = {k: win[k] for k in ("size", "calib", "frame_cam", "frame_affine", "frame_exposure", "gray", "im_host", "im_xy")}
    cmlw.save(full, keep)
    run_ref(full, "trace", os.path.join(tmp, "trace_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "trace_out.cmlw"))
    seen = np.concatenate([g[f"trc_status_f{f}"] for f in range(1, N)])
    assert set(np.unique(seen[seen >= 0])) == {0, 1, 2, 3, 4, 5}, "the scenario must produce every trace status"
    gold = {k: v for k, v in g.items() if not k.endswith("_seconds")}
    # second run with "Min iDepth H Act" raised so that the Hdd gate (return 0, DSOTracer.cpp:321-324, 341-344) fires on part of the points
    keep2 = dict(keep); keep2["min_idepth_h_act"] = np.array([3.0e4])
    cmlw.save(os.path.join(tmp, "trace_hact.cmlw"), keep2)
    run_ref(os.path.join(tmp, "trace_hact.cmlw"), "trace", os.path.join(tmp, "trace_hact_out.cmlw"))
    g2 = cmlw.load(os.path.join(tmp, "trace_hact_out.cmlw"))
    gold["hact_opt_rc"] = g2["trc_opt_rc"]; gold["hact_opt_idepth"] = g2["trc_opt_idepth"]; gold["hact_min_idepth_h_act"] = np.array([3.0e4])
    assert set(np.unique(np.concatenate([gold["trc_opt_rc"], gold["hact_opt_rc"]]))) >= {-1, 0, 1}, "the scenario must produce every activation outcome"
    print("activation rc", np.unique(gold["trc_opt_rc"], return_counts=True), "with raised gate", np.unique(gold["hact_opt_rc"], return_counts=True))
    cmlw.save(os.path.join(OUT, "trace_window.cmlw"), keep)
    cmlw.save(os.path.join(OUT, "trace_golden.cmlw"), gold)
    print("trace window", os.path.getsize(os.path.join(OUT, "trace_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "trace_golden.cmlw")) // 1024, "KB")


def activate():
    """DSOTracer::activatePoints of the reference (DSOTracer.cpp:62-278) on the trace window (tests/golden/trace_window.cmlw) plus 120 active points
    (distance map) with desiredPointDensity 300: minimum-distance adaptation, distance-map gating, activation, removal."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    base = cmlw.load(os.path.join(OUT, "trace_window.cmlw"))
    W, H = base["size"]; N = base["frame_cam"].shape[0]
    win = synth.make_window(int(W), int(H), N, 20, 4, True, seed=11, low_freq=True, with_gradients=False, with_depth=True)
    assert np.array_equal(win["gray"], base["gray"]), "the trace window must be reproducible"
    rng = np.random.default_rng(2)
    A = 120
    ah = rng.integers(0, N - 1, A).astype(np.int32); ax = rng.integers(8, W - 8, A); ay = rng.integers(8, H - 8, A)
    extra = dict(act_host=ah, act_xy=np.stack([ax, ay], 1).astype(np.float32), act_idepth=1.0 / win["truth_depth"][ah, ay, ax], desired_density=np.array([300], np.int32))
    full = dict(base); full.update(extra)
    cmlw.save(os.path.join(tmp, "activate.cmlw"), full)
    run_ref(os.path.join(tmp, "activate.cmlw"), "trace", os.path.join(tmp, "activate_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "activate_out.cmlw"))
    assert 0 < g["act_mapped"].sum() < g["act_order"].size and 0 < g["act_still_immature"].sum()
    gold = dict(extra)
    gold.update({k: g[k] for k in ("act_order", "act_mapped", "act_still_immature", "act_idepth_out", "act_min_distance", "act_urgent")})
    cmlw.save(os.path.join(OUT, "activate_golden.cmlw"), gold)       # inputs (act_*, desired_density) and outputs; the window is trace_window.cmlw
    print("activate golden", os.path.getsize(os.path.join(OUT, "activate_golden.cmlw")) // 1024, "KB  mapped", int(g["act_mapped"].sum()), "kept", int(g["act_still_immature"].sum()))


def fast():
    """Features::FAST::compute of the reference (SURVEY 8f NEXT #4, first unit of the ORB extractor) on a noisy 200x150 8-bit image with pasted blobs,
    thresholds 20 / 7 (ORB's iniThFAST / minThFAST) and 60."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    W, H = 200, 150
    win = synth.make_window(W, H, 2, 10, 1, False, seed=9, with_gradients=False)
    rng = np.random.default_rng(1)
    img = np.clip(win["gray"][0] + rng.normal(0, 6, (H, W)), 0, 255).astype(np.uint8)
    for _ in range(40):
        x, y = rng.integers(5, W - 10), rng.integers(5, H - 10)
        img[y:y + rng.integers(2, 6), x:x + rng.integers(2, 6)] = rng.integers(0, 256)
    w = dict(gray_u8=img, thresholds=np.array([20, 7, 60], np.int32))
    cmlw.save(os.path.join(tmp, "fast.cmlw"), w)
    run_ref(os.path.join(tmp, "fast.cmlw"), "fast", os.path.join(tmp, "fast_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "fast_out.cmlw"))
    assert all(g[f"fast_xy{k}"].shape[0] > 10 for k in range(3))
    gold = dict(w); gold.update({k: v for k, v in g.items() if k != "fast_seconds"})
    cmlw.save(os.path.join(OUT, "fast_golden.cmlw"), gold)
    print("fast golden", os.path.getsize(os.path.join(OUT, "fast_golden.cmlw")) // 1024, "KB", [int(g[f"fast_xy{k}"].shape[0]) for k in range(3)])


def prepare():
    """CaptureImageGenerator::generate of the reference (SURVEY 8f NEXT #3) with LUT, inverse vignette and a radtan pre-undistorter: 200x150 -> 160x120."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    w = synth.prepare_scenario(200, 150, 160, 120)
    cmlw.save(os.path.join(tmp, "prepare.cmlw"), w)
    run_ref(os.path.join(tmp, "prepare.cmlw"), "prepare", os.path.join(tmp, "prepare_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "prepare_out.cmlw"))
    assert 0.05 < np.isnan(g["prep_map"][..., 0]).mean() < 0.5, "the map must contain pixels outside the sensor image"
    gold = {k: v for k, v in g.items() if k != "prep_seconds"}
    cmlw.save(os.path.join(OUT, "prepare_window.cmlw"), w)
    cmlw.save(os.path.join(OUT, "prepare_golden.cmlw"), gold)
    print("prepare window", os.path.getsize(os.path.join(OUT, "prepare_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "prepare_golden.cmlw")) // 1024, "KB")


def select():
    """PixelSelector::compute of the reference (SURVEY 8f NEXT #4, PixelSelector part) on one 256x192 frame, five successive densities on the same
    selector (the potential carries over; densities 150 and 2500 trigger the re-sampling recursion in both directions, 600 the random sub-sampling)."""
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    W, H = 256, 192
    win = synth.make_window(W, H, 2, 10, 1, False, seed=9, low_freq=True)
    w = dict(size=np.array([W, H], np.int32), calib=win["calib"], gray=win["gray"][0], densities=np.array([600.0, 600.0, 150.0, 2500.0, 2500.0]))
    cmlw.save(os.path.join(tmp, "select.cmlw"), w)
    run_ref(os.path.join(tmp, "select.cmlw"), "select", os.path.join(tmp, "select_out.cmlw"))
    g = cmlw.load(os.path.join(tmp, "select_out.cmlw"))
    pots = [int(g[f"sel_pot_after{d}"][0]) for d in range(5)]
    assert len(set(pots)) >= 3, "the potential must move in both directions"
    gold = {k: v for k, v in g.items() if k != "sel_seconds"}
    cmlw.save(os.path.join(OUT, "select_window.cmlw"), w)
    cmlw.save(os.path.join(OUT, "select_golden.cmlw"), gold)
    print("select window", os.path.getsize(os.path.join(OUT, "select_window.cmlw")) // 1024, "KB  golden", os.path.getsize(os.path.join(OUT, "select_golden.cmlw")) // 1024, "KB", pots)


def full_size_summaries():
    """Compact summaries of the reference's run() at BASELINE.json's full sizes (c1..c4): final poses, affine, summed residual energy, accepted-step count,
    every 16th inverse depth and a digest of the surviving (point, target) residual set.  The windows are regenerated from libcml_b200.synth
    (seeded), so only the summary travels; tests/test_gpu_parity.py::test_against_reference_summary_full_size reads it on the GPU box.
        python oracle/make_golden.py summaries"""
    import hashlib
    tmp = "/tmp/cmlba_golden"
    os.makedirs(tmp, exist_ok=True)
    out = {}
    for cfg in ("c1", "c2", "c3", "c4"):
        win = synth.make_config(cfg, with_gradients=False)
        wp = os.path.join(tmp, f"{cfg}.cmlw"); op = os.path.join(tmp, f"{cfg}_run.cmlw")
        cmlw.save(wp, {k: v for k, v in win.items() if k != "grad"})
        run_ref(wp, "run", op)
        g = cmlw.load(op)
        key = np.sort(g["fin_alive_res_point"].astype(np.int64) * 64 + g["fin_alive_res_target"].astype(np.int64))
        out[f"{cfg}_ok"] = g["fin_ok"]
        out[f"{cfg}_w2c"] = g["fin_frame_pre_w2c"]
        out[f"{cfg}_affine"] = g["fin_frame_affine"]
        out[f"{cfg}_energy"] = np.array([g["fin_alive_res_energy"].sum()])
        out[f"{cfg}_accepted"] = g["accepted_count"]
        out[f"{cfg}_idepth16"] = g["fin_pt_idepth"][::16].copy()
        out[f"{cfg}_alive16"] = g["fin_pt_alive"][::16].astype(np.uint8)
        out[f"{cfg}_n_alive_res"] = np.array([key.size], np.int64)
        out[f"{cfg}_n_alive_pts"] = np.array([int(g["fin_pt_alive"].sum())], np.int64)
        out[f"{cfg}_res_digest"] = np.frombuffer(hashlib.sha1(key.tobytes()).digest(), dtype=np.uint8).copy()
        out[f"{cfg}_res_per_target"] = np.bincount(g["fin_alive_res_target"].astype(np.int64), minlength=win["frame_cam"].shape[0]).astype(np.int64)
        print(cfg, "accepted", int(g["accepted_count"][0]), "alive residuals", key.size, "energy", out[f"{cfg}_energy"])
    cmlw.save(os.path.join(OUT, "fullsize_summary.cmlw"), out)
    print("fullsize_summary.cmlw", os.path.getsize(os.path.join(OUT, "fullsize_summary.cmlw")) // 1024, "KB")


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    if len(sys.argv) > 1 and sys.argv[1] == "summaries":
        full_size_summaries(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "select":
        select(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fast":
        fast(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "activate":
        activate(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prepare":
        prepare(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tracer":
        tracer(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tracker":
        tracker(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "maintenance":
        maintenance(); maintenance(prior=True); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "rejection":
        rejection(); sys.exit(0)
    main()
