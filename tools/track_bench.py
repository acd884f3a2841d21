"""Measures the coarse tracker (SURVEY.md 8f NEXT #1) on a synthetic 640x480 window: device time of one optimize launch, end-to-end time of
cmltrk_track (host gray image in, Residual out), K-candidate batches.  Run as `python bench.py --component tracker` it also receives bench.py's
cpu_baseline callback and times the unmodified reference's DSOTracker::optimize on the host CPU for the same inputs.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import DSOTracker, synth  # noqa: E402


def scenario(W, H, N, ppk, seed):
    win = synth.make_window(W, H, N, ppk, 4, True, seed=seed, low_freq=True)
    rng = np.random.default_rng(seed + 1)
    win["track_ref"] = np.array([N - 2], np.int32); win["track_new"] = np.array([N - 1], np.int32)
    cam = win["truth_frame"][N - 1].copy(); cam[9:] += 3e-3 * rng.standard_normal(3)
    win["track_init_cam"] = cam
    win["pt_uncertainty"] = 1.0 / (rng.uniform(50, 5000, win["pt_host"].size) + 0.01)
    win["track_new_affine"] = np.array([0.0, 0.0])
    win["frame_cam"] = win["truth_frame"].copy(); win["frame_evalpt"] = win["truth_frame"].copy()
    win["pt_idepth"] = win["truth_idepth"] * (1 + 0.01 * rng.standard_normal(win["pt_host"].size))
    return win


def main(argv=None, reference=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--frames", type=int, default=7); ap.add_argument("--points", type=int, default=2000)
    ap.add_argument("--repeats", type=int, default=50); ap.add_argument("--cluster", type=int, default=16)
    ap.add_argument("--candidates", type=int, default=8); ap.add_argument("--threads", type=int, default=384)
    a = ap.parse_args(argv)
    win = scenario(a.width, a.height, a.frames, a.points, 31)
    N = a.frames; ref, new = N - 2, N - 1
    trk = DSOTracker(a.width, a.height, win["calib"], cluster_ctas=a.cluster, cta_threads=a.threads)
    keep = win["pt_host"] != new
    ref_exp = (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1])
    t0 = time.perf_counter()
    for _ in range(5):
        trk.makeCoarseDepthL0(win["gray"][ref], win["frame_cam"][ref], ref_exp, win["frame_cam"], win["pt_host"][keep], win["pt_xy"][keep], win["pt_idepth"][keep],
                              win["pt_uncertainty"][keep])
    coarse_ms = (time.perf_counter() - t0) / 5 * 1e3
    gray = win["gray"][new]; tau = win["frame_exposure"][new]
    for _ in range(5):
        r = trk.optimize(win["track_init_cam"], win["track_new_affine"], gray=gray, exposure_time=tau)
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        r = trk.optimize(win["track_init_cam"], win["track_new_affine"], gray=gray, exposure_time=tau)
    e2e_ms = (time.perf_counter() - t0) / a.repeats * 1e3
    stage = trk.frameBuffer()
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        stage[...] = gray                    # stands for the producer writing the image (not part of the tracker)
    fill_ms = (time.perf_counter() - t0) / a.repeats * 1e3
    for _ in range(3):
        trk.optimize(win["track_init_cam"], win["track_new_affine"], gray=stage, exposure_time=tau)
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        rz = trk.optimize(win["track_init_cam"], win["track_new_affine"], gray=stage, exposure_time=tau)
    e2e_pinned_ms = (time.perf_counter() - t0) / a.repeats * 1e3
    assert np.array_equal(rz.camera, r.camera)
    dev_ms = trk.benchOptimize(a.repeats)
    cyc = trk.read("cycles", np.int64)
    K = a.candidates
    cams = np.tile(win["track_init_cam"], (K, 1)); cams[:, 9:] += 1e-3 * np.random.default_rng(3).standard_normal((K, 3))
    trk.optimize(cams, np.zeros((K, 2)))
    devK_ms = trk.benchOptimize(a.repeats)
    out = {"workload": f"{a.width}x{a.height}, {N - 1} keyframes x {a.points} points, 5 levels", "pc_n": trk.read("pc_n", np.int32).tolist(),
           "iterations": int(r.iterations), "is_correct": bool(r.isCorrect), "cam_err_vs_truth": float(np.abs(r.camera - win["truth_frame"][new]).max()),
           "cluster_ctas": a.cluster, "cta_threads": a.threads, "optimize_device_ms": round(dev_ms, 4), "us_per_gn_step": round(dev_ms * 1e3 / max(r.iterations + 5, 1), 3),
           "evals": int(cyc[0]), "kcycles_advance_eval_reduce": [round(float(c) / 1e3, 1) for c in cyc[1:]], "track_e2e_ms": round(e2e_ms, 4), "track_e2e_from_frame_buffer_ms": round(e2e_pinned_ms, 4), "optimize_gpu_ms_in_call": round(float(rz.gpu_ms), 4), "h2d_bytes_per_track": int(gray.nbytes), "make_coarse_depth_e2e_ms": round(coarse_ms, 3),
           f"optimize_{K}_candidates_device_ms": round(devK_ms, 4)}
    g = reference("track", {k: v for k, v in win.items() if k != "grad"}, 5) if reference else None
    if g is not None:
        out["reference_cpu_optimize_ms"] = round(float(g["trk_seconds"][0]) * 1e3, 4)
        out["cam_diff_vs_reference"] = float(np.abs(r.camera - g["trk_cam"]).max())
        out["speedup_e2e_vs_reference_cpu"] = round(out["reference_cpu_optimize_ms"] / e2e_ms, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
