"""Measures the immature-point tracer (SURVEY.md 8f NEXT #2) on a synthetic 640x480 window: device time of one traceNewCoarse pass and of a
batched optimizeImmaturePoint, end-to-end call times.  Run as `python bench.py --component tracer` it also receives bench.py's cpu_baseline
callback and times the unmodified reference's trace() / optimizeImmaturePoint loops on the host CPU for the same inputs.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import DSOTracer, synth  # noqa: E402


def main(argv=None, reference=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--frames", type=int, default=7); ap.add_argument("--points", type=int, default=2000); ap.add_argument("--repeats", type=int, default=20)
    a = ap.parse_args(argv)
    W, H, N, per = a.width, a.height, a.frames, a.points
    win = synth.make_window(W, H, N, 20, 4, True, seed=11, low_freq=True)
    rng = np.random.default_rng(16)
    cams = win["truth_frame"]
    exps = [(win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1]) for i in range(N)]
    xy = {h: np.stack([rng.integers(8, W - 8, per), rng.integers(8, H - 8, per)], 1).astype(np.float32) for h in range(N - 1)}
    trc = DSOTracer(W, H, win["calib"])
    trace_ms, trace_e2e, traced = [], [], []
    for f in range(N):
        trc.addFrame(f, win["gray"][f], cams[f], exps[f])
        if f > 0:
            t0 = time.perf_counter()
            hist = trc.traceNewCoarse(f)
            trace_e2e.append((time.perf_counter() - t0) * 1e3); trace_ms.append(trc.last_gpu_ms); traced.append(int(hist.sum()))
        if f < N - 1:
            trc.makeNewTracesFrom(f, xy[f])
    pts = trc.getPoints()
    cand = np.nonzero(np.isfinite(pts["idepth_max"]) & (pts["host_frame_slot"] >= 0))[0]
    t0 = time.perf_counter()
    res = trc.optimizeImmaturePoint(cand)
    act_e2e = (time.perf_counter() - t0) * 1e3; act_ms = trc.last_gpu_ms
    out = {"workload": f"{W}x{H}, {N} frames, {per} immature points per keyframe", "points": int(pts.size), "traces": int(sum(traced)),
           "trace_pass_device_ms": [round(float(v), 4) for v in trace_ms], "trace_pass_e2e_ms": [round(float(v), 4) for v in trace_e2e],
           "traces_per_s_device": round(sum(traced) / (sum(trace_ms) * 1e-3)), "status_last_pass": np.bincount(pts["status"], minlength=6).tolist(),
           "activation_points": int(cand.size), "activation_device_ms": round(float(act_ms), 4), "activation_e2e_ms": round(act_e2e, 4),
           "activation_rc": {int(k): int(v) for k, v in zip(*np.unique(res["rc"], return_counts=True))}}
    g = None
    if reference:
        keep = {k: win[k] for k in ("size", "calib", "frame_affine", "frame_exposure", "gray")}
        keep["frame_cam"] = cams
        keep["im_host"] = np.concatenate([np.full(per, h, np.int32) for h in range(N - 1)]); keep["im_xy"] = np.concatenate([xy[h] for h in range(N - 1)])
        g = reference("trace", keep, 1)
    if g is not None:
        out["reference_cpu_trace_ms_total"] = round(float(g["trc_trace_seconds"][0]) * 1e3, 3)
        out["reference_cpu_activation_ms"] = round(float(g["trc_opt_seconds"][0]) * 1e3, 3)
        st = g[f"trc_status_f{N - 1}"]
        out["status_mismatches_vs_reference"] = int((st != pts["status"]).sum())
        todo = g["trc_opt_rc"] != -2
        out["activation_rc_mismatches_vs_reference"] = int((g["trc_opt_rc"][cand] != res["rc"]).sum()) if todo.sum() == cand.size else "candidate sets differ"
        out["speedup_trace_e2e_vs_reference_cpu"] = round(out["reference_cpu_trace_ms_total"] / sum(trace_e2e), 1)
        out["speedup_activation_e2e_vs_reference_cpu"] = round(out["reference_cpu_activation_ms"] / act_e2e, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
