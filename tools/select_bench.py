"""Measures the pixel selector (SURVEY.md 8f NEXT #4, PixelSelector part): one compute() per keyframe at 640x480 (density 600 -> ~2000 like the
reference's settings): call time, stream time.  Run as `python bench.py --component select` it also receives bench.py's cpu_baseline callback: the
unmodified reference's PixelSelector::compute on the host CPU plus an exact comparison of the selected corners.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import CaptureImageGenerator, PixelSelector, synth  # noqa: E402


def main(argv=None, reference=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=480); ap.add_argument("--repeats", type=int, default=20)
    a = ap.parse_args(argv)
    W, H = a.width, a.height
    win = synth.make_window(W, H, 2, 10, 1, False, seed=9, low_freq=True)
    gray = win["gray"][0]
    dens = [600.0, 600.0, 2000.0, 2000.0, 2000.0]
    cap = CaptureImageGenerator(W, H).generate(gray)
    sel = PixelSelector(W, H)
    out = {"workload": f"{W}x{H}, densities {dens}"}
    res = []
    for d in dens:
        t0 = time.perf_counter()
        xy, ty = sel.compute(cap, d)
        res.append((xy, ty, (time.perf_counter() - t0) * 1e3, sel.last_gpu_ms, sel.currentPotential))
    out["selected"] = [int(r[0].shape[0]) for r in res]; out["potential_after"] = [r[4] for r in res]
    out["compute_e2e_ms"] = [round(r[2], 3) for r in res]; out["compute_stream_ms"] = [round(float(r[3]), 3) for r in res]
    t0 = time.perf_counter()
    for _ in range(a.repeats):
        sel.compute(cap, 2000.0)
    out["steady_compute_e2e_ms"] = round((time.perf_counter() - t0) / a.repeats * 1e3, 3)
    g = reference("select", dict(size=np.array([W, H], np.int32), calib=win["calib"], gray=gray, densities=np.array(dens)), 1) if reference else None
    if g is not None:
        out["reference_cpu_compute_ms_best"] = round(float(g["sel_seconds"][0]) * 1e3, 3)
        out["identical_to_reference"] = [bool(np.array_equal(res[d][0], g[f"sel_xy{d}"]) and np.array_equal(res[d][1], g[f"sel_type{d}"])) for d in range(len(dens))]
        out["speedup_e2e_vs_reference_cpu"] = round(out["reference_cpu_compute_ms_best"] / out["steady_compute_e2e_ms"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
