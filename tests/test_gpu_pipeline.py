"""The device-resident hand-over between the components: a frame prepared once by cmlimg is consumed by the tracker (texels sampled in place),
the tracer and the bundle adjustment (device-to-device copies) without going back to the host.  Every consumer must produce exactly the bits
it produces from the host image."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import cmlw  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_tracker_from_device_levels_is_bit_identical():
    from libcml_b200 import CaptureImageGenerator, DSOTracker
    win = cmlw.load(os.path.join(GOLDEN, "track_window.cmlw")); g = cmlw.load(os.path.join(GOLDEN, "track_golden.cmlw"))
    H, W = win["gray"].shape[1:]
    ref, new = int(win["track_ref"][0]), int(win["track_new"][0])
    keep = win["pt_host"] != new
    ref_exp = (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1])
    args = (win["frame_cam"][ref], ref_exp, win["frame_cam"], win["pt_host"][keep], win["pt_xy"][keep], win["pt_idepth"][keep], win["pt_uncertainty"][keep])
    host = DSOTracker(W, H, win["calib"])
    host.makeCoarseDepthL0(win["gray"][ref], *args)
    a = host.optimize(g["a_init_cam"], g["a_new_affine"], gray=win["gray"][new], exposure_time=win["frame_exposure"][new])
    gen_ref, gen_new = CaptureImageGenerator(W, H), CaptureImageGenerator(W, H)
    dev = DSOTracker(W, H, win["calib"])
    dev.makeCoarseDepthL0Device(gen_ref.generate(win["gray"][ref]), *args)
    dev.setFrameDevice(gen_new.generate(win["gray"][new]), win["frame_exposure"][new])
    b = dev.optimize(g["a_init_cam"], g["a_new_affine"])
    assert np.array_equal(a.camera, b.camera) and np.array_equal(a.exposure, b.exposure) and np.array_equal(a.E, b.E) and a.iterations == b.iterations
    assert b.kernel_launches == 1                          # the optimisation alone: no pyramid kernels on this path
    for l in range(5):
        assert np.array_equal(host.read(f"pc{l}", np.float32), dev.read(f"pc{l}", np.float32))


def test_tracer_from_device_frame_is_bit_identical():
    from libcml_b200 import CaptureImageGenerator, DSOTracer
    win = cmlw.load(os.path.join(GOLDEN, "trace_window.cmlw"))
    H, W = win["gray"].shape[1:]
    N = 3
    ex = [(win["frame_exposure"][i], win["frame_affine"][i, 0], win["frame_affine"][i, 1]) for i in range(N)]
    gen = CaptureImageGenerator(W, H)

    def flow(device_frames):
        trc = DSOTracer(W, H, win["calib"])
        for f in range(N):
            if device_frames:
                trc.addFrameDevice(f, gen.generate(win["gray"][f]), win["frame_cam"][f], ex[f])
            else:
                trc.addFrame(f, win["gray"][f], win["frame_cam"][f], ex[f])
            if f > 0:
                trc.traceNewCoarse(f)
            sel = np.nonzero(win["im_host"] == f)[0]
            trc.makeNewTracesFrom(f, win["im_xy"][sel])
        pts = trc.getPoints()
        cand = np.nonzero(np.isfinite(pts["idepth_max"]))[0]
        return pts, trc.optimizeImmaturePoint(cand)
    p0, r0 = flow(False)
    p1, r1 = flow(True)
    assert p0.tobytes() == p1.tobytes() and r0.tobytes() == r1.tobytes()


def test_bundle_adjustment_from_device_texels_is_bit_identical():
    from libcml_b200 import CaptureImageGenerator, DSOBundleAdjustment, synth
    win = synth.make_config("tiny")
    H, W = win["gray"].shape[1:]
    N = win["frame_evalpt"].shape[0]
    P = win["pt_host"].size
    gen = CaptureImageGenerator(W, H)

    def run(device_frames):
        ba = DSOBundleAdjustment(device=0, iterations=int(win["iterations"][0]))
        ba.setCalibration(*[float(v) for v in win["calib"]], W, H)
        for f in range(N):
            a = (f, win["frame_evalpt"][f], win["frame_affine"][f, 0], win["frame_affine"][f, 1], win["frame_exposure"][f])
            if device_frames:
                ba.addNewFrameDevice(*a, gen.generate(win["gray"][f]).devicePtr("texel0"), bool(win["frame_init"][f]) if "frame_init" in win else False)
            else:
                ba.addNewFrameGray(*a, win["gray"][f], bool(win["frame_init"][f]) if "frame_init" in win else False)
        ba.addPoints(np.arange(P), win["pt_host"], win["pt_xy"], win["pt_idepth"])
        assert ba.run(win["frame_cam"], iterations=int(win["iterations"][0]))
        return ba.getFrames(), ba.getPoints()
    f0, p0 = run(False)
    f1, p1 = run(True)
    assert np.array_equal(f0["world_to_cam"], f1["world_to_cam"]) and np.array_equal(f0["affine"], f1["affine"])
    assert np.array_equal(p0["idepth"], p1["idepth"]) and np.array_equal(p0["id"], p1["id"])


def test_direct_pipeline_end_to_end_on_synthetic_truth():
    """prepare -> select -> trace -> activate -> bundle adjustment -> track, all on the device, against the analytic truth of the synthetic scene
    (tools/pipeline_demo.py): every stage must land where the geometry says."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pipeline_demo
    r = pipeline_demo.run()
    assert min(r["selected_per_keyframe"]) > 100 and r["activated"] > 0.5 * r["traced_points"] and r["activated"] + r["removed"] <= r["traced_points"]
    assert r["activation_idepth_median_rel_err"] < 5e-3 and r["activation_idepth_p90_rel_err"] < 2e-2          # tracer + activation recover the plane's depth
    assert r["ba_ok"] and r["ba_energy_last"] < 0.1 * r["ba_energy_first"]
    assert r["ba_reproj_px_after"] < 0.1 and r["ba_reproj_px_after"] < 0.25 * r["ba_reproj_px_before"]         # BA: 0.45 px -> 0.04 px
    assert r["track_ok"] and r["track_reproj_px_after"] < 0.15 and r["track_reproj_px_after"] < 0.25 * r["track_reproj_px_before"]


def test_degenerate_inputs_fail_cleanly():
    """Empty point sets and texture-less images: every component must return the reference's 'nothing to do' outcome, never hang or crash."""
    from libcml_b200 import CaptureImageGenerator, DSOTracer, DSOTracker, PixelSelector
    W, H = 128, 96
    K = (100.0, 100.0, 63.5, 47.5)
    flat = np.full((H, W), 100.0, np.float32)
    cam = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    cap = CaptureImageGenerator(W, H).generate(flat)
    # selector: no gradient anywhere -> no corners, the potential walks down to 1 and stops
    sel = PixelSelector(W, H)
    xy, ty = sel.compute(cap, 500.0)
    assert xy.shape == (0, 2) and ty.size == 0 and sel.currentPotential >= 1
    # tracker: no points -> fewer than 20 terms at the coarsest level -> not correct, camera untouched (DSOTracker.cpp:65-69)
    trk = DSOTracker(W, H, K)
    trk.makeCoarseDepthL0(flat, cam, (1.0, 0.0, 0.0), cam[None], np.zeros(0, np.int32), np.zeros((0, 2), np.float32), np.zeros(0), np.zeros(0))
    r = trk.optimize(cam, (0.0, 0.0), gray=flat)
    assert not r.isCorrect and np.array_equal(r.camera, cam) and r.iterations == 0
    # tracer: tracing / activating nothing
    trc = DSOTracer(W, H, K)
    trc.addFrame(0, flat, cam, (1.0, 0.0, 0.0)); trc.addFrame(1, flat, cam, (1.0, 0.0, 0.0))
    assert trc.traceNewCoarse(1).sum() == 0 and trc.optimizeImmaturePoint([]).size == 0
    ids = trc.makeNewTracesFrom(0, [[40.0, 40.0], [60.0, 50.0]])
    hist = trc.traceNewCoarse(1)                              # identical pose: zero baseline, the search direction is undefined -> OOB like the reference
    assert hist.sum() == 2
    assert set(trc.getPoints()["status"][ids]) <= {1, 2, 3, 4}


def test_stream_cycle_small():
    """The chained per-frame cycle of BASELINE.json configs[4] (tools/stream_bench.py = `bench.py --workload c5`) on a small sequence: every frame is
    prepared, tracked, traced, selected, activated and bundle-adjusted on the device; the sliding window stays at its size, the tracker stays on
    the truth and the BA keeps running."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stream_bench
    out = stream_bench.main(["--width", "320", "--height", "240", "--frames", "13", "--window", "6", "--density", "400"])
    assert out["steps"] == 5 and out["value"] > 0
    assert all(f["window"] == 6 and f["ba_iterations"] >= 1 and f["ba_residuals"] > 1000 and f["activated"] > 50 for f in out["frames"])
    assert out["track_translation_error_rel_to_motion"]["max"] < 0.05
