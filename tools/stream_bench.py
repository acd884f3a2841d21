"""BASELINE.json configs[4] (C5): the direct-odometry cycle of Hybrid::directMap (slam/modslam/direct/Mapping.cpp:47-134) as a STREAM, every
stage on the device, chained through device-resident frames:
    prepare (cmlimg) -> track against the newest keyframe (cmltrk) -> trace the immature points into the new frame (cmltrc) -> select pixels
    (cmlsel) -> new immature points -> activate -> photometric BA over the sliding window of N keyframes with window maintenance (cmlba)
    -> depth map of the new reference keyframe for the tracker.
Every frame of the synthetic sequence becomes a keyframe (the heaviest per-frame path; the reference inserts keyframes less often).  A frame crosses
PCIe once (the raw gray image).  Prints one JSON line: frames/s over the timed frames, per-stage milliseconds, sanity numbers against the truth
(tracked pose error, BA reprojection).  With --gpus N (torchrun) every rank runs its own replica of the stream (replicated streams, no
collective: frames/s adds up).  The reference CPU time per frame is assembled by bench.py from the reference's own stages (see `reference`).
    python bench.py --workload c5 [--width 1920 --height 1080 --frames 12 --window 8 --density 2000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import CaptureImageGenerator, DSOBundleAdjustment, DSOTracer, DSOTracker, PixelSelector, synth  # noqa: E402


def main(argv=None, reference=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--window", type=int, default=8, help="keyframes in the sliding window")
    ap.add_argument("--frames", type=int, default=16, help="frames of the sequence (the first `window` + 2 fill the window and size the buffers, untimed)")
    ap.add_argument("--density", type=int, default=2000, help="desired points per keyframe")
    ap.add_argument("--gpus", type=int, default=1); ap.add_argument("--steps", type=int, default=0); ap.add_argument("--warmup", type=int, default=0)
    a, _ = ap.parse_known_args(argv)
    W, H, NW, NF, density = a.width, a.height, a.window, a.frames, a.density
    trk_ready[0] = False
    if NF > 16:
        NF = 16
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    win = synth.make_window(W, H, NF, 10, 6, False, seed=5 + rank, low_freq=True, with_gradients=False, with_depth=True)
    truth = win["truth_frame"]; K = win["calib"]
    gray = torch.from_numpy(np.ascontiguousarray(win["gray"], dtype=np.float32)).pin_memory().numpy()
    ex = (1.0, 0.0, 0.0)
    rng = np.random.default_rng(7 + rank)
    G = NW + 3
    gens = [CaptureImageGenerator(W, H, device=local) for _ in range(G)]          # a ring of generators keeps the levels of every live frame on the device
    sel = PixelSelector(W, H, device=local); trc = DSOTracer(W, H, K, device=local); trk = DSOTracker(W, H, K, device=local)
    ba = DSOBundleAdjustment(device=local, iterations=6, max_frames=NW)
    ba.setCalibration(*[float(v) for v in K], W, H)
    caps, pose = {}, {}
    imm_ids, imm_xy, imm_ty = {}, {}, {}                 # immature points per host frame
    all_host = np.zeros(NF * 4 * density + 16, np.int64); all_xy = np.zeros((NF * 4 * density + 16, 2), np.float32)      # BA point id -> host frame id, pixel
    next_pid = [0]
    T, per_frame = {}, []

    def lap(name, t0):
        dt = (time.perf_counter() - t0) * 1e3
        T.setdefault(name, []).append(dt)
        return time.perf_counter()

    track_err, reproj = [], []
    for f in range(NF):
        timed = f >= NW + 2            # the window is full from frame NW on; its first two full cycles still grow buffers (N = NW + 1 inside run())
        torch.cuda.synchronize()
        t_frame = time.perf_counter(); t0 = t_frame
        cap = gens[f % G].generate(gray[f]); caps[f] = cap
        t0 = lap("prepare", t0) if timed else time.perf_counter()
        # ---- coarse tracking against the newest keyframe (the first frames of the sequence start from the truth)
        if f >= 2 and trk_ready[0]:
            guess = pose[f - 1].copy()                                                   # motion model: the last pose
            trk.setFrameDevice(cap, 1.0); r = trk.optimize(guess, (0.0, 0.0))
            cam = r.camera if r.isCorrect else truth[f]
            track_err.append(float(np.linalg.norm(cam[9:] - truth[f][9:]) / max(np.linalg.norm(truth[f][9:] - truth[f - 1][9:]), 1e-12)))
        else:
            cam = truth[f].copy()
        pose[f] = cam
        t0 = lap("track", t0) if timed else time.perf_counter()
        # ---- immature points: trace the existing ones into this frame, then seed new ones from the selector
        trc.addFrameDevice(f, cap, cam, ex)
        if f > 0:
            trc.traceNewCoarse(f)
        t0 = lap("trace", t0) if timed else time.perf_counter()
        xy, ty = sel.compute(cap, density)
        t0 = lap("select", t0) if timed else time.perf_counter()
        ids = trc.makeNewTracesFrom(f, xy)
        imm_ids[f], imm_xy[f], imm_ty[f] = ids, xy, ty
        t0 = lap("new_traces", t0) if timed else time.perf_counter()
        # ---- activation of traced immature points hosted in the older window frames
        hosts = [h for h in imm_ids if h != f and imm_ids[h].size]
        act_n = 0
        if hosts:
            cand = np.concatenate([imm_ids[h] for h in hosts]); cxy = np.concatenate([imm_xy[h] for h in hosts]); cty = np.concatenate([imm_ty[h] for h in hosts])
            chost = np.concatenate([np.full(imm_ids[h].size, h) for h in hosts])
            act_ids, act, rem_ids, st = trc.activatePoints(f, np.zeros((0, 2)), cand, desiredPointDensity=density * (NW - 1), types=cty)
            where = {int(i): k for k, i in enumerate(cand)}
            k_act = np.array([where[int(i)] for i in act_ids], dtype=np.int64)
            gone = set(int(i) for i in act_ids) | set(int(i) for i in rem_ids)
            for h in hosts:
                keep = np.array([int(i) not in gone for i in imm_ids[h]], dtype=bool)
                imm_ids[h], imm_xy[h], imm_ty[h] = imm_ids[h][keep], imm_xy[h][keep], imm_ty[h][keep]
            act_n = int(k_act.size)
        t0 = lap("activate", t0) if timed else time.perf_counter()
        # ---- photometric BA over the window + maintenance
        fr_ids = ba.getFrames()["id"] if f > 0 else np.zeros(0, np.int64)
        if fr_ids.size:
            ba.flagFramesForMarginalization(np.stack([pose[int(i)] for i in fr_ids]))
        ba.addNewFrameDevice(f, cam, 0.0, 0.0, 1.0, cap.devicePtr("texel0"), f == 0)
        if act_n:
            new_ids = np.arange(next_pid[0], next_pid[0] + act_n); next_pid[0] += act_n
            ba.addPoints(new_ids, chost[k_act], cxy[k_act], act["idepth"].astype(np.float64))
            all_host[new_ids] = chost[k_act]; all_xy[new_ids] = cxy[k_act]
        t0 = lap("ba.add", t0) if timed else time.perf_counter()
        ran = False
        if f >= 1 and ba.numPoints() > 0:
            fr_ids = ba.getFrames()["id"]
            if timed and os.environ.get("STREAM_HOST_TIMING"):
                ba.read("host_timing_reset", np.uint8)
            ran = ba.run(np.stack([pose[int(i)] for i in fr_ids]), iterations=6)
            t0 = lap("ba.run", t0) if timed else time.perf_counter()
            frs = ba.getFrames(); bp = ba.getPoints()
            for i, c in zip(frs["id"], frs["world_to_cam"]):
                pose[int(i)] = c
            t0 = lap("ba.results", t0) if timed else time.perf_counter()
            ba.tryMarginalize()
            ba.marginalizePointsF()
            for gone_f in ba.marginalizeFrames():
                trc.removeFrame(int(gone_f)); imm_ids.pop(int(gone_f), None); imm_xy.pop(int(gone_f), None); imm_ty.pop(int(gone_f), None)
            t0 = lap("ba.maintenance", t0) if timed else time.perf_counter()
        # ---- depth map of the new reference keyframe for the tracker
        if ran:
            frs = ba.getFrames(); bp = ba.getPoints()
            slot_of = np.full(NF + 1, -1, np.int32); slot_of[frs["id"]] = np.arange(frs["id"].size)
            hs_all = slot_of[all_host[bp["id"]]]
            keep = hs_all >= 0
            if keep.sum() > 50 and slot_of[f] >= 0:
                slot = {f: int(slot_of[f])}
                hs = hs_all[keep].astype(np.int32)
                pxy = all_xy[bp["id"][keep]]
                trk.makeCoarseDepthL0Device(cap, frs["world_to_cam"][slot[f]], ex, frs["world_to_cam"], hs, pxy, bp["idepth"][keep], np.full(int(keep.sum()), 1e-3))
                trk_ready[0] = True
        t0 = lap("tracker_depth", t0) if timed else time.perf_counter()
        torch.cuda.synchronize()
        if timed:
            per_frame.append((time.perf_counter() - t_frame) * 1e3)
            reproj.append({"frame": f, "window": int(ba.getFrames()["id"].size), "ba_points": int(ba.numPoints()), "activated": act_n,
                           "ba_residuals": int(ba.last_result.num_residuals) if ran else 0, "ba_iterations": int(ba.last_result.iterations_done) if ran else 0})
    ms = float(np.mean(per_frame))
    fps = 1e3 / ms
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.cpu()[0]); fps = world * 1e3 / ms
    out = {"metric": "frames/s of the direct-odometry cycle (every frame a keyframe) at %dx%d, %d-keyframe window" % (W, H, NW), "value": fps, "unit": "frames/s",
           "n_gpus": world, "steps": len(per_frame), "warmup": NW + 2, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (projection f64)", "data": "synthetic",
           "config": {"workload": "c5: %dx%d stream, sliding window of %d keyframes, %d points per keyframe desired" % (W, H, NW, density),
                      "parallelism": "replicated streams x%d (no collective)" % world},
           "stage_ms": {k: round(float(np.mean(v)), 3) for k, v in T.items()},
           "h2d_bytes_per_step": int(gray[0].nbytes), "frames": reproj,
           "track_translation_error_rel_to_motion": {"median": float(np.median(track_err)) if track_err else None, "max": float(np.max(track_err)) if track_err else None}}
    if reference and rank == 0:
        out["reference_cpu"] = reference(W, H, NW, density)
        if out["reference_cpu"] and out["reference_cpu"].get("ms_per_frame"):
            out["speedup_vs_reference_cpu"] = round(out["reference_cpu"]["ms_per_frame"] / ms * world, 1)
    if os.environ.get("STREAM_HOST_TIMING"):
        print(ba.read("host_timing", np.uint8).tobytes().decode(), file=sys.stderr)
        print({k: [round(x, 2) for x in v] for k, v in T.items()}, file=sys.stderr)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return out


trk_ready = [False]

if __name__ == "__main__":
    main()
