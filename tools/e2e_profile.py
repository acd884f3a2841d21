"""Development aid: wall-time breakdown of the end-to-end call sequence bench.py times (run on a GPU box).
python tools/e2e_profile.py [workload] [repeats]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from libcml_b200 import DSOBundleAdjustment, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W, H, N, ppk, iters, affine = synth.CONFIGS[wl]
win = synth.make_config(wl, seed=1234)
P = win["pt_host"].size
ba = DSOBundleAdjustment(device=0, iterations=iters, async_image_upload=1)
cams = win["frame_cam"]
gray = torch.from_numpy(np.ascontiguousarray(win["gray"], dtype=np.float32)).pin_memory().numpy()
acc = {}
def lap(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0) * 1e3; return t1
ids = np.arange(P)
for i in range(reps + 1):
    if i == 1:
        acc.clear(); ba.read("host_timing_reset", np.uint8)
    torch.cuda.synchronize()
    t = time.perf_counter(); t00 = t
    ba.reset(); ba.setCalibration(*[float(v) for v in win["calib"]], W, H); t = lap("py.reset+calib", t)
    for f in range(N):
        ba.addNewFrameGray(f, win["frame_evalpt"][f], win["frame_affine"][f, 0], win["frame_affine"][f, 1], win["frame_exposure"][f], gray[f], False)
    t = lap("py.addNewFrame x N", t)
    ba.addPoints(ids, win["pt_host"], win["pt_xy"], win["pt_idepth"]); t = lap("py.addPoints", t)
    ok = ba.run(cams, iterations=iters); t = lap("py.run", t)
    fr = ba.getFrames(); pts = ba.getPoints(); t = lap("py.getters", t)
    acc["py.total"] = acc.get("py.total", 0.0) + (t - t00) * 1e3
r = ba.last_result
print(f"workload {wl}: R={r.num_residuals} iterations={r.iterations_done} gpu_ms={r.gpu_ms:.3f} launches={r.kernel_launches}")
for k, v in acc.items():
    print(f"{k:28s} {v / reps:10.3f} ms/step")
print("---- engine host timers (totals over", reps, "steps)")
print(ba.read("host_timing", np.uint8).tobytes().decode())
