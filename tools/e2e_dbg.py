import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
from libcml_b200 import DSOBundleAdjustment, synth
W, H, N, ppk, iters, affine = synth.CONFIGS["c2"]
win = synth.make_config("c2", seed=1234)
P = win["pt_host"].size
ba = DSOBundleAdjustment(device=0, iterations=iters, async_image_upload=1)
cams = ba.loadWindow(win)
mode = sys.argv[1]
if "prep" in mode: ba.prepare(cams)
if "pass" in mode:
    ba.prepare(cams); ba.benchPass(50, 5, True); ba.benchPass(50, 5, False)
gnp = torch.from_numpy(win["grad"]).pin_memory().numpy()
ts = []
for i in range(21):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ba.reset()
    ba.setCalibration(*[float(v) for v in win["calib"]], W, H)
    for f in range(N):
        ba.addNewFrame(f, win["frame_evalpt"][f], win["frame_affine"][f, 0], win["frame_affine"][f, 1], win["frame_exposure"][f], gnp[f], False)
    ba.addPoints(np.arange(P), win["pt_host"], win["pt_xy"], win["pt_idepth"])
    ok = ba.run(cams, iterations=iters)
    fr = ba.getFrames(); pts = ba.getPoints()
    torch.cuda.synchronize()
    if i: ts.append(time.perf_counter() - t0)
print(mode, "e2e ms", 1e3 * np.mean(ts), "min", 1e3 * np.min(ts))
print(ba.read("host_timing", np.uint8).tobytes().decode())
