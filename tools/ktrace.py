"""Development aid: device timeline of one timed pass (CMLBA_KTRACE=1): when the first CTA of every kernel was scheduled, when it got past its
dependency wait, when its last warp finished (globaltimer, ns relative to the first stamp).   CMLBA_KTRACE=1 python tools/ktrace.py [workload] [run]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CMLBA_KTRACE", "1")
from libcml_b200 import DSOBundleAdjustment, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "pass"
names = ["(timed region starts)", "accumulate", "schur", "stitch", "assemble", "solve", "point_step", "post_linearize", "pack_post", "restore_state"]
win = synth.make_config(wl)
ba = DSOBundleAdjustment(device=0)
cams = ba.loadWindow(win)
if mode == "pass":
    ba.prepare(cams)
    for cold in (True, False):
        br = ba.benchPass(5, 3, cold)
        kt = ba.read("ktrace", np.uint64).reshape(-1, 3).astype(np.int64)
        live = [k for k in range(len(names)) if kt[k, 2] > 0]
        t0 = min(kt[k, 0] for k in live)
        print(f"{'cold' if cold else 'warm'} pass: {br.ms_pass * 1e3:.1f} us by events (timeline of the last timed pass below)")
        for k in sorted(live, key=lambda k: kt[k, 0]):
            print(f"  {names[k]:15s} scheduled {kt[k, 0] - t0:7d}  past wait {kt[k, 1] - t0:7d}  done {kt[k, 2] - t0:7d}   busy {kt[k, 2] - kt[k, 1]:6d} ns")
else:
    ba.run(cams, iterations=int(win["iterations"][0]))
    r = ba.last_result
    kt = ba.read("ktrace", np.uint64).reshape(-1, 3).astype(np.int64)
    live = [k for k in range(len(names)) if kt[k, 2] > 0]
    t0 = min(kt[k, 0] for k in live)
    print(f"run: gpu_ms {r.gpu_ms:.3f}, {r.kernel_launches} launches, {r.iterations_done} iterations; first schedule / last completion per kernel type")
    for k in sorted(live, key=lambda k: kt[k, 0]):
        print(f"  {names[k]:15s} first scheduled {kt[k, 0] - t0:7d}  last done {kt[k, 2] - t0:7d}")
