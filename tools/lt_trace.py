"""Development aid: per-warp SM-clock timeline of linearize_tile_kernel (CMLBA_LT_MODE=2 or 3).  python tools/lt_trace.py [workload]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libcml_b200 import DSOBundleAdjustment, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
win = synth.make_config(wl)
ba = DSOBundleAdjustment(device=0)
cams = ba.loadWindow(win)
ba.prepare(cams)
ba.benchPass(3, 2, os.environ.get("LT_COLD", "1") == "1")
tr = ba.read("lt_trace", np.int64).reshape(-1, 16, 32)
for cta in (0, 1, 73, 147):
    if cta >= tr.shape[0]:
        continue
    t0 = tr[cta][tr[cta] > 0].min()
    print(f"== CTA {cta}")
    for wv in range(16):
        v = tr[cta, wv]; v = v[v > 0]
        if v.size:
            print(f"  warp {wv:2d}: " + " ".join(f"{(x - t0) / 1965.0:6.2f}" for x in v))
