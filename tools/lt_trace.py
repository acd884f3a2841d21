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
br = ba.benchPass(3, 2, os.environ.get("LT_COLD", "1") == "1")
print("bench: pass", br.ms_pass * 1e3, "us, linearize", br.ms_linearize * 1e3, "us")
tr = ba.read("lt_trace", np.int64).reshape(-1, 16, 32)
g0 = tr[:, :, 30]; g1 = tr[:, :, 31]
ok = g0 > 0
gmin = g0[ok].min()
print(f"globaltimer (ns): first warp start 0, last warp start {g0[ok].max() - gmin}, first end {g1[ok].min() - gmin}, last end {g1[ok].max() - gmin}")
ends = np.array([g1[c][ok[c]].max() - gmin for c in range(tr.shape[0]) if ok[c].any()]); starts = np.array([g0[c][ok[c]].min() - gmin for c in range(tr.shape[0]) if ok[c].any()])
print("CTA start ns: min/median/max", starts.min(), np.median(starts), starts.max(), " CTA end ns: min/median/max", ends.min(), np.median(ends), ends.max())
print("slowest CTAs:", np.argsort(-ends)[:6], np.sort(-ends)[:6] * -1)
for cta in (0, 73, int(np.argmax(ends))):
    t0 = tr[cta, :, :30][tr[cta, :, :30] > 0].min()
    print(f"== CTA {cta}")
    for wv in range(16):
        v = tr[cta, wv, :30]
        if (v > 0).any():
            print(f"  warp {wv:2d}: " + " ".join((f"{(x - t0) / 1965.0:6.2f}" if x > 0 else "     -") + (" |" if k % 7 == 0 else "") for k, x in enumerate(v[:22])))
