"""Development aid: per-warp SM-clock timeline of linearize_tile_kernel (CMLBA_LT_MODE bit 1 set).  python tools/lt_trace.py [workload]
Stamps per warp pass: start (headers arrived) | projection | tiles landed | taps | photometric sums | classification + Jacobians | stores."""
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
print("lanes that sampled:", int(tr[:, :, 28].sum()), " of them read the image instead of the staged tile:", int(tr[:, :, 29].sum()),
      " warps with at least one such lane:", int((tr[:, :, 29] > 0).sum()), "of", int((tr[:, :, 28] > 0).sum()))
tr = tr.copy(); tr[:, :, 28:30] = 0
ok = g0 > 0
gmin = g0[ok].min()
print(f"globaltimer (ns): first warp start 0, last warp start {g0[ok].max() - gmin}, first end {g1[ok].min() - gmin}, last end {g1[ok].max() - gmin}")
live = [c for c in range(tr.shape[0]) if ok[c].any()]
ends = np.array([g1[c][ok[c]].max() - gmin for c in live]); starts = np.array([g0[c][ok[c]].min() - gmin for c in live])
print("CTA start ns: min/median/max", starts.min(), np.median(starts), starts.max(), " CTA end ns: min/median/max", ends.min(), np.median(ends), ends.max())
# aggregate: median over all warps of every stamp, relative to the CTA's first stamp (us at 1965 MHz)
names = ["start", "proj", "tiles", "taps", "photo", "jac", "stores"]
rel = []
for c in live:
    v = tr[c, :, :28].astype(np.float64)
    t0 = v[v > 0].min()
    v = np.where(v > 0, (v - t0) / 1965.0, np.nan)
    rel.append(v)
rel = np.stack(rel)                                      # [cta][warp][stamp]
full = rel[:, :, 1:22]
pro = rel[:, :, 22:27]
print("prologue: median " + " ".join(f"{n}={x:6.2f}" for n, x in zip(["record", "pre-sync", "post-sync", "pairs", "prefetch"], np.nanmedian(pro.reshape(-1, 5), axis=0)))
      + "   max " + " ".join(f"{x:6.2f}" for x in np.nanmax(pro.reshape(-1, 5), axis=0)))
for ps in range(2):
    row = np.nanmedian(full[:, :, 7 * ps:7 * ps + 7].reshape(-1, 7), axis=0)
    mx = np.nanmax(full[:, :, 7 * ps:7 * ps + 7].reshape(-1, 7), axis=0)
    if np.isnan(row).all():
        continue
    print(f"pass {ps}: median " + " ".join(f"{n}={x:6.2f}" for n, x in zip(names, row)) + "   max " + " ".join(f"{x:6.2f}" for x in mx))
for cta in (live[0], live[len(live) // 2], live[int(np.argmax(ends))]):
    print(f"== CTA {cta}")
    for wv in range(16):
        v = rel[live.index(cta), wv]
        if not np.isnan(v).all():
            print(f"  warp {wv:2d}: " + " ".join((f"{x:6.2f}" if not np.isnan(x) else "     -") + (" |" if k % 7 == 0 else "") for k, x in enumerate(v[:22])))
