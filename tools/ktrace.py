"""Development aid: device timeline (CMLBA_KTRACE=1) of the last timed pass of cmlba_bench_pass, or of a whole run(): per launch, when its first
CTA was scheduled, when the first CTA got past its dependency wait, when its last warp finished (globaltimer, ns after the reset kernel that opens
the region).  The sampling kernel carries no stamps (register budget): it ends where the next kernel gets past its wait.
    CMLBA_KTRACE=1 python tools/ktrace.py [workload] [pass|run]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CMLBA_KTRACE", "1")
from libcml_b200 import DSOBundleAdjustment, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "pass"
SITES = {0: "(region starts)", 1: "linearize_tile", 2: "post_linearize/pack_post", 4: "accumulate", 8: "schur", 16: "stitch_pair", 32: "assemble", 64: "solve", 128: "point_step/restore_state"}


def show(ba, title):
    kt = ba.read("ktrace", np.uint64).reshape(-1, 4).astype(np.int64)
    t0 = kt[0, 0]
    print(title)
    prev_done = 0
    for i in range(kt.shape[0]):
        sched, wait, done, site = kt[i]
        name = SITES.get(int(site), str(site))
        if done <= 0:
            print(f"  {i:3d} {name:26s} (no stamps)")
            continue
        print(f"  {i:3d} {name:26s} scheduled {sched - t0:8d}  past wait {wait - t0:8d}  done {done - t0:8d}   busy {done - wait:6d}   idle before {wait - t0 - prev_done:6d} ns")
        prev_done = max(prev_done, done - t0)


win = synth.make_config(wl)
ba = DSOBundleAdjustment(device=0)
cams = ba.loadWindow(win)
if mode == "pass":
    ba.prepare(cams)
    for cold in (True, False):
        br = ba.benchPass(5, 3, cold)
        show(ba, f"{'cold' if cold else 'warm'} pass: {br.ms_pass * 1e3:.1f} us by events; timeline of the last timed pass:")
else:
    ba.run(cams, iterations=int(win["iterations"][0]))
    r = ba.last_result
    show(ba, f"run: gpu_ms {r.gpu_ms:.3f}, {r.kernel_launches} launches, {r.iterations_done} iterations")
    st = ba.read("ktrace_solve", np.uint64).astype(np.int64)[:11]
    names = ["enter", "system loaded", "damped", "scaled", "factorised+forward", "backward", "orthogonalised", "x stored", "frame states", "xAd", "pair constants"]
    print("last solve_kernel, ns since its entry: " + ", ".join(f"{n} {st[k] - st[0]}" for k, n in enumerate(names) if st[k] > 0))
