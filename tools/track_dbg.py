import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_tracker import load, make_tracker
import tracker_oracle as T
win, g = load()
np.set_printoptions(precision=9, linewidth=200)
for cluster in (1, 8):
    trk, ref, new = make_tracker(win, cluster_ctas=cluster)
    for case in ("a", "b"):
        r = trk.optimize(g[f"{case}_init_cam"], g[f"{case}_new_affine"], gray=win["gray"][new], exposure_time=win["frame_exposure"][new])
        print(cluster, case, "it", r.iterations, "nT", r.numTermsInE, g[f"{case}_trk_numTermsInE"], "E", r.E, g[f"{case}_trk_E"])
        print("   cam diff", np.abs(r.camera - g[f"{case}_trk_cam"]).max(), "aff", r.exposure, g[f"{case}_trk_affine"], "rep", r.levelCutoffRepeat, "flow", r.flowVector, g[f"{case}_trk_flow"])
# oracle on the device point clouds
L = 5
pcs = [trk.read(f"pc{l}", np.float32).reshape(-1, 4) for l in range(L)]
pyr = T.build_pyramid(win["gray"][new], L)
Rr, tr = win["frame_cam"][ref][:9].reshape(3, 3), win["frame_cam"][ref][9:]
st = g["a_init_cam"]; Rn, tn = st[:9].reshape(3, 3), st[9:]
R0 = Rn @ Rr.T
TR = []
o = T.optimize(pcs, [p[1] for p in pyr], [T.level_K(win["calib"], l) for l in range(L)], (R0, tn - R0 @ tr),
               (win["frame_exposure"][ref], win["frame_affine"][ref, 0], win["frame_affine"][ref, 1]), (win["frame_exposure"][new], 0.0, 0.0), params=dict(trace=TR))
print("oracle it", o["iterations"], "nT", o["numTermsInE"], "E", o["E"])
for t in TR: print(t)
