"""Development aid: the large-motion edge scenario of tests/test_gpu_edge.py, first linearization and full run against the numpy oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libcml_b200 import DSOBundleAdjustment, synth
import ba_oracle as O
from parity_util import DeviceView, map_residuals, rel
win = synth.make_window(W=120, H=90, N=5, pts_per_kf=80, iterations=3, affine=False, seed=5, pose_noise=2e-3)
win["frame_cam"] = win["frame_cam"].copy(); win["frame_cam"][:, 9] += np.linspace(0, 0.25, 5); win["frame_evalpt"] = win["frame_cam"].copy()
ba = DSOBundleAdjustment(device=0)
cams = ba.loadWindow(win)
ba.prepare(cams)
dv = DeviceView(ba, win)
E0 = ba.linearizeAll(False)
ow = O.Window(win)
O.prepare(ow) if hasattr(O, "prepare") else None
print("E0 gpu", E0)
ns = ba.read("res_new_state", np.uint8); ne = ba.read("res_new_energy_wo", np.float32); ctr = ba.read("res_center", np.float32).reshape(-1, 3)
np.savez(f"gpurun_out/dbg_lm_{os.environ.get('CMLBA_LT_EXACT', '0')}.npz", point=dv.res_point, target=dv.res_target, ns=ns, ne=ne, ctr=ctr)
ba.close()
ba = DSOBundleAdjustment(device=0)
cams = ba.loadWindow(win)
ok = ba.run(cams, iterations=3)
ow = O.Window(win); ok2 = O.run(ow)
rs = ba.getResiduals()
mine = set(zip(rs["point_id"].tolist(), rs["target_frame_id"].tolist()))
theirs = set(zip(ow.res_point[ow.res_alive].tolist(), ow.res_target[ow.res_alive].tolist()))
print("run ok", ok, ok2, "alive mine/theirs", len(mine), len(theirs), "sym diff", sorted(mine ^ theirs))
print("energy", ba.last_result.energy_first, ba.last_result.energy_last, "oracle", getattr(ow, "fin_energy", None))
