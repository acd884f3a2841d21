"""Development aid: stage-by-stage comparison of the CUDA path against the committed golden vectors.
Run on a GPU box:  python tools/gpu_check.py [tiny|tiny_affine]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from libcml_b200 import DSOBundleAdjustment
from parity_util import *

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
win, g = load_golden(name)
N = win["frame_evalpt"].shape[0]; n = 8 * N + 4
ba = DSOBundleAdjustment()
ba.enableDebugDump()
cams = ba.loadWindow(win)
ba.prepare(cams)
dv = DeviceView(ba, win)
fr = dv.frames()
print("state", rel(fr["state"], g["pre_frame_state"]), "preR", rel(fr["preR"], g["pre_frame_pre_w2c"][:, :9]), "pret", rel(fr["pret"], g["pre_frame_pre_w2c"][:, 9:]))
AH = ba.read("AH", np.float64).reshape(N, N, 8, 8); AT = ba.read("AT", np.float64).reshape(N, N, 8, 8)
gAH = g["pre_ad_host"].reshape(N, N, 8, 8).transpose(1, 0, 2, 3); gAT = g["pre_ad_target"].reshape(N, N, 8, 8).transpose(1, 0, 2, 3)   # ref index h + N*t
print("AH", rel(AH, gAH), "AT", rel(AT, gAT))
print("colors", rel(dv.point_array("pt_colors", np.float32, 8), g["pre_pt_colors"]), "weights", rel(dv.point_array("pt_weights", np.float32, 8), g["pre_pt_weights"]))

def check_lin(stage):
    pre = stage + "_"
    m = dv.map_to(g, pre)
    ns = ba.read("res_new_state", np.uint8)[m]; ne = ba.read("res_new_energy", np.float32)[m]; neo = ba.read("res_new_energy_wo", np.float32)[m]
    print(stage, "new_state mismatch", int((ns != g[pre + "res_new_state"]).sum()), "energy", rel(ne, g[pre + "res_new_energy"]), "wo", rel(neo, g[pre + "res_new_energy_wo"]))
    th = dv.frames()["energy_th"]; print("   th", th, g[pre + "frame_energy_th"])
    if pre + "rJ_resF" in g:
        J = decode_rj(ba.read("rj", np.float32), ba.read("dbg", np.float32))
        ok = g[pre + "res_new_state"] == 0
        for k in ["resF", "Jpdxi", "Jpdc", "Jpdd", "JIdx", "JabF", "JIdx2", "JabJIdx", "Jab2"]:
            print("   ", k, rel(J[k][m][ok], g[pre + "rJ_" + k][ok]))
        okc = g[pre + "res_new_state"] != 1
        print("    center", rel(ba.read("res_center", np.float32).reshape(-1, 3)[m][okc], g[pre + "res_center"][okc]))

E0 = ba.linearizeAll(False)
print("E0", E0, g["lin0_energy"])
check_lin("lin0")
ba.applyActiveRes()
m = dv.map_to(g, "app0_")
T = ba.read("T", np.float32).reshape(dv.P, N, 16)
Tc = np.zeros_like(T); Tc[dv.pt_order] = T
jp = Tc[g["app0_res_point"], g["app0_res_target"], :8]
print("app0 good mismatch", int((ba.read("res_good", np.uint8)[m] != g["app0_res_good"]).sum()), "JpJdF", rel(jp, g["app0_res_JpJdF"]),
      "state", int((ba.read("res_state", np.uint8)[m] != g["app0_res_state"]).sum()), "energy", rel(ba.read("res_energy", np.float32)[m], g["app0_res_energy"]))
it = 0
while f"sol{it}_x" in g:
    pre = f"sol{it}_"
    ba.solveSystem(it)
    acc = unpack_acc(ba.read("acc", np.float64), N)
    print(pre, "acc", rel(acc, g[pre + "acc_active"]))
    s = split_sys(ba.read("sys", np.float64), n)
    print("   HA", rel(s["HA"][4:, 4:], g[pre + "HA_top"][4:, 4:]), "HAfull", rel(s["HA"], g[pre + "HA_top"]), "bA", rel(s["bA"], g[pre + "bA_top"][:, 0]),
          "Hsc", rel(s["HS"], g[pre + "H_sc"]), "bsc", rel(s["bS"], g[pre + "b_sc"][:, 0]))
    for k, gk in [("pt_Hdd", "pt_Hdd"), ("pt_bd", "pt_bd"), ("pt_HdiF", "pt_HdiF"), ("pt_bdSumF", "pt_bdSumF"), ("pt_idepth_hessian", "pt_idepth_hessian")]:
        print("   ", k, rel(dv.point_array(k, np.float32), g[pre + gk]), end="")
    print("   Hcd", rel(dv.point_array("pt_Hcd", np.float32, 4), g[pre + "pt_Hcd"]))
    x = ba.read("x", np.float64)
    H, b = solve_reference_system(g, pre, N)
    xr = g[pre + "x"]
    print("   x fwd", rel(x, xr), "backward err (mine)", np.abs(H[4:, 4:] @ x[4:] - b[4:]).max() / np.abs(b[4:]).max(), "(ref)", np.abs(np.tril(H[4:, 4:]) @ xr[4:] + np.tril(H[4:, 4:], -1).T @ xr[4:] - b[4:]).max() / np.abs(b[4:]).max())
    print("   pt_step", rel(dv.point_array("pt_step", np.float64), g[pre + "pt_step"]))
    fr = dv.frames()
    print("   frame state", rel(fr["state"], g[f"step{it}_frame_state"]), "idepth", rel(dv.point_array("pt_idepth", np.float64), g[f"step{it}_pt_idepth"]),
          "canbreak", ba.doStepFromBackup(), g[f"step{it}_canbreak"])
    E = ba.linearizeAll(False)
    print("   E", E, g[f"lin{it+1}_energy"])
    check_lin(f"lin{it+1}")
    ba.applyActiveRes()
    it += 1
E = ba.linearizeAll(True)
print("fin E", E, g["fin_energy"], "dropped", dv.ctrl()["num_dropped"])
fr = dv.frames()
print("fin state", rel(fr["state"], g["fin_frame_state"]), "evalR", rel(fr["evalR"], g["fin_frame_evalpt"][:, :9]), "numgood", int((dv.point_array("pt_num_good", np.int32) != g["fin_pt_num_good"]).sum()),
      "mrb", rel(dv.point_array("pt_max_rel_baseline", np.float32), g["fin_pt_max_rel_baseline"]))
# full run() on a fresh handle
ba2 = DSOBundleAdjustment()
cams = ba2.loadWindow(win)
ok = ba2.run(cams, iterations=int(win["iterations"][0]))
r = ba2.last_result
print("run ok", ok, "iters", r.iterations_done, g["iterations_done"], "E first/last", r.energy_first, r.energy_last, g["lin0_energy"], g["fin_energy"], "dropped", r.num_dropped, "outliers", r.num_outliers, "gpu_ms", r.gpu_ms, "launches", r.kernel_launches)
f2 = ba2.getFrames(); p2 = ba2.getPoints()
print("run w2c", rel(f2["world_to_cam"], g["fin_frame_pre_w2c"]), "state", rel(f2["state"], g["fin_frame_state"]), "aff", np.abs(f2["affine"] - g["fin_frame_affine"]).max(), "th", f2["energy_th"], g["fin_frame_energy_th"])
alive = g["fin_pt_alive"].astype(bool)
print("run idepth", rel(p2["idepth"], g["fin_pt_idepth"][p2["id"]]), "npts", p2["id"].size, alive.sum(), "gft", int((np.sort(p2["id"][p2["good_for_tracking"] != 0]) != g["fin_good_points_for_tracking"]).sum()) if (p2["good_for_tracking"] != 0).sum() == g["fin_good_points_for_tracking"].size else "size-mismatch")
rs = ba2.getResiduals()
print("run residuals", rs["point_id"].size, g["fin_alive_res_point"].size)
