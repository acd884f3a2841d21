import os, sys, subprocess, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from libcml_b200 import DSOBundleAdjustment, synth, cmlw
from parity_util import rel
REF = "/root/repo/oracle/_ref/cmlba_ref"
for cfg in sys.argv[1:]:
    win = synth.make_config(cfg)
    t0 = time.time()
    cmlw.save("/tmp/w.cmlw", {k: v for k, v in win.items() if k != "grad"})
    subprocess.run([REF, "--window", "/tmp/w.cmlw", "--mode", "run", "--out", "/tmp/o.cmlw"], check=True, capture_output=True, timeout=900)
    g = cmlw.load("/tmp/o.cmlw")
    t1 = time.time()
    ba = DSOBundleAdjustment(device=0)
    cams = ba.loadWindow(win)
    ok = ba.run(cams, iterations=int(win["iterations"][0]))
    r = ba.last_result
    fr = ba.getFrames(); pts = ba.getPoints(); rs = ba.getResiduals()
    mine = set(zip(rs["point_id"].tolist(), rs["target_frame_id"].tolist()))
    theirs = set(zip(g["fin_alive_res_point"].tolist(), g["fin_alive_res_target"].tolist()))
    print(cfg, "ok", ok, bool(g["fin_ok"][0]), "ref %.1fs" % (t1 - t0), "gpu_ms %.3f" % r.gpu_ms, "iters", r.iterations_done, "R", r.num_residuals,
          "pose", rel(fr["world_to_cam"], g["fin_frame_pre_w2c"]), "aff", np.abs(fr["affine"] - g["fin_frame_affine"]).max(),
          "idepth", rel(pts["idepth"], g["fin_pt_idepth"][pts["id"]]), "res diff", len(mine ^ theirs), "of", len(theirs), flush=True)
    ba.close()
