"""CPU, world_size 2 over gloo: the multi-GPU scheme of SURVEY.md 8(e) / DESIGN.md section 5, executed with the numpy oracle.

Points (with all their residuals) are sharded across ranks, frames and images replicated.  Checked against the
single-process result computed on the same window:
  * the reduced system [H_A | b_A | H_sc | b_sc] is the SUM of the ranks' stitched partial systems (stitching is linear
    in the accumulators) -> one all-reduce per GN iteration;
  * the energy and the 0.7-quantile frameEnergyTH of the newest frame (BA:2419-2464) come out identical when every rank
    contributes its candidate energies to one all-gather (what pack_post_kernel + post_linearize_kernel do on the device)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_q):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import ba_oracle as O
    from libcml_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    win = synth.make_window(W=160, H=120, N=4, pts_per_kf=50, iterations=2, affine=True, seed=3)
    P = win["pt_host"].size
    sel = np.arange(P)[np.arange(P) % world == rank]
    shard = dict(win); shard["pt_host"] = win["pt_host"][sel]; shard["pt_xy"] = win["pt_xy"][sel]; shard["pt_idepth"] = win["pt_idepth"][sel]

    def linearize(w):
        O.compute_adjoints(w); O.compute_delta(w)
        e = O.linearize_all(w, False)
        act = np.nonzero(w.res_alive)[0]
        cand = w.res_new_energy_wo[act][(w.res_new_energy_wo[act] >= 0) & (w.res_target[act] == w.N - 1)].astype(np.float32)
        return e, cand

    def system(w):
        O.apply_active_res(w)
        acc = O.accumulate_top(w)
        HA, bA = O.stitch_top(w, acc, False)
        O.accumulate_sc(w, True)
        Hsc, bsc = O.stitch_sc(w)
        return np.concatenate([HA.ravel(), bA.ravel(), Hsc.ravel(), bsc.ravel()])

    def threshold(vals):   # setNewFrameEnergyTH (BA:2440-2461) on a candidate list
        F32 = np.float32
        nth = int(F32(0.7) * F32(vals.size))
        e = np.partition(vals, nth)[nth]
        th = F32(np.sqrt(e)) * F32(1.5)
        th = F32(26.0) * F32(0.5) + th * F32(0.5)
        return float(F32(th * th))

    # ---- sharded
    ws = O.Window(shard)
    e_loc, cand_loc = linearize(ws)
    gathered = [None] * world
    dist.all_gather_object(gathered, (float(e_loc), cand_loc))
    e_glob = sum(g[0] for g in gathered)                       # fixed rank order
    th_glob = threshold(np.concatenate([g[1] for g in gathered]))
    ws.frame_energy_th[ws.N - 1] = th_glob                     # every rank installs the global threshold
    sys_loc = torch.from_numpy(system(ws))
    dist.all_reduce(sys_loc)                                   # the one collective of a GN iteration
    # ---- single process on the full window
    w1 = O.Window(win)
    e_one, cand_one = linearize(w1)
    th_one = w1.frame_energy_th[w1.N - 1]
    sys_one = system(w1)
    out_q.put((rank, e_glob, e_one, th_glob, th_one, float(np.abs(sys_loc.numpy() - sys_one).max() / np.abs(sys_one).max()),
               int(np.concatenate([g[1] for g in gathered]).size), int(cand_one.size)))
    dist.destroy_process_group()


def test_sharded_points_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_glob, e_one, th_glob, th_one, sys_err, n_g, n_1 in res:
        assert n_g == n_1                                           # every candidate of the newest frame is owned by exactly one rank
        assert th_glob == th_one, (th_glob, th_one)                 # the global quantile is exact (same multiset of candidates)
        assert abs(e_glob - e_one) <= 1e-9 * abs(e_one)             # energy: sum of the ranks' sums
        assert sys_err < 1e-5, sys_err                              # fp32 accumulators summed in a different order
