"""CPU restatement of the reference's window-maintenance DECISIONS -- TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module; the product (libcml_b200/) never does.  Parity is pinned: tests/test_maintenance_oracle.py
checks every function against tests/golden/maint_golden.cmlw, produced by the unmodified reference (oracle/make_golden.py maintenance).

Reference (under /root/reference/src/cml/optimization/dso):
  flag_frames        DSOBundleAdjustment.cpp:603-708  flagFramesForMarginalization
  is_oob             DSOBundleAdjustment.cpp:2515-2554 isOOB
  try_marginalize    DSOBundleAdjustment.cpp:2240-2363 tryMarginalize
  remove counters    DSOContext.h:94-110 (removePoint), :207-218 (removeResiduals)
"""
import numpy as np

IN, OOB, OUTLIER = 0, 1, 2


def flag_frames(cams, keyid, aff_a, exposure, n_res, n_immature, n_marg, n_out, flagged, max_frames, min_age=1):
    """cams [N,12] world->cam (R row-major, t); returns the updated flag vector (BA:603-708)."""
    N = len(keyid)
    flagged = np.array(flagged, bool).copy()
    nflag = 0
    for i in range(N):
        inn = float(n_res[i] + n_immature[i]); out = float(n_marg[i] + n_out[i])
        ref_to_fh = np.exp(aff_a[i] - aff_a[N - 1]) * exposure[i] / exposure[N - 1]           # map/Exposure.h:119-123
        not_enough = inn < 0.05 * (inn + out)
        too_big = abs(np.log(ref_to_fh)) > 0.7 and N - nflag > max_frames - 2
        if not_enough or too_big:
            flagged[i] = True; nflag += 1
    if N - nflag >= max_frames:
        R = [np.asarray(c[:9]).reshape(3, 3) for c in cams]; t = [np.asarray(c[9:]) for c in cams]

        def rel_t(a, b):                                         # a.to(b) = b o a^-1 (map/Camera.h:289-300): translation
            return t[b] - R[b] @ R[a].T @ t[a]
        smallest, pick = 1.0, -1
        latest = keyid[N - 1]
        for r in range(N):
            if keyid[r] > latest - min_age or keyid[r] == 0:
                continue
            score = 0.0
            for k in range(N):
                if k == r or keyid[k] > latest - min_age + 1:
                    continue
                score += 1.0 / (1e-5 + np.linalg.norm(rel_t(r, k)))
            score *= -np.sqrt(np.linalg.norm(rel_t(r, N - 1)))
            if score < smallest:
                smallest, pick = score, r
        if pick >= 0:
            flagged[pick] = True
    return flagged


def is_oob(states, targets, flagged, num_good, last0, last1):
    """states/targets of the point's residuals (BA:2515-2554)."""
    num_in = int(np.sum(states == IN)); vis = int(np.sum((states == IN) & flagged[targets]))
    if num_in >= 3 and num_good > 4 + 10 and num_in - vis < 3:
        return True
    if last0 == OOB:
        return True
    if num_in < 2:
        return False
    return last0 == OUTLIER and last1 == OUTLIER


def try_marginalize(pt_host, pt_idepth, pt_num_good, pt_idh, pt_last0, pt_last1, res_point, res_target, res_state, flagged, alive, min_idh=50.0):
    """Returns (drop, marginalize) boolean vectors over points (BA:2240-2363)."""
    P = len(pt_host)
    drop = np.zeros(P, bool); marg = np.zeros(P, bool)
    order = np.argsort(res_point, kind="stable")
    rp, rt, rs = res_point[order], res_target[order], res_state[order]
    lo = np.searchsorted(rp, np.arange(P)); hi = np.searchsorted(rp, np.arange(P), side="right")
    for p in range(P):
        if not alive[p]:
            continue
        st, tg = rs[lo[p]:hi[p]], rt[lo[p]:hi[p]]
        if pt_idepth[p] < 0 or st.size == 0:
            drop[p] = True
        elif is_oob(st, tg, flagged, pt_num_good[p], pt_last0[p], pt_last1[p]) or flagged[pt_host[p]]:
            if st.size >= 3 and pt_num_good[p] >= 4 and pt_idh[p] > min_idh:
                marg[p] = True
            else:
                drop[p] = True
    return drop, marg


def counters_after_removal(n_marg, n_out, res_point, res_target, gone, marginalized):
    """numMarginalized / numResidualsOut after removePoint(point, marginalize) for every point in `gone` (DSOContext.h:94-110, 207-218)."""
    n_marg = np.array(n_marg).copy(); n_out = np.array(n_out).copy()
    for p, t in zip(res_point, res_target):
        if gone[p]:
            n_out[t] += 1
            if marginalized[p]:
                n_marg[t] += 1
    return n_marg, n_out
