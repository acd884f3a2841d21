// tracer.cuh -- device code of the immature-point tracer (SURVEY.md 8f NEXT #2).
//
// Reference anchors (under /root/reference/src/cml/optimization/dso):
//   trc_init_kernel      DSOTracer.cpp:538-583 (makeNewTracesFrom), DSOTracer.h:14-32 (point defaults)
//   trc_trace_kernel     DSOTracer.cpp:585-832 (trace), called for every immature point by traceNewCoarse (:17-60)
//   trc_activate_kernel  DSOTracer.cpp:280-411 (optimizeImmaturePoint), :413-494 (linearizeResidual)
//
// B200 design.  Both loops of the reference are per-point and independent; the work per point is a few hundred bilinear samples.
// One WARP per point: in trace the lanes take the steps of the epipolar search (32 at a time), in activation the lanes take
// (target frame, pattern pixel) pairs.  Everything that decides an outcome follows the reference's arithmetic -- fp64 geometry,
// the fp32 running position `ptx += dx`, fp32 running sums of Hdd / bd / energy in the reference's order (the terms travel by
// shuffle to a sequential fold, every lane folding redundantly) -- so statuses and return codes match exactly and the values
// to ~1e-6.  Points, their state and the frames' images stay resident on the device between calls; a call uploads only poses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmltrc {

constexpr int TRC_MAXF = 16;
enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };
enum { RES_IN = 0, RES_OOB = 1, RES_OUTLIER = 2 };

struct TrcParams {
    int W, H;
    double fx, fy, cx, cy;
    float huber, outlier_th, outlier_th_sum, max_pix_search, max_slack_interval, step_size, min_improvement, test_radius, extra_slack, min_idepth_h_act;
    int gn_iterations;
};

struct PointsDev {           // SoA over immature points (index = point id)
    int *host;               // frame slot of the host, -1 = removed
    float2 *xy;
    int *status;
    double *idmin, *idmax, *u, *v, *interval, *quality;
    double *gradH;           // [P][4]
    double *energyTH;
    float *weights;          // [P][8] (kept for parity dumps; the reference stores but never reads them in this path)
};

struct FrameImg { const float *gray; const float4 *grad; };

struct PairDev {             // host slot -> target: everything fp64 like the reference
    double KRKi[9], Kt[3];   // trace
    double R[9], t[3];       // activation (hostToTarget)
    double a, b;             // exposure transition host -> target
};

__constant__ int c_pat[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// level-0 derivative texels (I, dx, dy, 0); border ring zero like gradientImage
__global__ void __launch_bounds__(256) trc_grad_kernel(const float *__restrict__ g, float4 *__restrict__ out, const int w, const int h) {
    for (int i = blockIdx.x * 256 + threadIdx.x; i < w * h; i += gridDim.x * 256) {
        const int x = i % w, y = i / w;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x > 0 && y > 0 && x < w - 1 && y < h - 1) { o.x = g[i]; o.y = (g[i + 1] - g[i - 1]) * 0.5f; o.z = (g[i + w] - g[i - w]) * 0.5f; }
        out[i] = o;
    }
}

// Array2D::interpolate on a float image: m00 w00 + m10 w10 + m01 w01 + m11 w11 in fp32, left to right
__device__ __forceinline__ float interp_gray(const float *__restrict__ img, const int w, const float x, const float y) {
    const int ix = (int) x, iy = (int) y;
    const float dx = x - (float) ix, dy = y - (float) iy, dxdy = dx * dy;
    const float *p = img + (size_t) iy * w + ix;
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p[0], 1.f - dx - dy + dxdy), __fmul_rn(p[1], dx - dxdy)), __fmul_rn(p[w], dy - dxdy)), __fmul_rn(p[w + 1], dxdy));
}
__device__ __forceinline__ float4 interp_grad(const float4 *__restrict__ img, const int w, const float x, const float y) {
    const int ix = (int) x, iy = (int) y;
    const float dx = x - (float) ix, dy = y - (float) iy, dxdy = dx * dy;
    const float4 *p = img + (size_t) iy * w + ix;
    const float4 a = p[0], b = p[1], c = p[w], d = p[w + 1];
    const float w00 = 1.f - dx - dy + dxdy, w10 = dx - dxdy, w01 = dy - dxdy;
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, w00), __fmul_rn(b.x, w10)), __fmul_rn(c.x, w01)), __fmul_rn(d.x, dxdy));
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.y, w00), __fmul_rn(b.y, w10)), __fmul_rn(c.y, w01)), __fmul_rn(d.y, dxdy));
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.z, w00), __fmul_rn(b.z, w10)), __fmul_rn(c.z, w01)), __fmul_rn(d.z, dxdy));
    o.w = 0.f;
    return o;
}

// makeNewTracesFrom: one thread per new point
__global__ void __launch_bounds__(128) trc_init_kernel(const TrcParams P, const PointsDev pts, const int first, const int count, const int slot, const FrameImg img) {
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k >= count) return;
    const int p = first + k;
    const float2 xy = pts.xy[p];
    double g00 = 0, g01 = 0, g11 = 0;
    for (int i = 0; i < 8; i++) {
        // Vector2f(corner + shift): the sum is formed in fp64 and cast to fp32 by the call
        const float4 g = interp_grad(img.grad, P.W, (float) ((double) xy.x + c_pat[i][0]), (float) ((double) xy.y + c_pat[i][1]));
        const double gx = g.y, gy = g.z;
        g00 += gx * gx; g01 += gx * gy; g11 += gy * gy;
        pts.weights[(size_t) p * 8 + i] = (float) sqrt((double) P.outlier_th_sum / ((double) P.outlier_th_sum + (gx * gx + gy * gy)));
    }
    pts.gradH[(size_t) p * 4] = g00; pts.gradH[(size_t) p * 4 + 1] = g01; pts.gradH[(size_t) p * 4 + 2] = g01; pts.gradH[(size_t) p * 4 + 3] = g11;
    pts.energyTH[p] = 8.0 * (double) P.outlier_th;
    pts.host[p] = slot;
    pts.status[p] = IPS_UNINITIALIZED;
    pts.idmin[p] = 1.0 / 1000.0; pts.idmax[p] = nan("");
    pts.u[p] = -1.0; pts.v[p] = -1.0; pts.interval[p] = -1.0; pts.quality[p] = 10000.0;
}

__device__ __forceinline__ bool inside(const TrcParams &P, const double x, const double y, const double pad) {
    return x >= pad && y >= pad && x < (double) P.W - pad && y < (double) P.H - pad;
}

// trace(): one warp per immature point, lanes = steps of the search.  pairs[host slot] -> the frame being traced.
__global__ void __launch_bounds__(128) trc_trace_kernel(const TrcParams P, const PointsDev pts, const int num_points, const int target_slot, const PairDev *__restrict__ pairs,
                                                       const FrameImg *__restrict__ frames, int *__restrict__ counts /* [6] status histogram of this pass */) {
    const int p = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= num_points) return;
    const int hs = pts.host[p];
    if (hs < 0 || hs == target_slot) return;          // removed, or frame == referenceFrame (returns lastTraceStatus untouched)
    const int last = pts.status[p];
    auto leave = [&](const int st, const double u, const double v, const double interval) {
        if (lane == 0) { pts.status[p] = st; pts.u[p] = u; pts.v[p] = v; pts.interval[p] = interval; atomicAdd(counts + st, 1); }
    };
    if (last == IPS_OOB) { if (lane == 0) atomicAdd(counts + IPS_OOB, 1); return; }
    const PairDev &pr_ = pairs[hs];
    const float2 xy = pts.xy[p];
    const double X = xy.x, Y = xy.y;
    const double pr0 = pr_.KRKi[0] * X + pr_.KRKi[1] * Y + pr_.KRKi[2], pr1 = pr_.KRKi[3] * X + pr_.KRKi[4] * Y + pr_.KRKi[5], pr2 = pr_.KRKi[6] * X + pr_.KRKi[7] * Y + pr_.KRKi[8];
    const double maxPix = (double) (P.W + P.H) * (double) P.max_pix_search;
    const double idmin = pts.idmin[p], idmax = pts.idmax[p];
    const double m0 = pr0 + pr_.Kt[0] * idmin, m1 = pr1 + pr_.Kt[1] * idmin, m2 = pr2 + pr_.Kt[2] * idmin;
    const double minx = m0 / m2, miny = m1 / m2;
    if (!inside(P, minx, miny, 4.0)) { leave(IPS_OOB, -1.0, -1.0, 0.0); return; }
    double maxx, maxy, interval;
    if (isfinite(idmax)) {
        const double a0 = pr0 + pr_.Kt[0] * idmax, a1 = pr1 + pr_.Kt[1] * idmax, a2 = pr2 + pr_.Kt[2] * idmax;
        maxx = a0 / a2; maxy = a1 / a2;
        if (!inside(P, maxx, maxy, 5.0)) { leave(IPS_OOB, -1.0, -1.0, 0.0); return; }
        interval = sqrt((maxx - minx) * (maxx - minx) + (maxy - miny) * (maxy - miny));
        if (interval < (double) P.max_slack_interval) { leave(IPS_SKIPPED, (maxx + minx) / 2.0, (maxy + miny) / 2.0, interval); return; }
    } else {
        interval = maxPix;
        const double a0 = pr0 + pr_.Kt[0] * 0.01, a1 = pr1 + pr_.Kt[1] * 0.01, a2 = pr2 + pr_.Kt[2] * 0.01;
        const double dxx = a0 / a2 - minx, dyy = a1 / a2 - miny;
        const double inv = 1.0 / sqrt(dxx * dxx + dyy * dyy);
        maxx = minx + interval * dxx * inv; maxy = miny + interval * dyy * inv;
        if (!inside(P, maxx, maxy, 5.0)) { leave(IPS_OOB, -1.0, -1.0, 0.0); return; }
    }
    if (!(idmin < 0 || (m2 > 0.75 && m2 < 1.5))) { leave(IPS_OOB, -1.0, -1.0, 0.0); return; }
    double dx = (double) P.step_size * (maxx - minx), dy = (double) P.step_size * (maxy - miny);
    const double *G = pts.gradH + (size_t) p * 4;
    const double a = dx * (G[0] * dx + G[1] * dy) + dy * (G[2] * dx + G[3] * dy);
    const double b = dy * (G[0] * dy - G[1] * dx) - dx * (G[2] * dy - G[3] * dx);
    double errPx = (double) 0.2f + (double) 0.2f * (a + b) / a;
    if (errPx * (double) P.min_improvement > interval && isfinite(idmax)) { leave(IPS_BADCONDITION, (maxx + minx) / 2.0, (maxy + miny) / 2.0, interval); return; }
    if (errPx > 10) errPx = 10;
    dx /= interval; dy /= interval;
    if (interval > maxPix) interval = maxPix;
    int numSteps = (int) ((double) 1.9999f + interval / (double) P.step_size);
    const double randShift = minx * 1000 - floor(minx * 1000);
    float ptx = (float) (minx - randShift * dx), pty = (float) (miny - randShift * dy);
    if (!isfinite(dx) || !isfinite(dy)) { leave(IPS_OOB, -1.0, -1.0, 0.0); return; }
    if (numSteps >= 100) numSteps = 99;
    // reference colours of the 8 pattern pixels (getGrayPatch: integer pixel of the corner + offset), brightness transferred
    const FrameImg hf = frames[hs], tf = frames[target_slot];
    const int ixc = (int) xy.x, iyc = (int) xy.y;
    double refc[8], rx[8], ry[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        refc[k] = pr_.a * (double) hf.gray[(size_t) (iyc + c_pat[k][1]) * P.W + ixc + c_pat[k][0]] + pr_.b;
        rx[k] = pr_.KRKi[0] * c_pat[k][0] + pr_.KRKi[1] * c_pat[k][1];
        ry[k] = pr_.KRKi[3] * c_pat[k][0] + pr_.KRKi[4] * c_pat[k][1];
    }
    const double huber = (double) P.huber;
    double err[4] = {1e300, 1e300, 1e300, 1e300};       // this lane's steps: lane, lane + 32, lane + 64, lane + 96
    double bestE = 1e10, bestU = 0, bestV = 0;
    int bestI = -1;
    // the search position is the reference's fp32 running sum `ptx += dx`: every lane walks the whole (cheap) recurrence and keeps
    // the positions of its own steps, then all lanes evaluate their steps together
    float sx[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < numSteps; i++) {
        const bool mine = (i & 31) == lane;
#pragma unroll
        for (int j = 0; j < 4; j++) if (mine && (i >> 5) == j) { sx[j] = ptx; sy[j] = pty; }
        ptx = (float) ((double) ptx + dx); pty = (float) ((double) pty + dy);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = lane + 32 * j;
        if (i < numSteps) {
            double e = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double qx = (double) sx[j] + rx[k], qy = (double) sy[j] + ry[k];
                if (!inside(P, qx, qy, 3.0)) { e += 1e5; continue; }
                const double r = (double) interp_gray(tf.gray, P.W, (float) qx, (float) qy) - refc[k];
                const double hw = fabs(r) < huber ? 1.0 : huber / fabs(r);
                e += hw * r * r * (2 - hw);
            }
            err[j] = e;
            if (e < bestE) { bestE = e; bestU = sx[j]; bestV = sy[j]; bestI = i; }
        }
    }
    // first minimum over all steps = min energy, ties to the smaller index
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double oE = __shfl_xor_sync(0xffffffffu, bestE, off), oU = __shfl_xor_sync(0xffffffffu, bestU, off), oV = __shfl_xor_sync(0xffffffffu, bestV, off);
        const int oI = __shfl_xor_sync(0xffffffffu, bestI, off);
        const bool take = (oI >= 0) && (bestI < 0 || oE < bestE || (oE == bestE && oI < bestI));
        if (take) { bestE = oE; bestU = oU; bestV = oV; bestI = oI; }
    }
    double second = 1e10;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = lane + 32 * j;
        if (i < numSteps && ((double) i < (double) bestI - (double) P.test_radius || (double) i > (double) bestI + (double) P.test_radius) && err[j] < second) second = err[j];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) second = fmin(second, __shfl_xor_sync(0xffffffffu, second, off));
    if (lane != 0) return;
    const double newQ = second / bestE;
    if (newQ < pts.quality[p] || numSteps > 10) pts.quality[p] = newQ;
    if (bestE >= pts.energyTH[p] * (double) P.extra_slack) {
        leave(last == IPS_OUTLIER ? IPS_OOB : IPS_OUTLIER, -1.0, -1.0, 0.0);
        return;
    }
    double lo, hi;
    if (dx * dx > dy * dy) {
        lo = (pr2 * (bestU - errPx * dx) - pr0) / (pr_.Kt[0] - pr_.Kt[2] * (bestU - errPx * dx));
        hi = (pr2 * (bestU + errPx * dx) - pr0) / (pr_.Kt[0] - pr_.Kt[2] * (bestU + errPx * dx));
    } else {
        lo = (pr2 * (bestV - errPx * dy) - pr1) / (pr_.Kt[1] - pr_.Kt[2] * (bestV - errPx * dy));
        hi = (pr2 * (bestV + errPx * dy) - pr1) / (pr_.Kt[1] - pr_.Kt[2] * (bestV + errPx * dy));
    }
    if (lo > hi) { const double s = lo; lo = hi; hi = s; }
    pts.idmin[p] = lo; pts.idmax[p] = hi;
    leave(IPS_GOOD, bestU, bestV, 2 * errPx);
}

// optimizeImmaturePoint: one warp per listed point; lane = (target within a round of 4) * 8 + pattern pixel.
// pairs[host slot * TRC_MAXF + target slot]; targets = the live frame slots in window order.
struct ActivateOut { int rc; float idepth; unsigned in_mask; };     // in_mask: bit per target (window order) whose residual ended IN

__global__ void __launch_bounds__(128) trc_activate_kernel(const TrcParams P, const PointsDev pts, const int n, const int *__restrict__ ids, const int min_obs,
                                                          const int num_slots, const int *__restrict__ slots, const PairDev *__restrict__ pairs,
                                                          const FrameImg *__restrict__ frames, ActivateOut *__restrict__ out) {
    const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= n) return;
    const int p = ids[q];
    const int hs = pts.host[p];
    const float2 xy = pts.xy[p];
    const int k = lane & 7, sub = lane >> 3;
    const double ux = ((double) xy.x + c_pat[k][0] - P.cx) / P.fx, uy = ((double) xy.y + c_pat[k][1] - P.cy) / P.fy;
    const FrameImg hf = frames[hs];
    const float4 gt = hf.grad[(size_t) ((int) xy.y + c_pat[k][1]) * P.W + (int) xy.x + c_pat[k][0]];     // getDerivativePatch
    const float c = P.outlier_th_sum;
    const double weight = (double) sqrtf(c / (c + (gt.y * gt.y + gt.z * gt.z)));
    const double energyTH = pts.energyTH[p], huber = (double) P.huber;
    // targets in window order, the host skipped; lane t keeps the residual state of target t
    int nres = 0;
    for (int s = 0; s < num_slots; s++) if (slots[s] != hs) nres++;
    int my_state = RES_IN, my_new_state = RES_OUTLIER;
    double my_energy = 0.0, my_new_energy = 0.0;

    // one evaluation of all residuals at `idepth`: returns (energy, Hdd, bd) in the reference's fp32 running sums
    auto evaluate = [&](const float idepth, const float slack, float &Hdd, float &bd) -> float {
        float total = 0.f;
        Hdd = 0.f; bd = 0.f;
        for (int base = 0; base < nres; base += 4) {
            const int ti = base + sub;                 // target index of this lane
            bool oob = false;
            double eterm = 0, hterm = 0, bterm = 0;
            // state of this lane's target lives in lane ti
            const int st_t = __shfl_sync(0xffffffffu, my_state, ti < nres ? ti : 0);
            if (ti < nres && st_t != RES_OOB) {
                int s = 0, cnt = -1, slot = 0;
                for (; s < num_slots; s++) { if (slots[s] != hs) cnt++; if (cnt == ti) { slot = slots[s]; break; } }
                const PairDev &pr = pairs[hs * TRC_MAXF + slot];
                const double id = (double) idepth;
                const double q0 = pr.R[0] * ux + pr.R[1] * uy + pr.R[2] + pr.t[0] * id, q1 = pr.R[3] * ux + pr.R[4] * uy + pr.R[5] + pr.t[1] * id,
                             q2 = pr.R[6] * ux + pr.R[7] * uy + pr.R[8] + pr.t[2] * id;
                const double px = q0 / q2, py = q1 / q2;
                const double jx = P.fx * px + P.cx, jy = P.fy * py + P.cy;
                const double dres = 1.0 / q2;
                if (!inside(P, jx, jy, 1.0) || dres <= 0) oob = true;
                else {
                    const float4 gv = interp_grad(frames[slot].grad, P.W, (float) jx, (float) jy);
                    const double r = (double) gv.x - (pr.a * (double) gt.x + pr.b);
                    double hw = fabs(r) < huber ? 1.0 : huber / fabs(r);
                    eterm = weight * weight * hw * r * r * (2 - hw);
                    const double dxi = (double) gv.y * P.fx, dyi = (double) gv.z * P.fy;
                    const double d_id = dxi * dres * (pr.t[0] - pr.t[2] * px) + dyi * dres * (pr.t[1] - pr.t[2] * py);
                    hw *= weight * weight;
                    hterm = (hw * d_id) * d_id; bterm = (hw * r) * d_id;
                }
            }
            const unsigned oob_mask = __ballot_sync(0xffffffffu, oob);
            // sequential fold in the reference's order (target by target, pattern pixel by pattern pixel), every lane folding the same values
            for (int j = 0; j < 4 && base + j < nres; j++) {
                const int st_j = __shfl_sync(0xffffffffu, my_state, base + j);
                const double old_e = __shfl_sync(0xffffffffu, my_energy, base + j);
                const unsigned om = (oob_mask >> (8 * j)) & 0xffu;
                const int first_oob = om ? __ffs(om) - 1 : 8;
                float energyLeft = 0.f;
                int new_state; double ret;
                if (st_j == RES_OOB) { new_state = RES_OOB; ret = old_e; }
                else {
                    for (int kk = 0; kk < 8; kk++) {
                        const double e = __shfl_sync(0xffffffffu, eterm, 8 * j + kk), hh = __shfl_sync(0xffffffffu, hterm, 8 * j + kk), bb = __shfl_sync(0xffffffffu, bterm, 8 * j + kk);
                        if (kk < first_oob) {
                            energyLeft = (float) ((double) energyLeft + e);
                            Hdd = (float) ((double) Hdd + hh); bd = (float) ((double) bd + bb);
                        }
                    }
                    if (first_oob < 8) { new_state = RES_OOB; ret = old_e; }        // early return keeps the partial Hdd / bd, drops the energy
                    else {
                        if ((double) energyLeft > energyTH * (double) slack) { energyLeft = (float) (energyTH * (double) slack); new_state = RES_OUTLIER; }
                        else new_state = RES_IN;
                        ret = (double) energyLeft;
                        if (lane == base + j) my_new_energy = ret;
                    }
                }
                if (lane == base + j) my_new_state = new_state;
                total = (float) ((double) total + ret);
            }
        }
        return total;
    };
    auto commit = [&]() { my_state = my_new_state; my_energy = my_new_energy; };

    ActivateOut o; o.rc = 0; o.idepth = 0.f; o.in_mask = 0u;
    float lastHdd, lastbd;
    float cur = (float) ((pts.idmax[p] + pts.idmin[p]) * (double) 0.5f);
    // first pass: the reference commits every residual right after linearising it; the energies it reads back are those of the same pass
    float lastEnergy = evaluate(cur, 1000.f, lastHdd, lastbd);
    commit();
    bool done = false;
    if (!isfinite(lastEnergy) || lastHdd < P.min_idepth_h_act) { o.rc = 0; done = true; }
    float lambda = 0.1f;
    for (int it = 0; !done && it < P.gn_iterations; it++) {
        float H = lastHdd;
        H *= 1 + lambda;
        const float step = (float) ((1.0 / (double) H) * (double) lastbd);
        const float newId = cur - step;
        float newHdd, newbd;
        const float newEnergy = evaluate(newId, 1.f, newHdd, newbd);
        if (!isfinite(lastEnergy) || newHdd < P.min_idepth_h_act) { o.rc = 0; done = true; break; }
        if (newEnergy < lastEnergy) {
            cur = newId; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
            commit();
            lambda *= 0.5f;
        } else lambda *= 5.f;
        if ((double) fabsf(step) < 0.0001 * (double) cur) break;
    }
    if (!done) {
        const unsigned in_mask = __ballot_sync(0xffffffffu, lane < nres && my_state == RES_IN);
        if (!isfinite(cur) || cur <= 0.f) o.rc = -1;
        else if (__popc(in_mask) < min_obs || !isfinite(energyTH)) o.rc = -1;
        else { o.rc = 1; o.idepth = cur; o.in_mask = in_mask; }
    }
    if (lane == 0) out[q] = o;
}

}  // namespace cmltrc
