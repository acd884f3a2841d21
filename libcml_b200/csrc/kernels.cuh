// kernels.cuh -- hand-written sm_100a kernels of the photometric-BA hot path.
//
// Reference behaviour restated on the device (file:line under /root/reference/src/cml):
//   linearize_tile_kernel (linearize.cuh)  optimization/dso/DSOBundleAdjustment.cpp:62-316 (linearize) + :2051-2093 (applyRes, as a
//                       double-buffered candidate) + :1568-1599 (fixLinearization bookkeeping)
//   accumulate_kernel   :1648-1779 (addToHessianTop ACTIVE) + MatrixAccumulators.h:776-937 (AccumulatorApprox)
//   schur_kernel        :1880-1937 (addToHessianSC)
//   stitch_kernel       :1781-1878 (stitchDoubleTop) + :1939-2043 (stitchDoubleSC), one CTA per host frame
//   solve_kernel        :1284-1337 (solveLevenbergMarquardt) + :1196-1261 (orthogonalize) + :1427-1451 (frame steps, xAd)
//                       + DSOFrame.h:110-124 (setState) + :259-273 (pair precompute)
//   point_step_kernel   :1455-1487 (point steps) + :976-1026 (doStepFromBackup, point part and convergence test)
//   post_linearize_kernel  :2419-2464 (setNewFrameEnergyTH) + run() accept bookkeeping :843-879
// No tensor cores: there is no dense contraction on this path (north_star).  All cross-thread reductions
// use fixed-order trees, so results are bitwise reproducible run to run.
#pragma once
#include "dev_types.h"
#include "se3.h"

namespace cmlba {

constexpr int DBG_STRIDE = 52;  // resF[8] JIdx[16] JabF[16] Jpdd[2] JIdx2[3] JabJIdx[4] Jab2[3]

__constant__ int c_sx[8] = {0, -1, 1, -2, 0, 2, -1, 0};   // PredefinedPattern::star8 (types.h:1395-1407)
__constant__ int c_sy[8] = {-2, -1, -1, 0, 0, 0, 1, 2};
// the same table as compile-time constants (folded after unrolling; the __constant__ copy costs one LDC per use)
__host__ __device__ __forceinline__ constexpr int pat_sx(int i) { return i == 0 ? 0 : i == 1 ? -1 : i == 2 ? 1 : i == 3 ? -2 : i == 4 ? 0 : i == 5 ? 2 : i == 6 ? -1 : 0; }
__host__ __device__ __forceinline__ constexpr int pat_sy(int i) { return i == 0 ? -2 : i == 1 ? -1 : i == 2 ? -1 : i == 3 ? 0 : i == 4 ? 0 : i == 5 ? 0 : i == 6 ? 1 : 2; }

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch (Engine::launch_pdl): a kernel of the chain lets its successor be scheduled as soon as all of its own
// CTAs are running (the successor's CTAs then sit in pdl_wait until this grid has completed and its writes are visible), so the
// launch latency of kernel k+1 overlaps the execution of kernel k.  No-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

// development timeline (CMLBA_KTRACE=1, tools/ktrace.py): one object at the top of every kernel of the chain; the host points w.ktrace at
// the slot of the launch (Engine::launch_pdl), slot 0 holds the time of the reset
struct KTrace {
    unsigned long long *p;
    static __device__ __forceinline__ unsigned long long now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ KTrace(const DevWin &w, const int) : p(w.ktrace) { if (p && (threadIdx.x & 31) == 0) atomicMin(p, now()); }
    __device__ __forceinline__ void begin() const { if (p && (threadIdx.x & 31) == 0) atomicMin(p + 1, now()); }
    __device__ __forceinline__ ~KTrace() { if (p && (threadIdx.x & 31) == 0) atomicMax(p + 2, now()); }
};
#define KTRACE_ENTER(k) KTrace kt_(w, (k)); pdl_enter(); kt_.begin()
// slot 0 (the sampling kernel, which carries no stamps of its own: register budget) gets the time of this reset = the start of the timed region
__global__ void ktrace_reset_kernel(unsigned long long *kt) { const int i = threadIdx.x; if (i < 3) kt[i] = KTrace::now(); for (int j = 3 + i; j < 3 * 128; j += 96) kt[j] = (j % 3 == 2) ? 0ull : ~0ull; }

// Guard of a pre-launched kernel (`respect_done` argument): 0 = always run; 1 = skip once the Gauss-Newton loop is done (Ctrl.done); 2 = the
// closing sequence of run() launched AHEAD of the host's look at Ctrl.done: run only when the loop is done and the closing linearization
// has not been executed yet.
__device__ __forceinline__ bool launch_skipped(const int guard, const int done, const int final_done) {
    return guard == 1 ? done != 0 : guard == 2 ? !(done != 0 && final_done == 0) : false;
}

// Frame state -> PRE_worldToCam (DSOFrame::setState, DSOFrame.h:110-124)
__device__ inline void frame_set_state(FrameDev &f, const double *state, const DevWin &w) {
    const double sc[10] = {w.scaleT, w.scaleT, w.scaleT, w.scaleR, w.scaleR, w.scaleR, w.scaleA, w.scaleB, w.scaleA, w.scaleB};
    for (int k = 0; k < 10; k++) { f.state[k] = state[k]; f.state_scaled[k] = sc[k] * state[k]; }
    Pose E; for (int k = 0; k < 9; k++) E.R[k] = f.evalR[k]; for (int k = 0; k < 3; k++) E.t[k] = f.evalt[k];
    Pose P = pose_mul(se3_exp(f.state_scaled), E);
    for (int k = 0; k < 9; k++) f.preR[k] = P.R[k];
    for (int k = 0; k < 3; k++) f.pret[k] = P.t[k];
}

// DSOFramePrecomputed::precompute (DSOFrame.h:259-273) for pair (h,t)
__device__ inline void pair_precompute(const DevWin &w, int h, int t) {
    const FrameDev &fh = w.frames[h], &ft = w.frames[t];
    PairPre &pp = w.pairs[h * w.N + t];
    Pose Ph, Pt, Eh, Et;
    for (int k = 0; k < 9; k++) { Ph.R[k] = fh.preR[k]; Pt.R[k] = ft.preR[k]; Eh.R[k] = fh.evalR[k]; Et.R[k] = ft.evalR[k]; }
    for (int k = 0; k < 3; k++) { Ph.t[k] = fh.pret[k]; Pt.t[k] = ft.pret[k]; Eh.t[k] = fh.evalt[k]; Et.t[k] = ft.evalt[k]; }
    Pose T = pose_mul(Pt, pose_inv(Ph));
    Pose T0 = pose_mul(Et, pose_inv(Eh));
    for (int k = 0; k < 9; k++) { pp.R[k] = T.R[k]; pp.R0[k] = T0.R[k]; }
    for (int k = 0; k < 3; k++) { pp.t[k] = T.t[k]; pp.t0[k] = T0.t[k]; }
    // aff_g2l() = (ab_exposure, state_scaled[6], state_scaled[7]); Exposure::to (map/Exposure.h:119-123)
    double a = exp(ft.state_scaled[6] - fh.state_scaled[6]) * ft.exposure / fh.exposure;
    pp.a = a;
    pp.b = ft.state_scaled[7] - a * fh.state_scaled[7];
    pp.b0 = (float) (fh.state_zero[7] * (double) w.scaleB);
    pp.pad = 0.f;
    for (int k = 0; k < 3; k++) { pp.Af[k] = (float) (T.R[3 * k] * w.fxi); pp.Bf[k] = (float) (T.R[3 * k + 1] * w.fyi); }
}

__global__ void pairs_kernel(const DevWin w, const int guard) {
    if (guard && launch_skipped(guard, w.ctrl->done, w.ctrl->final_done)) return;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < w.N * w.N) pair_precompute(w, i / w.N, i % w.N);
}

// run() epilogue BA:885-889: newest frame gets a new FEJ evaluation point, then pairs are refreshed
__global__ void set_evalpt_newest_kernel(const DevWin w, const int guard) {
    if (guard && launch_skipped(guard, w.ctrl->done, w.ctrl->final_done)) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        FrameDev &f = w.frames[w.N - 1];
        for (int k = 0; k < 9; k++) f.evalR[k] = f.preR[k];
        for (int k = 0; k < 3; k++) f.evalt[k] = f.pret[k];
        double nz[10];
        for (int k = 0; k < 10; k++) nz[k] = 0.0;
        nz[6] = f.state[6]; nz[7] = f.state[7];
        frame_set_state(f, nz, w);
        for (int k = 0; k < 10; k++) f.state_zero[k] = nz[k];
    }
}

// ------------------------------------------------------------------------------------------------
// HOT LOOPS 1+2 fused: linearize (BA:62-316) + applyRes (BA:2051-2093) + addToHessianTop(ACTIVE) (BA:1648-1779) live in
// linearize.cuh (linearize_tile_kernel: TMA-staged target tiles, one residual per lane, per-(host,target)-run partial blocks).
// entry e (0..95) of the packed 13x13 block as a product of record fields; x = Jp_x[10], y = Jp_y[10], Q = JIdx2 * Jp,
// B = the 2x3 top-right multipliers, BR = the 6 bottom-right sums.  e is a compile-time constant after unrolling.
__host__ __device__ __forceinline__ float acc_entry(const int e, const float *x, const float *y, const float *Qx, const float *Qy, const float *Bx, const float *By, const float *BR) {
    if (e < 55) {
        int r = 0, base = 0;
        while (e >= base + (10 - r)) { base += 10 - r; r++; }
        const int c = r + (e - base);
        return x[r] * Qx[c] + y[r] * Qy[c];
    }
    if (e < 85) { const int r = (e - 55) / 3, k = (e - 55) % 3; return x[r] * Bx[k] + y[r] * By[k]; }
    if (e < 91) return BR[e - 85];
    return 0.f;
}

// sum over the warp of 32 per-lane values v[0..31]; lane L returns the total of entry L
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], const int lane) {
#pragma unroll
    for (int hsz = 16; hsz >= 1; hsz >>= 1) {
        const bool up = (lane & hsz) != 0;
#pragma unroll
        for (int k = 0; k < hsz; k++) {
            const float keep = up ? v[k + hsz] : v[k];
            const float send = up ? v[k] : v[k + hsz];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, hsz);
        }
    }
    return v[0];
}

}  // namespace cmlba
#include "linearize.cuh"
namespace cmlba {

// ------------------------------------------------------------------------------------------------
// HOT LOOP 3: Schur complement of the inverse depths (addToHessianSC, BA:1880-1937).  One CTA per chunk of <=64
// points hosted in ONE frame.  With s = sqrt(Hdd^-1) every accumulator of the reference is one block of a single
// symmetric rank update of the augmented per-point vector z = s * [ JpJdF over all targets (8N) | Hcd (4) | bdSum ]:
//   D = sum z_u z_u^T   E = sum z_u z_c^T   EB = sum z_u z_b   Hcc = sum z_c z_c^T   bc = sum z_c z_b
// (the N^3 8x8 accumulators of BA:1040 per host are the (8N)^2 matrix D).  4x4 register tiles, 2 LDS.128 per 16 FMA;
// only tiles on or below the diagonal of D are computed; each is stored together with its mirror image.
constexpr int SCZ_PAD = 8;     // z_c (4) z_b (1) pad (3)
__host__ __device__ __forceinline__ size_t schur_smem_bytes(int N) { return sizeof(float) * ((size_t) SC_CHUNK * N * T_STRIDE + (size_t) SC_CHUNK * (8 * N + SCZ_PAD)); }

// addToHessianTop (BA:1648-1779, MatrixAccumulators.h:776-937) from the Jacobian records of the sampling kernel: one warp per (bin,
// slice).  Every lane keeps the 91 sums of the packed 13x13 block in registers over its residuals (lane, lane + 32, ...: the next
// record is in flight while the current one is evaluated), then one transposing butterfly per 32 entries sums the lanes.
__global__ void __launch_bounds__(32) accumulate_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(1);
    const int job = blockIdx.x, bin = job / ACC_SLICES, sl = job - bin * ACC_SLICES, lane = threadIdx.x;
    // everything the first record load needs is requested together (one latency): control block, bin bounds
    const int done_ld = w.ctrl->done, cur = w.ctrl->cur;
    const int b0 = w.res_bin_begin[bin], b1 = w.res_bin_begin[bin + 1];
    if (respect_done && done_ld) { if (lane == 0) atomicAdd(&w.ctrl->acc_done_count, 1); return; }      // skipped jobs count too: the host's target stays in step
    const int len = (b1 - b0 + ACC_SLICES - 1) / ACC_SLICES, a0 = min(b0 + sl * len, b1), a1 = min(a0 + len, b1);
    float acc[ACC_N];
#pragma unroll
    for (int k = 0; k < ACC_N; k++) acc[k] = 0.f;
    const float4 *rj4 = reinterpret_cast<const float4 *>(w.rj[cur]);
    float4 nx[9];                                    // the next record of this lane: in flight while the current one is evaluated
    int r = a0 + lane;
    if (r < a1) {
#pragma unroll
        for (int k = 0; k < 9; k++) nx[k] = __ldg(rj4 + (size_t) r * 9 + k);
    }
    while (r < a1) {
        float rec[RJ_STRIDE];
#pragma unroll
        for (int k = 0; k < 9; k++) { rec[4 * k] = nx[k].x; rec[4 * k + 1] = nx[k].y; rec[4 * k + 2] = nx[k].z; rec[4 * k + 3] = nx[k].w; }
        r += 32;
        if (r < a1) {
#pragma unroll
            for (int k = 0; k < 9; k++) nx[k] = __ldg(rj4 + (size_t) r * 9 + k);
        }
        if (rec[35] != 0.f) {                        // a good residual (the others carry no record)
            float Qx[10], Qy[10];
            const float a00 = rec[20], a01 = rec[21], a11 = rec[22];
#pragma unroll
            for (int k = 0; k < 10; k++) { Qx[k] = a00 * rec[k] + a01 * rec[10 + k]; Qy[k] = a01 * rec[k] + a11 * rec[10 + k]; }
#pragma unroll
            for (int e = 0; e < 91; e++) acc[e] += acc_entry(e, rec, rec + 10, Qx, Qy, rec + 23, rec + 26, rec + 29);
        }
    }
    // sum over the 32 lanes: transposing butterfly, lane L ends up with entry (g, L)
#pragma unroll
    for (int g = 0; g < 3; g++) {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = acc[g * 32 + k];
        w.acc_bin[(size_t) job * ACC_N + g * 32 + lane] = warp_transpose_sum(v, lane);
    }
    // this job's slices are in place: publish (stitch_pair_kernel, on the main stream, waits until every job of this launch has counted)
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(&w.ctrl->acc_done_count, 1);
}

__global__ void __launch_bounds__(256) schur_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(2);
    extern __shared__ __align__(16) float sm[];
    const int N = w.N, NB = 8 * N, ZS = NB + SCZ_PAD;
    const int done_ld = w.ctrl->done, cur = w.ctrl->cur;                 // requested together with the chunk bounds below (one latency)
    const int begin_ld = w.sc_chunk_begin[blockIdx.x], cnt_ld = w.sc_chunk_count[blockIdx.x];
    if (respect_done && done_ld) return;
    float *sT = sm;                                  // [SC_CHUNK][N][T_STRIDE] raw Schur rows
    float *sZ = sT + SC_CHUNK * N * T_STRIDE;        // [SC_CHUNK][ZS]         augmented, scaled
    const int c = blockIdx.x, tid = threadIdx.x;
    const int begin = begin_ld, cnt = cnt_ld;
    // the per-point scalars of the second phase are requested together with the rows (same round trip)
    float priorF_ld = 0.f, idz_ld = 0.f; double id_ld = 0.0;
    if (tid < cnt) { priorF_ld = w.pt_priorF[begin + tid]; id_ld = w.pt_idepth[begin + tid]; idz_ld = w.pt_idepth_zero[begin + tid]; }
    {   // stage the rows of the chunk's points (contiguous in T)
        const float4 *src = reinterpret_cast<const float4 *>(w.T[cur] + (size_t) begin * N * T_STRIDE);
        float4 *dst = reinterpret_cast<float4 *>(sT);
        const int n4 = cnt * N * (T_STRIDE / 4), tot4 = SC_CHUNK * N * (T_STRIDE / 4);
        for (int i = tid; i < tot4; i += 256) dst[i] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // per point sums (BA:1895-1907), one thread per point
    if (tid < SC_CHUNK) {
        float sh = 0.f, hc0 = 0.f, hc1 = 0.f, hc2 = 0.f, hc3 = 0.f, bs = 0.f;
        if (tid < cnt) {
            const int p = begin + tid;
            float Hdd = 0.f, bd = 0.f, h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f; int ng = 0;
            for (int t = 0; t < N; t++) {
                const float4 a = *reinterpret_cast<const float4 *>(sT + (tid * N + t) * T_STRIDE + 8);    // bd Hdd Hcd0 Hcd1
                const float4 b = *reinterpret_cast<const float4 *>(sT + (tid * N + t) * T_STRIDE + 12);   // Hcd2 Hcd3 good pad
                bd += a.x; Hdd += a.y; h0 += a.z; h1 += a.w; h2 += b.x; h3 += b.y; ng += (b.z != 0.f);
            }
            const float priorF = priorF_ld;
            float idh = 0.f, bdSum = 0.f, hd = 0.f;
            if (ng > 0) {
                float Hs = Hdd + priorF;
                if (Hs < 1e-10f) Hs = 1e-10f;
                idh = Hs;
                hd = (float) (1.0 / (double) Hs);
                const float deltaF = (float) (id_ld - (double) idz_ld);
                bdSum = w.marg_mode ? bd : bd + priorF * deltaF;       // addToHessianSC(shiftPriorToZero) (BA:1902)
                sh = sqrtf(hd);
            } else {
                w.pt_max_rel_bs[p] = 0.f;        // BA:1885-1893
            }
            w.pt_Hdd[p] = Hdd; w.pt_bd[p] = bd;
            w.pt_Hcd[p * 4 + 0] = h0; w.pt_Hcd[p * 4 + 1] = h1; w.pt_Hcd[p * 4 + 2] = h2; w.pt_Hcd[p * 4 + 3] = h3;
            w.pt_HdiF[p] = hd; w.pt_bdSumF[p] = bdSum; w.pt_idepth_hessian[p] = idh; w.pt_ngood_cur[p] = ng;
            hc0 = sh * h0; hc1 = sh * h1; hc2 = sh * h2; hc3 = sh * h3; bs = sh * bdSum;
        }
        float *z = sZ + tid * ZS;
        z[NB + 0] = hc0; z[NB + 1] = hc1; z[NB + 2] = hc2; z[NB + 3] = hc3; z[NB + 4] = bs; z[NB + 5] = 0.f; z[NB + 6] = 0.f; z[NB + 7] = 0.f;
        sT[tid * N * T_STRIDE + 15] = sh;            // pad slot of the point's first row carries the scale to the next phase
    }
    __syncthreads();
    for (int i = tid; i < SC_CHUNK * N * 2; i += 256) {      // z_u = s * JpJdF
        const int row = i >> 1, half = i & 1, pnt = row / N, t = row - pnt * N;
        const float sh = sT[pnt * N * T_STRIDE + 15];
        float4 v = *reinterpret_cast<const float4 *>(sT + row * T_STRIDE + half * 4);
        v.x *= sh; v.y *= sh; v.z *= sh; v.w *= sh;
        *reinterpret_cast<float4 *>(sZ + pnt * ZS + t * 8 + half * 4) = v;
    }
    __syncthreads();
    float *out = w.sc_part + (size_t) c * w.sc_stride;
    float *oE = out + NB * NB, *oEB = oE + NB * 4, *oHcc = oEB + NB, *obc = oHcc + 16;
    // tiles: (tr, tc), tc <= tr over the 2N x 2N tile grid of D, then tile columns 2N (z_c) and 2N+1 (z_b) for every tr
    const int ntr = 2 * N, ntri = ntr * (ntr + 1) / 2, ntiles = ntri + 2 * ntr;
    for (int id = tid; id < ntiles; id += 256) {
        int tr, tc;
        if (id < ntri) {
            tr = (int) ((sqrtf(8.f * (float) id + 1.f) - 1.f) * 0.5f);
            while (tr * (tr + 1) / 2 > id) tr--;
            while ((tr + 1) * (tr + 2) / 2 <= id) tr++;
            tc = id - tr * (tr + 1) / 2;
        } else { tr = (id - ntri) >> 1; tc = ntr + ((id - ntri) & 1); }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
        const float *pa = sZ + tr * 4, *pb = sZ + tc * 4;
#pragma unroll 4
        for (int pnt = 0; pnt < SC_CHUNK; pnt++) {
            const float4 a = *reinterpret_cast<const float4 *>(pa + pnt * ZS);
            const float4 b = *reinterpret_cast<const float4 *>(pb + pnt * ZS);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] += av[i] * bv[j];
        }
        if (tc < ntr) {
#pragma unroll
            for (int i = 0; i < 4; i++) *reinterpret_cast<float4 *>(out + (size_t) (tr * 4 + i) * NB + tc * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (tc < tr) {   // mirror image, so that readers stream whole rows
#pragma unroll
                for (int j = 0; j < 4; j++) *reinterpret_cast<float4 *>(out + (size_t) (tc * 4 + j) * NB + tr * 4) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            }
        } else if (tc == ntr) {
#pragma unroll
            for (int i = 0; i < 4; i++) *reinterpret_cast<float4 *>(oE + (tr * 4 + i) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) oEB[tr * 4 + i] = acc[i][0];
        }
    }
    if (tid >= 224 && tid < 244) {       // Hcc (4x4), bc (4): last warp, otherwise idle in the tile loop when N <= 8
        const int k = tid - 224;
        float acc = 0.f;
        const int i = k < 16 ? (k >> 2) : (k - 16), j = k < 16 ? (k & 3) : 4;
        for (int pnt = 0; pnt < SC_CHUNK; pnt++) acc += sZ[pnt * ZS + NB + i] * sZ[pnt * ZS + NB + j];
        if (k < 16) oHcc[k] = acc; else obc[k - 16] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// index of entry (r,c), r<=c, of the packed 13x13 block
__device__ __forceinline__ int acc_index(int r, int c) {
    if (r > c) { int t = r; r = c; c = t; }
    if (c < 10) return r * 10 - (r * (r - 1)) / 2 + (c - r);        // top-left 10x10, row-major upper triangle
    if (r < 10) return 55 + r * 3 + (c - 10);                       // top-right 10x3
    const int rr = r - 10, cc = c - 10;                             // bottom-right 3x3 upper triangle
    return 85 + (rr == 0 ? cc : (rr == 1 ? 2 + cc : 5));
}

// Stitching (fp64), fully parallel: one CTA per ordered frame pair (host i, other frame j).  It sums the chunk
// partials it needs in fixed order (13x13 block of bin i->j; rows j of the host's Schur blocks D, E, EB), applies the
// adjoints and writes every block product that involves this pair to its own slot of st_out; assemble_kernel
// then gathers the slots into the reduced system.  No atomics, no zero-filled per-host matrices.
// AT is diagonal (identity pose block, -a, -1, row-scaled; BA:1075-1092): only its diagonal is used.
//   slot layout (doubles):  A_tt[64] A_it[64] A_ii[64] A_tC[32] A_iC[32] A_CC[16] bA_t[8] bA_i[8] bA_C[4] pad[4]
//                           S_ji[64] S_ii[64] S_jC[32] S_iC[32] bS_j[8] bS_i[8] | S_jk[N][64]
//   slot (i,i) holds Hcc[16] bc[4] of host i instead.
constexpr int ST_A_TT = 0, ST_A_IT = 64, ST_A_II = 128, ST_A_TC = 192, ST_A_IC = 224, ST_A_CC = 256, ST_BA_T = 272, ST_BA_I = 280, ST_BA_C = 288,
              ST_S_JI = 296, ST_S_II = 360, ST_S_JC = 424, ST_S_IC = 456, ST_BS_J = 488, ST_BS_I = 496, ST_S_JK = 504;
__host__ __device__ __forceinline__ int st_stride(int N) { return ST_S_JK + 64 * N; }
constexpr int ST_THREADS = 576;     // 8 * 64 + 40 partial sums + ... : one round of the chunk sums for N = 8

__global__ void __launch_bounds__(ST_THREADS) stitch_pair_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(3);
    extern __shared__ __align__(16) double smd[];
    const int N = w.N, NB = 8 * N, tid = threadIdx.x;
    const int i = blockIdx.x / N, j = blockIdx.x % N;
    double *out = w.st_out + (size_t) blockIdx.x * st_stride(N);
    const int done_ld = respect_done ? w.ctrl->done : 0;
    const int cb = w.host_chunk_begin[i], ce = w.host_chunk_begin[i + 1];       // requested together with the control block (one latency)
    if (done_ld) return;
    if (i == j) {   // calibration block of host i's Schur complement (BA:1908-1909, 2026-2027)
        if (tid < 20) {
            const float *src = w.sc_part + (size_t) cb * w.sc_stride + NB * NB + NB * 5 + tid;
            double s = 0.0;
            for (int c = cb; c < ce; c++, src += w.sc_stride) s += (double) __ldg(src);
            out[tid] = s;
        }
        return;
    }
    double *Dj = smd;              // [8][NB]  rows of frame j of D_i
    double *Ej = Dj + 8 * NB;      // [8][4]   (Dj, Ej, EBj contiguous: filled by one loop)
    double *EBj = Ej + 32;         // [8]
    double *G = EBj + 8;           // [N][8][8] AH_ik
    double *atd = G + N * 64;      // [N][8]   diag(AT_ik)
    double *A = atd + NB;          // [ACC_N]  packed 13x13 block of bin (i -> j)
    double *Y = A + ACC_N;         // [8][8]   sum_k D_jk AH_ik^T
    double *M = Y + 64;            // [8][8]   AH_ij A8
    double *Yp = M + 64;           // [N][8][8] D_jk AH_ik^T per k (summed into Y in fixed order)
    // the adjoints of host i are requested before the chunk sums below start to wait on their own loads (one round trip instead of two)
    constexpr int G_IT = (MAXF * 64 + ST_THREADS - 1) / ST_THREADS;
    double g_ld[G_IT];
#pragma unroll
    for (int it = 0; it < G_IT; it++) { const int e = tid + it * ST_THREADS; g_ld[it] = e < N * 64 ? w.AH[(size_t) (i * N) * 64 + e] : 0.0; }
    const double atd_ld = tid < NB ? w.AT[((size_t) (i * N + (tid >> 3))) * 64 + (tid & 7) * 9] : 0.0;
    for (int e = tid; e < 8 * NB + 40; e += ST_THREADS) {
        int off;
        if (e < 8 * NB) off = j * 8 * NB + e;
        else if (e < 8 * NB + 32) off = NB * NB + j * 32 + (e - 8 * NB);
        else off = NB * NB + NB * 4 + j * 8 + (e - 8 * NB - 32);
        const float *src = w.sc_part + (size_t) cb * w.sc_stride + off;
        double s = 0.0;
#pragma unroll 16
        for (int c = cb; c < ce; c++, src += w.sc_stride) s += (double) __ldg(src);
        Dj[e] = s;                 // Dj, Ej, EBj are contiguous
    }
    if (w.acc_target > 0) {   // accumulate_kernel runs on the side stream: wait (bounded) until all of its jobs have published their slices
        if (tid == 0) {
            int v = 0; long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(&w.ctrl->acc_done_count) : "memory");
                if (v >= w.acc_target || ++spins > (1l << 24)) break;
                __nanosleep(40);
            }
            if (v < w.acc_target) { w.ctrl->failed = 1; w.ctrl->pad1 = 1; w.ctrl->done = 1; }      // never seen with a correct launch sequence; fail the run (own status) instead of reading stale blocks
        }
        __syncthreads();
    }
    if (tid < ACC_N) {   // 13x13 block of bin (i -> j): the slices of accumulate_kernel, fixed order
        const float *src = w.acc_bin + (size_t) (j * N + i) * ACC_SLICES * ACC_N + tid;
        double a = 0.0;
        float part[ACC_SLICES];
#pragma unroll
        for (int sl = 0; sl < ACC_SLICES; sl++) part[sl] = src[sl * ACC_N];     // plain loads: all in flight together (ld.global.cg is a strong access the compiler keeps in order: 16 serial L2 round trips)
#pragma unroll
        for (int sl = 0; sl < ACC_SLICES; sl++) a += (double) part[sl];
        A[tid] = a;
    }
#pragma unroll
    for (int it = 0; it < G_IT; it++) { const int e = tid + it * ST_THREADS; if (e < N * 64) G[e] = g_ld[it]; }
    if (tid < NB) atd[tid] = atd_ld;
    __syncthreads();
    const double *AHj = G + j * 64, *atj = atd + j * 8;
    for (int e = tid; e < N * 64 + 64; e += ST_THREADS) {       // short dependent chains: 8 products per thread
        const int r = (e >> 3) & 7, c = e & 7;
        double s = 0.0;
        if (e < N * 64) { const int k = e >> 6; for (int m = 0; m < 8; m++) s += Dj[r * NB + k * 8 + m] * G[k * 64 + c * 8 + m]; Yp[e] = s; }
        else { for (int m = 0; m < 8; m++) s += AHj[r * 8 + m] * A[acc_index(4 + m, 4 + c)]; M[e - N * 64] = s; }
    }
    __syncthreads();
    if (tid < 64) { double s = 0.0; for (int k = 0; k < N; k++) s += Yp[k * 64 + tid]; Y[tid] = s; }
    __syncthreads();
    const int tot = st_stride(N);
    for (int e = tid; e < tot; e += ST_THREADS) {
        double v = 0.0;
        if (e < ST_A_TC) {                       // 8x8 active blocks
            const int q = e & 63, r = q >> 3, c = q & 7;
            if (e < ST_A_IT) v = atj[r] * A[acc_index(4 + r, 4 + c)] * atj[c];
            else if (e < ST_A_II) v = M[q] * atj[c];
            else for (int l = 0; l < 8; l++) v += M[r * 8 + l] * AHj[c * 8 + l];
        } else if (e < ST_A_CC) {                // 8x4 calibration columns
            const int q = (e - ST_A_TC) & 31, r = q >> 2, c = q & 3;
            if (e < ST_A_IC) v = atj[r] * A[acc_index(4 + r, c)];
            else for (int m = 0; m < 8; m++) v += AHj[r * 8 + m] * A[acc_index(4 + m, c)];
        } else if (e < ST_BA_T) { const int q = e - ST_A_CC; v = A[acc_index(q >> 2, q & 3)]; }
        else if (e < ST_BA_I) { const int r = e - ST_BA_T; v = atj[r] * A[acc_index(4 + r, 12)]; }
        else if (e < ST_BA_C) { const int r = e - ST_BA_I; for (int m = 0; m < 8; m++) v += AHj[r * 8 + m] * A[acc_index(4 + m, 12)]; }
        else if (e < ST_S_JI) { const int q = e - ST_BA_C; v = q < 4 ? A[acc_index(q, 12)] : 0.0; }
        else if (e < ST_S_II) { const int q = e - ST_S_JI; v = atj[q >> 3] * Y[q]; }
        else if (e < ST_S_JC) { const int q = e - ST_S_II, r = q >> 3, c = q & 7; for (int m = 0; m < 8; m++) v += AHj[r * 8 + m] * Y[m * 8 + c]; }
        else if (e < ST_S_IC) { const int q = e - ST_S_JC; v = atj[q >> 2] * Ej[q]; }
        else if (e < ST_BS_J) { const int q = e - ST_S_IC, r = q >> 2, c = q & 3; for (int m = 0; m < 8; m++) v += AHj[r * 8 + m] * Ej[m * 4 + c]; }
        else if (e < ST_BS_I) { const int r = e - ST_BS_J; v = atj[r] * EBj[r]; }
        else if (e < ST_S_JK) { const int r = e - ST_BS_I; for (int m = 0; m < 8; m++) v += AHj[r * 8 + m] * EBj[m]; }
        else { const int q = e - ST_S_JK, k = q >> 6, r = (q >> 3) & 7, c = q & 7; v = atj[r] * Dj[r * NB + k * 8 + c] * atd[k * 8 + c]; }
        out[e] = v;
    }
}

// Gathers the pair slots into sys = [HA | bA | H_sc | b_sc] (the multi-GPU allreduce payload), one thread per
// element, fixed summation order.  Both matrices come out completed exactly like the tails of stitchDoubleTop
// (BA:1857-1876: H[h,t] += H[t,h]^T, calibration rows mirrored) and stitchDoubleSC (BA:2033-2037).
// ---- reduced-system exchange over NVLink peer memory ------------------------------------------------------------------
// Every rank owns one cudaIpc-mapped buffer [flags u64[16] | pad | slots[2 epochs][world sources]].  assemble_kernel PUSHES
// every element of this rank's partial sys into slot[epoch & 1][rank] of EVERY rank (P2P stores are fire-and-forget, no
// remote latency on the critical path); its last CTA then stores the epoch into flags[rank] of every peer (release,
// system scope).  The consumer (the prologue of solve_kernel, or p2p_allreduce_kernel) spins on its LOCAL flags and sums its
// LOCAL slots in rank order: identical bits on every rank, no NCCL launch, no remote load.
// Two epochs suffice: a rank can only be one exchange ahead of the slowest reader (its next signal needs every peer's flag).
__device__ __forceinline__ unsigned long long *p2p_flags(const DevWin &w, int rk) { return reinterpret_cast<unsigned long long *>(w.p2p_base[rk]); }
__device__ __forceinline__ double *p2p_slot(const DevWin &w, int owner, int src) {
    return reinterpret_cast<double *>(w.p2p_base[owner] + 256) + ((size_t) (w.p2p_epoch & 1ull) * w.world + src) * P2P_SLOT_DOUBLES;
}
// post-linearize records: region after the sys slots, [2 epochs][world sources][P2P_POST_DOUBLES]; flags u64[16] at byte 128
__device__ __forceinline__ double *p2p_post_slot(const DevWin &w, int owner, int src) {
    return reinterpret_cast<double *>(w.p2p_base[owner] + 256) + (size_t) 2 * w.world * P2P_SLOT_DOUBLES + ((size_t) (w.p2p_post_epoch & 1ull) * w.world + src) * P2P_POST_DOUBLES;
}
__device__ __forceinline__ void p2p_signal(const DevWin &w) {      // one thread, after a __threadfence_system() by the writers
    for (int q = 0; q < w.world; q++) {
        unsigned long long *f = p2p_flags(w, q) + w.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(f), "l"(w.p2p_epoch) : "memory");
    }
}
// whole CTA (>= world threads).  Returns false on timeout (a peer never arrived): the caller flags the run as failed.
__device__ __forceinline__ bool p2p_reduce(const DevWin &w, double *dst, const int count) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    if ((int) threadIdx.x < w.world) {
        const unsigned long long *f = p2p_flags(w, w.rank) + threadIdx.x;
        unsigned long long v = 0; long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(f) : "memory");
            if (v >= w.p2p_epoch) break;
            if (++spins > (1l << 26)) { s_ok = 0; break; }     // ~10 s (the ranks are aligned by NCCL at the start of run()): never hang the device on a protocol error
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (!s_ok) return false;
    const double *base = p2p_slot(w, w.rank, 0);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        double sum = 0.0;
        for (int q = 0; q < w.world; q++) sum += __ldcg(base + (size_t) q * P2P_SLOT_DOUBLES + e);
        dst[e] = sum;
    }
    __syncthreads();
    return true;
}
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const DevWin w, const int respect_done) {
    pdl_enter();                                     // launched while assemble_kernel still runs; its own pushes must be complete before the sums are stored
    if (respect_done && w.ctrl->done) return;
    if (!p2p_reduce(w, w.sys, 2 * w.n * w.n + 2 * w.n) && threadIdx.x == 0) { w.ctrl->failed = 1; w.ctrl->pad0 = 1; w.ctrl->done = 1; }
}

// cross-rank barrier on the same epoch flags (an exchange without payload): aligns the ranks before a timed pass so that a rank
// whose L2 flush finished early does not charge its wait for the others to the pass.  Bounded spin like p2p_reduce.
__global__ void __launch_bounds__(32) p2p_barrier_kernel(const DevWin w) {
    if (threadIdx.x == 0) { __threadfence_system(); p2p_signal(w); }
    if ((int) threadIdx.x < w.world) {
        const unsigned long long *f = p2p_flags(w, w.rank) + threadIdx.x;
        unsigned long long v = 0; long spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(f) : "memory");
            if (v >= w.p2p_epoch || ++spins > (1l << 26)) break;
            __nanosleep(100);
        }
    }
}

// element e of this rank's sys: local buffer, or (peer exchange) the rank's slot in every rank's buffer
__device__ __forceinline__ void sys_emit(const DevWin &w, const bool p2p, const int e, const double v) {
    if (!p2p) { w.sys[e] = v; return; }
    for (int q = 0; q < w.world; q++) p2p_slot(w, q, w.rank)[e] = v;
}
// Every element of sys is a fixed-order sum of pair-slot entries of one of three shapes: a strided sequence over the frames k that skips
// up to two of them (sequence A), a second such sequence (B), and up to two single entries.  The element only selects the shapes (integer
// work); ONE copy of the gather loops then does the loads, so the kernel stays a few hundred instructions (the per-case unrolled
// version was 127 KB of SASS and spent its time fetching instructions).
__device__ __forceinline__ double gather_seq(const double *st, const int N, const int off, const int stride, const int skip1, const int skip2) {
    double v = 0.0;
    for (int k0 = 0; k0 < N; k0 += 8) {
        // eight unconditional loads into eight registers (a skipped k reads entry 0 of the sequence, which exists, and is discarded): they are
        // all in flight before the first add.  Written as "load if used" the compiler chained load -> add -> load through one register.
        double x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { const int k = k0 + j; const bool use = k < N && k != skip1 && k != skip2; x[j] = st[off + (use ? k : 0) * stride]; if (!use) x[j] = 0.0; }
#pragma unroll
        for (int j = 0; j < 8; j++) v += x[j];
    }
    return v;
}
__device__ __forceinline__ void assemble_body(const DevWin &w, const bool p2p) {
    const int N = w.N, n = w.n, nn = n * n, S = st_stride(N);
    const double *st = w.st_out;
    const int nelem = 2 * nn + 2 * n;
    if ((int) blockIdx.x * 256 >= nelem) {
        // calibration block of the active part: HA[C,C] (16) and bA[C] (4) sum over all N(N-1) pairs -> one warp per entry
        const int k = ((int) blockIdx.x * 256 - ((nelem + 255) / 256) * 256 + (int) threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (k >= 20) return;
        const int off = k < 16 ? ST_A_CC + k : ST_BA_C + (k - 16);
        double v = 0.0;
        for (int q = lane; q < N * N; q += 32) { const int i = q / N, j = q - i * N; double x = 0.0; if (i != j) x = st[(size_t) q * S + off]; v += x; }
        v = warp_sum_d(v);
        if (lane == 0) sys_emit(w, p2p, k < 16 ? (k >> 2) * n + (k & 3) : nn + (k - 16), v);
        return;
    }
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= nelem) return;
    const bool schur = e >= nn + n;
    const int q = schur ? e - nn - n : e;
    // shapes: sequence A (offA + k * strA, k != skA1, skA2), sequence B, singles one1 / one2 (added in this order)
    int offA = -1, strA = 0, skA1 = -1, skA2 = -1, offB = -1, strB = 0, skB = -1, one1 = -1, one2 = -1;
    auto row_col = [&](const int a, const int o_row, const int o_col) {       // sum_k!=a slot(a,k)[o_row] + sum_k!=a slot(k,a)[o_col]
        offA = a * N * S + o_row; strA = S; skA1 = a;
        offB = a * S + o_col; strB = N * S; skB = a;
    };
    if (q < nn) {
        int r = q / n, c = q - r * n;
        if (r < 4 && c >= 4) { const int t = r; r = c; c = t; }          // calibration rows mirror the columns
        if (r < 4) {                                                      // (C,C)
            if (!schur) return;                                           // written by the warp-per-entry blocks above
            offA = r * 4 + c; strA = (N + 1) * S;                         // sum_i slot(i,i)
        } else if (c < 4) {                                               // (frame a, C)
            const int a = (r - 4) >> 3, rr = (r - 4) & 7;
            row_col(a, (schur ? ST_S_IC : ST_A_IC) + rr * 4 + c, (schur ? ST_S_JC : ST_A_TC) + rr * 4 + c);
        } else {
            const int a = (r - 4) >> 3, rr = (r - 4) & 7, b = (c - 4) >> 3, cc = (c - 4) & 7;
            if (a == b) {
                if (!schur) row_col(a, ST_A_II + rr * 8 + cc, ST_A_TT + rr * 8 + cc);
                else row_col(a, ST_S_II + rr * 8 + cc, ST_S_JK + a * 64 + rr * 8 + cc);
            } else {
                // the (lo,hi) orientation is summed the same way from both sides: the result is bitwise symmetric
                const int lo = a < b ? a : b, hi = a < b ? b : a, rl = a < b ? rr : cc, rh = a < b ? cc : rr;   // element (lo rl, hi rh)
                if (!schur) { one1 = (lo * N + hi) * S + ST_A_IT + rl * 8 + rh; one2 = (hi * N + lo) * S + ST_A_IT + rh * 8 + rl; }
                else {
                    offA = lo * S + ST_S_JK + hi * 64 + rl * 8 + rh; strA = N * S; skA1 = lo; skA2 = hi;      // sum_k!=lo,hi slot(k,lo)
                    one1 = (hi * N + lo) * S + ST_S_JI + rl * 8 + rh; one2 = (lo * N + hi) * S + ST_S_JI + rh * 8 + rl;
                }
            }
        }
    } else {
        const int r = q - nn;
        if (r < 4) {
            if (!schur) return;
            offA = 16 + r; strA = (N + 1) * S;
        } else {
            const int a = (r - 4) >> 3, rr = (r - 4) & 7;
            row_col(a, (schur ? ST_BS_I : ST_BA_I) + rr, (schur ? ST_BS_J : ST_BA_T) + rr);
        }
    }
    double v = 0.0;
    if (offA >= 0) v = gather_seq(st, N, offA, strA, skA1, skA2);
    if (offB >= 0) v += gather_seq(st, N, offB, strB, skB, -1);
    if (one1 >= 0) { const double x1 = st[one1], x2 = st[one2]; v += x1; v += x2; }
    sys_emit(w, p2p, e, v);
}

__global__ void __launch_bounds__(256) assemble_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(4);
    if (respect_done && w.ctrl->done) return;
    const bool p2p = w.p2p_on && w.world > 1;
    assemble_body(w, p2p);
    if (p2p) {   // the last CTA to finish publishes the epoch to every peer
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int ticket = atomicAdd(&w.ctrl->asm_done_count, 1);
            if (ticket == (int) gridDim.x - 1) { w.ctrl->asm_done_count = 0; __threadfence_system(); p2p_signal(w); }
        }
    }
}

}  // namespace cmlba
#include "tail.cuh"
namespace cmlba {

// ------------------------------------------------------------------------------------------------
// Reduced camera system: assemble, damp, Jacobi-scale, LDL^T (lower triangle, like Eigen's default), solve,
// orthogonalise against the gauge nullspaces, frame steps + new frame states + pair constants, xAd.
__host__ __device__ __forceinline__ bool solve_stages_hs(const int n) { return ((size_t) 2 * n * n + 5 * n + 256 + 8 * MAXF) * sizeof(double) <= (size_t) 200 * 1024; }
__host__ __device__ __forceinline__ size_t solve_smem_doubles(const int n) { return (size_t) (solve_stages_hs(n) ? 2 : 1) * n * n + 5 * n + 256 + 8 * MAXF; }

// 1 / x for a normal, finite x: hardware seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps (quadratic: >= 52 bits up to rounding, <= 2 ulp).
// The IEEE division is a ~30-instruction dependent chain and the LDL^T pivots are strictly sequential.
__device__ __forceinline__ double rcp_f64(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

__global__ void __launch_bounds__(256) solve_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(5);
    // development: phase stamps of the LAST solve of a traced run (tools/ktrace.py), 32 slots behind the launch slots
#define SOLVE_STAMP(k) do { if (kt_.p && tid_ == 0) w.ktrace_base[3 * 128 + (k)] = KTrace::now(); } while (0)
    const int tid_ = threadIdx.x;
    SOLVE_STAMP(0);
    Ctrl *ctrl = w.ctrl;
    if (respect_done && ctrl->done) return;
    extern __shared__ __align__(16) double smd[];
    const int N = w.N, n = w.n, m = 8 * N, tid = threadIdx.x, nn = n * n;
    double *H = smd;            // [n][n]
    double *b = H + nn;         // [n]
    double *s = b + n;          // [n]
    double *x = s + n;          // [n]
    double *red = x + n;        // [256]
    double *invd = red + 256;   // [m] reciprocals of the LDL^T pivots
    double *sysHA = w.sys, *sysbA = w.sys + nn, *sysHS = w.sys + nn + n, *sysbS = w.sys + 2 * nn + n;
    if (w.p2p_on && w.world > 1) {   // sum of the ranks' partial systems over NVLink (replaces ncclAllReduce + its launch)
        if (!p2p_reduce(w, w.sys, 2 * nn + 2 * n)) { if (tid == 0) { ctrl->failed = 1; ctrl->pad0 = 1; ctrl->done = 1; } return; }
    }
    // H <- HA, scratch <- H_sc (already completed by assemble_kernel): both matrices are requested together, eight elements per thread in flight
    const bool stage_hs = solve_stages_hs(n);   // large windows (N > 11): the Schur part is read from global memory where it is used
    const double *Hs = stage_hs ? H + nn + 4 * n + 256 + m : sysHS;      // [n][n] Schur part, staged next to the system (solve_smem_doubles() reserves it)
    for (int e0 = 0; e0 < nn; e0 += 256 * 4) {
        double va[4], vs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const int e = e0 + k * 256 + tid; va[k] = e < nn ? sysHA[e] : 0.0; vs[k] = (stage_hs && e < nn) ? sysHS[e] : 0.0; }
#pragma unroll
        for (int k = 0; k < 4; k++) { const int e = e0 + k * 256 + tid; if (e < nn) { H[e] = va[k]; if (stage_hs) (H + nn + 4 * n + 256 + m)[e] = vs[k]; } }
    }
    __syncthreads();
    SOLVE_STAMP(1);   // system loaded
    const double lambda = w.fix_lambda ? w.fixed_lambda : ctrl->lambda;
    const int wid = tid >> 5, lane = tid & 31;
    // b = bL + bM + bA - b_sc ; H = HL + HM + HA (BA:1299-1300); HL = diag(prior), bL = prior*delta_prior (BA:1857-1865)
    for (int e = tid; e < n; e += 256) {
        double v = sysbA[e] - sysbS[e];
        if (e >= 4) {
            const FrameDev &f = w.frames[(e - 4) >> 3];
            const int k = (e - 4) & 7;
            v += f.prior[k] * f.state[k];                       // delta_prior = state - prior_zero, prior_zero == 0
            s[e] = f.state[k] - f.state_zero[k];                // delta (BA:1401), staged for the H_M product
        } else s[e] = 0.0;
        b[e] = v;
    }
    __syncthreads();
    if (w.has_HM) {   // bM_top = b_M + H_M * delta (BA:1401); zero when disableMarginalization (BA:1395-1398)
        for (int e = wid; e < n; e += 8) {
            double hm = 0.0;
            for (int c = lane; c < n; c += 32) hm += w.HM[(size_t) e * n + c] * s[c];
            hm = warp_sum_d(hm);
            if (lane == 0) b[e] += w.bM[e] + hm;
        }
    }
    double nrm[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // squared Frobenius norms of HA HL HM Hsc | bA bL bM bsc (the Statistic series of BA.h:221-228, BA:1415-1418)
    for (int e = tid; e < nn; e += 256) {
        const int r = e / n, c = e % n;
        double v = H[e];
        nrm[0] += v * v;
        if (w.has_HM) { const double hm = w.HM[e]; v += hm; nrm[2] += hm * hm; }
        if (r == c && r >= 4) { const double pr = w.frames[(r - 4) >> 3].prior[(r - 4) & 7]; v += pr; nrm[1] += pr * pr; }
        if (r == c) v *= (1.0 + lambda);                        // BA:1306-1308
        const double hs = Hs[e];
        nrm[3] += hs * hs;
        v -= hs * (1.0 / (1.0 + lambda));                       // BA:1309
        H[e] = v;
    }
    for (int e = tid; e < n; e += 256) {
        const double ba = sysbA[e], bs = sysbS[e];
        nrm[4] += ba * ba; nrm[7] += bs * bs;
        if (e >= 4) { const FrameDev &f = w.frames[(e - 4) >> 3]; const double bl = f.prior[(e - 4) & 7] * f.state[(e - 4) & 7]; nrm[5] += bl * bl; }
        if (w.has_HM) { const double bm = w.bM[e]; nrm[6] += bm * bm; }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) nrm[k] = warp_sum_d(nrm[k]);
    if (lane == 0) for (int k = 0; k < 8; k++) red[wid * 8 + k] = nrm[k];
    __syncthreads();
    if (tid < 8) { double t = 0.0; for (int q = 0; q < 8; q++) t += red[q * 8 + tid]; ctrl->stats[5 + tid] = sqrt(t); }
    __syncthreads();
    for (int e = tid; e < n; e += 256) s[e] = 1.0 / sqrt(H[e * n + e] + 10.0);   // BA:1312
    __syncthreads();
    SOLVE_STAMP(2);   // damped, scale factors
    // scaled system on the trailing m x m block (calibration fixed, BA:1319); only the lower triangle is read
#define AA(r, c) H[(size_t) (4 + (r)) * n + 4 + (c)]
    for (int e = tid; e < m * m; e += 256) { const int r = e / m, c = e % m; if (r >= c) AA(r, c) = AA(r, c) * s[4 + r] * s[4 + c]; }
    for (int e = tid; e < m; e += 256) x[e] = s[4 + e] * b[4 + e];
    __syncthreads();
    SOLVE_STAMP(3);   // scaled
    // LDL^T without pivoting (SPD after damping + priors), right-looking in panels of 8 columns (= one frame).  The right-hand
    // side rides along as an extra row, so the forward substitution L y = rhs costs no pass of its own:
    //   (1) warp 0 factors the 8x8 diagonal block in registers (shuffles) and forward-substitutes the panel's 8 rhs entries,
    //   (2) one thread per row below solves its panel row (operands preloaded: the chain is 36 dependent FMAs) and updates its rhs,
    //   (3) rank-8 update of the trailing lower triangle in 4x4 register tiles.  Three barriers per panel.
    double *blk = red;          // [64] L11[c][j] * d_j of the current panel (j < c), reused every panel
    for (int k0 = 0; k0 < m; k0 += 8) {
        if (wid == 0) {
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = (lane < 8 && c <= lane) ? AA(k0 + lane, k0 + c) : 0.0;
            double xv = lane < 8 ? x[k0 + lane] : 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dk = __shfl_sync(0xffffffffu, a[k], k);
                const double idk = rcp_f64(dk);                    // one reciprocal per pivot (on the critical path of all 8N pivots); rows are scaled by multiplication
                if (lane == k) invd[k0 + k] = idk;
                const double lrk = a[k] * idk;
                if (lane > k) a[k] = lrk;
                const double yk = __shfl_sync(0xffffffffu, xv, k);  // y_k is final once the pivots 0..k-1 have been applied
                if (lane > k && lane < 8) xv -= lrk * yk;
                if (lane > k && lane < 8) blk[lane * 8 + k] = lrk * dk;
#pragma unroll
                for (int c = k + 1; c < 8; c++) {
                    const double lck = __shfl_sync(0xffffffffu, a[k], c);
                    if (lane >= c) a[c] -= lrk * dk * lck;
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) if (lane < 8 && c <= lane) AA(k0 + lane, k0 + c) = a[c];
            if (lane < 8) x[k0 + lane] = xv;
        }
        __syncthreads();
        const int rem = m - k0 - 8;
        if (tid < rem) {
            const int r = k0 + 8 + tid;
            double ld[28], idv[8], yv[8], l[8], ar[8];
#pragma unroll
            for (int c = 1, q = 0; c < 8; c++)
#pragma unroll
                for (int jj = 0; jj < c; jj++, q++) ld[q] = blk[c * 8 + jj];
#pragma unroll
            for (int c = 0; c < 8; c++) { idv[c] = invd[k0 + c]; yv[c] = x[k0 + c]; ar[c] = AA(r, k0 + c); }
            double xr = x[r];
#pragma unroll
            for (int c = 0, q = 0; c < 8; c++) {
                double v = ar[c];
#pragma unroll
                for (int jj = 0; jj < c; jj++, q++) v -= l[jj] * ld[q];
                l[c] = v * idv[c];
                xr -= l[c] * yv[c];
            }
#pragma unroll
            for (int c = 0; c < 8; c++) AA(r, k0 + c) = l[c];
            x[r] = xr;
        }
        __syncthreads();
        const int nt = rem >> 2, ntiles = nt * (nt + 1) / 2;
        for (int id = tid; id < ntiles; id += 256) {
            int tr = (int) ((sqrtf(8.f * (float) id + 1.f) - 1.f) * 0.5f);
            while (tr * (tr + 1) / 2 > id) tr--;
            while ((tr + 1) * (tr + 2) / 2 <= id) tr++;
            const int tc = id - tr * (tr + 1) / 2;
            const int r0 = k0 + 8 + 4 * tr, c0 = k0 + 8 + 4 * tc;
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) acc[i][jj] = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double dk = AA(k0 + k, k0 + k);
                double lr[4], lc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { lr[i] = AA(r0 + i, k0 + k); lc[i] = AA(c0 + i, k0 + k) * dk; }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int jj = 0; jj < 4; jj++) acc[i][jj] += lr[i] * lc[jj];
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) if (r0 + i >= c0 + jj) AA(r0 + i, c0 + jj) -= acc[i][jj];
        }
        __syncthreads();
    }
    SOLVE_STAMP(4);   // factorised + forward
    // z = y / d, then backward L^T x = z: column dots over the rows below (one warp per column), 8x8 block by shuffles
    for (int e = tid; e < m; e += 256) x[e] = x[e] * invd[e];
    __syncthreads();
    for (int k0 = m - 8; k0 >= 0; k0 -= 8) {
        {   // warp c: sum_{r >= k0+8} L[r][k0+c] x[r]
            double part = 0.0;
            for (int r = k0 + 8 + lane; r < m; r += 32) part += AA(r, k0 + wid) * x[r];
            part = warp_sum_d(part);
            if (lane == 0) red[128 + wid] = part;
        }
        __syncthreads();
        if (wid == 0) {
            double lc[8];   // column `lane` of the unit lower block: L[j][lane], j > lane
#pragma unroll
            for (int j = 1; j < 8; j++) lc[j] = (lane < 8 && j > lane) ? AA(k0 + j, k0 + lane) : 0.0;
            double v = lane < 8 ? x[k0 + lane] - red[128 + lane] : 0.0;
#pragma unroll
            for (int j = 7; j >= 1; j--) {
                const double xj = __shfl_sync(0xffffffffu, v, j);
                if (lane < j) v -= lc[j] * xj;
            }
            if (lane < 8) x[k0 + lane] = v;
        }
        __syncthreads();
    }
#undef AA
    SOLVE_STAMP(5);   // backward
    // x = SVecI * y, calibration entries 0 (BA:1314-1319)
    for (int e = tid; e < n; e += 256) b[e] = (e >= 4) ? s[e] * x[e - 4] : 0.0;   // b now holds x
    __syncthreads();
    if (ctrl->iteration >= 2) {                                  // orthogonalize(x) (BA:1332-1334), projector precomputed on the host
        // P is symmetric: thread (g, r) sums P[c][r] * x[c] over the g-th slice of c (coalesced over r, eight loads in flight), then the slices are added
        const int groups = max(1, min(256 / n, 8)), chunk = (n + groups - 1) / groups, g = tid / n, r = tid - g * n;
        if (g < groups) {
            double v = 0.0;
            const int c_lo = g * chunk, c_hi = min(n, c_lo + chunk);
#pragma unroll 8
            for (int c = c_lo; c < c_hi; c++) v += w.Pns[(size_t) c * n + r] * b[c];
            red[g * n + r] = v;
        }
        __syncthreads();
        if (tid < n) { double v = 0.0; for (int q = 0; q < groups; q++) v += red[q * n + tid]; s[tid] = b[tid] - v; }
        __syncthreads();
        for (int e = tid; e < n; e += 256) b[e] = s[e];
        __syncthreads();
    }
    SOLVE_STAMP(6);   // orthogonalised
    for (int e = tid; e < n; e += 256) w.x[e] = b[e];
    // statistics (BA:1415-1425)
    if (tid < 32) {
        double v = 0.0;
        for (int e = tid; e < n; e += 32) v += b[e] * b[e];
        v = warp_sum_d(v);
        if (tid == 0) ctrl->stats[4] = sqrt(v);
    }
    SOLVE_STAMP(7);
    // frame steps + states (BA:1433-1441, 957-973; DSOFrame::doStepFromBackup) and convergence sums
    if (tid < N) {
        FrameDev &f = w.frames[tid];
        double step[10];
        bool fin = true;
        for (int k = 0; k < 8; k++) { step[k] = -b[4 + 8 * tid + k]; fin = fin && isfinite(step[k]); }
        step[8] = step[9] = 0.0;
        if (!fin) for (int k = 0; k < 10; k++) step[k] = 0.0;    // DSOFrame::setStep (DSOFrame.h:205-214)
        if (w.update_points_only) for (int k = 0; k < 6; k++) step[k] = 0.0;
        double ns[10];
        for (int k = 0; k < 10; k++) { f.state_backup[k] = f.state[k]; f.step[k] = step[k]; ns[k] = f.state[k] + step[k]; }
        frame_set_state(f, ns, w);
    }
    __syncthreads();
    if (tid == 0) {
        float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
        for (int i = 0; i < N; i++) {
            const double *st = w.frames[i].step;
            sumA += (float) (st[6] * st[6]); sumB += (float) (st[7] * st[7]);
            sumT += (float) (st[0] * st[0] + st[1] * st[1] + st[2] * st[2]);
            sumR += (float) (st[3] * st[3] + st[4] * st[4] + st[5] * st[5]);
        }
        ctrl->sumA = sumA / N; ctrl->sumB = sumB / N; ctrl->sumT = sumT / N; ctrl->sumR = sumR / N;
    }
    SOLVE_STAMP(8);   // frame states
    // xAd[h*N+t] = x_h^T AH + x_t^T AT (BA:1447)
    for (int e = tid; e < N * N * 8; e += 256) {
        const int ht = e >> 3, c = e & 7, h = ht / N, t = ht % N;
        const double *ah = w.AH + (size_t) ht * 64, *at = w.AT + (size_t) ht * 64;
        double v = 0.0, pa[8], pt[8];
#pragma unroll
        for (int r = 0; r < 8; r++) { pa[r] = ah[r * 8 + c]; pt[r] = at[r * 8 + c]; }      // 16 independent loads in flight
#pragma unroll
        for (int r = 0; r < 8; r++) v += b[4 + 8 * h + r] * pa[r] + b[4 + 8 * t + r] * pt[r];
        w.xAd[e] = v;
    }
    SOLVE_STAMP(9);   // xAd
    for (int e = tid; e < N * N; e += 256) pair_precompute(w, e / N, e % N);
    SOLVE_STAMP(10);
#undef SOLVE_STAMP
}

// ------------------------------------------------------------------------------------------------
// Per-point back-substitution and step (BA:1455-1487, 976-994) + convergence test (BA:1013-1026)
__global__ void __launch_bounds__(256) point_step_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(6);
    Ctrl *ctrl = w.ctrl;
    if (respect_done && ctrl->done) return;
    const int N = w.N, cur = ctrl->cur;
    __shared__ double s_xad[MAXF * MAXF * 8];
    for (int e = threadIdx.x; e < N * N * 8; e += 256) s_xad[e] = w.xAd[e];
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    double nid = 0.0, pe = 0.0; int cntid = 0; int bad = 0;
    if (p < w.P) {
        double step = 0.0;
        if (w.pt_ngood_cur[p] > 0) {
            const int h = w.pt_host[p];
            double bb = (double) w.pt_bdSumF[p];
            for (int c = 0; c < 4; c++) bb -= (-w.x[c]) * (double) w.pt_Hcd[p * 4 + c];
            const float4 *row = reinterpret_cast<const float4 *>(w.T[cur] + (size_t) p * N * T_STRIDE);
            const double *xa = s_xad + (size_t) h * N * 8;
#pragma unroll 4
            for (int t = 0; t < N; t++) {
                const float4 a = __ldg(row + t * (T_STRIDE / 4)), c = __ldg(row + t * (T_STRIDE / 4) + 1);
                bb -= xa[t * 8 + 0] * a.x + xa[t * 8 + 1] * a.y + xa[t * 8 + 2] * a.z + xa[t * 8 + 3] * a.w + xa[t * 8 + 4] * c.x + xa[t * 8 + 5] * c.y + xa[t * 8 + 6] * c.z + xa[t * 8 + 7] * c.w;
            }
            step = -bb * (double) w.pt_HdiF[p];
            if (!isfinite(step)) bad = 1;
        }
        w.pt_step[p] = step;
        const float backup = (float) w.pt_idepth[p];               // backupState (BA:919-922): idepth_backup is float
        w.pt_idepth_backup[p] = backup;
        const double newid = (double) backup + step;
        if (isfinite(newid) && newid > 0.0) {
            w.pt_idepth[p] = newid;
            w.pt_idepth_zero[p] = (float) newid;
            nid = (double) fabsf(backup); cntid = 1;
        }
        const float deltaF = (float) (w.pt_idepth[p] - (double) w.pt_idepth_zero[p]);   // computeDelta (BA:1180-1190) after the step
        pe = (double) (deltaF * deltaF * w.pt_priorF[p]);
    }
    __shared__ double s_n[8]; __shared__ double s_pe[8]; __shared__ int s_c[8]; __shared__ int s_bad[8];
    double sn = warp_sum_d(nid);
    const double spe = warp_sum_d(pe);
    int sc = cntid, sb = bad;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sc += __shfl_down_sync(0xffffffffu, sc, o); sb += __shfl_down_sync(0xffffffffu, sb, o); }
    if ((threadIdx.x & 31) == 0) { s_n[threadIdx.x >> 5] = sn; s_pe[threadIdx.x >> 5] = spe; s_c[threadIdx.x >> 5] = sc; s_bad[threadIdx.x >> 5] = sb; }
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        double tn = 0, tpe = 0; int tc = 0, tb = 0;
        for (int k = 0; k < 8; k++) { tn += s_n[k]; tpe += s_pe[k]; tc += s_c[k]; tb += s_bad[k]; }
        w.pt_part[blockIdx.x * 4 + 0] = tn; w.pt_part[blockIdx.x * 4 + 1] = (double) tc; w.pt_part[blockIdx.x * 4 + 2] = (double) tb; w.pt_part[blockIdx.x * 4 + 3] = tpe;
        __threadfence();
        const int ticket = atomicAdd(&ctrl->sc_done_count, 1);
        s_last = (ticket == (int) gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        double tn = 0, tc = 0, tb = 0, tpe = 0;
        for (int k = threadIdx.x; k < (int) gridDim.x; k += 32) { tn += w.pt_part[k * 4]; tc += w.pt_part[k * 4 + 1]; tb += w.pt_part[k * 4 + 2]; tpe += w.pt_part[k * 4 + 3]; }
        tn = warp_sum_d(tn); tc = warp_sum_d(tc); tb = warp_sum_d(tb); tpe = warp_sum_d(tpe);
    if (threadIdx.x == 0) {
        ctrl->prior_energy_pts = tpe;
        ctrl->sumNID = tn; ctrl->numID = (int) tc; ctrl->pt_bad = (int) tb;
        if (w.world > 1) { ctrl->sc_done_count = 0; return; }          // decided on the all-gathered sums in post_linearize_kernel
        const float sumNID = (float) (tn / (tc > 0 ? tc : 1.0));
        const float th = w.th_opt;
        ctrl->canbreak = (sqrtf(ctrl->sumA) < 0.0005f * th && sqrtf(ctrl->sumB) < 0.00005f * th && sqrtf(ctrl->sumR) < 0.00005f * th &&
                          sqrtf(ctrl->sumT) * sumNID < 0.00005f * th) ? 1 : 0;
        if (tb > 0) { ctrl->failed = 1; ctrl->done = 1; }            // BA:1489-1492
        ctrl->sc_done_count = 0;
    }
    }
}

// ------------------------------------------------------------------------------------------------
// After every linearization: total energy, setNewFrameEnergyTH (exact k-th smallest by radix select, the value
// std::nth_element leaves at index floor(0.7 n)), accept bookkeeping of run() (forceAccept path).
// mode 0: first linearization of run() (+applyActiveRes)  1: GN iteration  2: final linearizeAll(true)  3: stage call (no flip)
__global__ void __launch_bounds__(1024) post_linearize_kernel(const DevWin w, const int mode, const int respect_done) {
    KTRACE_ENTER(7);
    Ctrl *ctrl = w.ctrl;
    if (respect_done && launch_skipped(respect_done, ctrl->done, ctrl->final_done)) return;
    const int tid = threadIdx.x;
    __shared__ double s_red[32];
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_k, s_n;
    const bool multi = w.world > 1;
    const size_t rec_d = (size_t) w.post_stride;                        // doubles per rank record
    if (multi && w.p2p_post_on) {   // records were pushed into this rank's buffer by every rank's pack_post_kernel: wait for all of them
        __shared__ int s_ok;
        if (tid == 0) s_ok = 1;
        __syncthreads();
        if (tid < w.world) {
            const unsigned long long *f = p2p_flags(w, w.rank) + 16 + tid;
            unsigned long long v = 0; long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(f) : "memory");
                if (v >= w.p2p_post_epoch) break;
                if (++spins > (1l << 26)) { s_ok = 0; break; }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (!s_ok) { if (tid == 0) { ctrl->failed = 1; ctrl->pad0 = 1; ctrl->done = 1; } return; }
    }
    // energy: fixed-order sum of the block partials (multi-GPU: of the ranks' sums, all-gathered by pack_post_kernel)
    double e = 0.0;
    if (!multi) for (int i = tid; i < w.n_chunks; i += 1024) e += w.energy_part[i];
    else if (tid < w.world) e = w.post_recv[tid * rec_d];
    e = warp_sum_d(e);
    if ((tid & 31) == 0) s_red[tid >> 5] = e;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int k = 0; k < 32; k++) t += s_red[k]; s_red[0] = t; }
    // candidate list: local residuals towards the newest frame, or the gathered records
    const int c_begin = multi ? 0 : w.newest_begin, c_end = multi ? w.world * w.cand_cap : w.R;
    auto cand_at = [&](int i, bool &alive_i) -> float {
        if (!multi) { alive_i = w.r_alive[i] != 0; return w.r_new_energy_wo[i]; }
        const int rk = i / w.cand_cap, k = i - rk * w.cand_cap;
        alive_i = true;
        return reinterpret_cast<const float *>(w.post_recv + rk * rec_d + 8)[k];
    };
    // candidates: energies >= 0 of alive residuals whose target is the newest frame.  Each thread keeps up to PL_CACHE of
    // them in registers (one exposed load latency for the count + the four radix passes); longer tails are re-read.
    constexpr int PL_CACHE = 16;
    if (tid == 0) s_n = 0;
    __syncthreads();
    const double energy = s_red[0];
    float cache[PL_CACHE];
    unsigned int cnt = 0;
#pragma unroll
    for (int k = 0; k < PL_CACHE; k++) {
        const int i = c_begin + tid + k * 1024;
        bool al = false; float v = -1.f;
        if (i < c_end) v = cand_at(i, al);
        cache[k] = al ? v : -1.f;
    }
#pragma unroll
    for (int k = 0; k < PL_CACHE; k++) cnt += cache[k] >= 0.f ? 1u : 0u;
    for (int i = c_begin + tid + PL_CACHE * 1024; i < c_end; i += 1024) { bool al; const float v = cand_at(i, al); cnt += (al && v >= 0.f) ? 1u : 0u; }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0 && cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    const unsigned int nv = s_n;
    float th_new;
    if (nv == 0) {
        th_new = 12.f * 12.f * 8.f;                                  // BA:2432-2436
    } else {
        if (tid == 0) { s_k = (unsigned int) (int) (0.7f * (float) nv); s_prefix = 0; }   // nthIdx (BA:2448)
        __syncthreads();
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 24 - 8 * pass;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
            const unsigned int prefix = s_prefix;
            const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
            for (int k = 0; k < PL_CACHE; k++) {
                const unsigned int u = __float_as_uint(cache[k]);
                if (cache[k] >= 0.f && (u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
            }
            for (int i = c_begin + tid + PL_CACHE * 1024; i < c_end; i += 1024) {
                bool al; const float v = cand_at(i, al);
                if (al && v >= 0.f) {
                    const unsigned int u = __float_as_uint(v);
                    if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
                }
            }
            __syncthreads();
            if (tid < 32) {      // digit d with acc(d) <= k < acc(d) + hist[d]: warp scan over 8 digits per lane
                unsigned int loc[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { loc[q] = hist[tid * 8 + q]; sum += loc[q]; }
                unsigned int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
                unsigned int acc = incl - sum;
                const unsigned int k = s_k;
                __syncwarp();
                if (k >= acc && k < incl) {
                    int d = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) { if (k >= acc + loc[q]) { acc += loc[q]; d = q + 1; } else break; }
                    s_k = k - acc;
                    s_prefix = prefix | ((unsigned int) (tid * 8 + d) << shift);
                }
            }
            __syncthreads();
        }
        const float nth = sqrtf(__uint_as_float(s_prefix));          // BA:2455
        float th = nth * 1.5f;
        th = 26.0f * 0.5f + th * (1.f - 0.5f);
        th_new = th * th;
    }
    if (tid == 0) {
        ctrl->energy_new = energy;
        if (multi && mode == 1) {
            // doStepFromBackup's convergence test (BA:1013-1026) and failure flag on the sums over all ranks
            double tn = 0, tc = 0, tb = 0, tpe = 0;
            for (int rk = 0; rk < w.world; rk++) { const double *h = w.post_recv + rk * rec_d; tn += h[1]; tc += h[2]; tb += h[3]; tpe += h[4]; }
            const float sumNID = (float) (tn / (tc > 0 ? tc : 1.0));
            const float th = w.th_opt;
            ctrl->canbreak = (sqrtf(ctrl->sumA) < 0.0005f * th && sqrtf(ctrl->sumB) < 0.00005f * th && sqrtf(ctrl->sumR) < 0.00005f * th &&
                              sqrtf(ctrl->sumT) * sumNID < 0.00005f * th) ? 1 : 0;
            ctrl->prior_energy_pts = tpe;
            if (tb > 0) { ctrl->failed = 1; ctrl->done = 1; }
        }
        // calcLEnergy (BA:2118-2208) without linearized residuals: frame priors + point priors; calcMEnergy (BA:2095-2116) is 0 (H_M = 0)
        double EL = 0.0;
        if (!w.force_accept && w.has_HM) {   // calcMEnergy (BA:2095-2116): |delta . (2 b_M + H_M delta)|
            double em = 0.0;
            for (int r = 0; r < w.n; r++) {
                double row = 2.0 * w.bM[r];
                for (int c = 4; c < w.n; c++) { const FrameDev &g = w.frames[(c - 4) >> 3]; const int k = (c - 4) & 7; row += w.HM[(size_t) r * w.n + c] * (g.state[k] - g.state_zero[k]); }
                if (r >= 4) { const FrameDev &g = w.frames[(r - 4) >> 3]; const int k = (r - 4) & 7; em += (g.state[k] - g.state_zero[k]) * row; }
            }
            EL += fabs(em);
        }
        if (!w.force_accept) {
            for (int i = 0; i < w.N; i++) for (int k = 0; k < 8; k++) { const double d = w.frames[i].state[k]; EL += d * w.frames[i].prior[k] * d; }
            EL += ctrl->prior_energy_pts;
        }
        ctrl->energyL_new = EL;
        bool keep_th = false;
        if (!isfinite(energy)) { ctrl->failed = 1; ctrl->done = 1; }
        else if (mode == 0) { ctrl->cur ^= 1; ctrl->energy_last = energy; ctrl->energy_first = energy; ctrl->energyL_last = EL; }
        else if (mode == 1) {
            const int it = ctrl->iteration;
            if (w.force_accept || energy + EL < ctrl->energy_last + ctrl->energyL_last) {
                // applyActiveRes, lambda *= 0.25 (BA:843-864)
                ctrl->cur ^= 1; ctrl->energy_last = energy; ctrl->energyL_last = EL; ctrl->lambda *= 0.25; ctrl->accepted += 1;
            } else {
                // loadSateBackup + re-linearization at the backup = the committed linearization stays (BA:866-876); restore_state_kernel undoes the step
                ctrl->lambda *= 1e2; ctrl->rejected += 1; ctrl->rejected_at = it + 1; keep_th = true;
            }
            ctrl->iteration = it + 1;
            if (ctrl->canbreak && it >= 1) ctrl->done = 1;             // BA:879
        } else if (mode == 2) { ctrl->cur ^= 1; ctrl->energy_last = energy; ctrl->final_done = 1; }
        if (!keep_th) w.frames[w.N - 1].energy_th = th_new;            // a rejected linearization leaves frameEnergyTH as the re-linearization would
    }
}

// multi-GPU: this rank's record for the all-gather that precedes post_linearize_kernel (see DevWin::post_send)
__global__ void __launch_bounds__(1024) pack_post_kernel(const DevWin w, const int respect_done) {
    KTRACE_ENTER(8);
    Ctrl *ctrl = w.ctrl;
    (void) respect_done;   // always packs: every rank must feed the collective, even after its own early exit
    const int tid = threadIdx.x;
    __shared__ double s_red[32];
    double e = 0.0;
    for (int i = tid; i < w.n_chunks; i += 1024) e += w.energy_part[i];
    e = warp_sum_d(e);
    if ((tid & 31) == 0) s_red[tid >> 5] = e;
    __syncthreads();
    const int nloc = w.R - w.newest_begin;
    if (w.p2p_post_on) {
        // push this rank's record into every rank's buffer (fire-and-forget NVLink stores), then publish the epoch
        for (int q = 0; q < w.world; q++) {
            double *h = p2p_post_slot(w, q, w.rank);
            if (tid == 0) {
                double t = 0; for (int k = 0; k < 32; k++) t += s_red[k];
                h[0] = t; h[1] = ctrl->sumNID; h[2] = (double) ctrl->numID; h[3] = (double) ctrl->pt_bad; h[4] = ctrl->prior_energy_pts; h[5] = h[6] = h[7] = 0.0;
            }
            float *c = reinterpret_cast<float *>(h + 8);
            for (int k = tid; k < w.cand_cap; k += 1024) {
                float v = -1.f;
                if (k < nloc) { const int i = w.newest_begin + k; v = w.r_alive[i] ? w.r_new_energy_wo[i] : -1.f; }
                c[k] = v;
            }
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) for (int q = 0; q < w.world; q++) {
            unsigned long long *f = p2p_flags(w, q) + 16 + w.rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(f), "l"(w.p2p_post_epoch) : "memory");
        }
        return;
    }
    if (tid == 0) {
        double t = 0; for (int k = 0; k < 32; k++) t += s_red[k];
        double *h = w.post_send;
        h[0] = t; h[1] = ctrl->sumNID; h[2] = (double) ctrl->numID; h[3] = (double) ctrl->pt_bad; h[4] = ctrl->prior_energy_pts; h[5] = h[6] = h[7] = 0.0;
    }
    float *c = reinterpret_cast<float *>(w.post_send + 8);
    for (int k = tid; k < w.cand_cap; k += 1024) {
        float v = -1.f;
        if (k < nloc) { const int i = w.newest_begin + k; v = w.r_alive[i] ? w.r_new_energy_wo[i] : -1.f; }
        c[k] = v;
    }
}

// applyRes without the bookkeeping of run(): the candidate linearization becomes the committed one
__global__ void commit_candidate_kernel(const DevWin w) { if (threadIdx.x == 0 && blockIdx.x == 0) w.ctrl->cur ^= 1; }

// forceAccept = false: undo a rejected step (loadSateBackup, BA:928-946).  `it1` = the iteration counter value the step belongs to.
__global__ void __launch_bounds__(256) restore_state_kernel(const DevWin w, const int it1) {
    KTRACE_ENTER(9);
    Ctrl *ctrl = w.ctrl;
    if (ctrl->rejected_at != it1) return;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p < w.P) { const float b = w.pt_idepth_backup[p]; w.pt_idepth[p] = (double) b; w.pt_idepth_zero[p] = b; }
    if (blockIdx.x == 0) {
        if (threadIdx.x < w.N) { FrameDev &f = w.frames[threadIdx.x]; double st[10]; for (int k = 0; k < 10; k++) st[k] = f.state_backup[k]; frame_set_state(f, st, w); }
        __syncthreads();
        for (int e = threadIdx.x; e < w.N * w.N; e += 256) pair_precompute(w, e / w.N, e % w.N);
        if (threadIdx.x == 0) {
            double pe = 0.0;   // point prior energy of the restored state is recomputed lazily: deltaF = 0 after loadSateBackup (idepth_zero = idepth_backup)
            ctrl->prior_energy_pts = pe;
        }
    }
}

// sum_p deltaF^2 * priorF at the start of run() (forceAccept = false only)
__global__ void __launch_bounds__(256) point_prior_energy_kernel(const DevWin w) {
    __shared__ double s[8];
    double pe = 0.0;
    for (int p = threadIdx.x; p < w.P; p += 256) { const float d = (float) (w.pt_idepth[p] - (double) w.pt_idepth_zero[p]); pe += (double) (d * d * w.pt_priorF[p]); }
    pe = warp_sum_d(pe);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = pe;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int k = 0; k < 8; k++) t += s[k]; w.ctrl->prior_energy_pts = t; }
}

// ------------------------------------------------------------------------------------------------
// addPoints (BA:382-415, DSOContext.h:86-91): reference colours by INTEGER pixel read, weights from the
// interpolated host gradient.  One thread per (point, pattern pixel).
__global__ void point_init_kernel(const DevWin w, const int p_begin, const int p_count, float *colors, float *weights) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_count * 8) return;
    const int p = p_begin + (i >> 3), k = i & 7;
    const float4 *img = w.img[w.pt_host[p]];
    const float x = w.pt_x[p], y = w.pt_y[p];
    const int ix0 = (int) x + c_sx[k], iy0 = (int) y + c_sy[k];
    colors[(size_t) p * 8 + k] = img[(size_t) iy0 * w.W + ix0].x;
    const float fx = x + (float) c_sx[k], fy = y + (float) c_sy[k];
    const int ix = (int) fx, iy = (int) fy;
    const float dx = fx - (float) ix, dy = fy - (float) iy, dxdy = dx * dy;
    const float w00 = 1.f - dx - dy + dxdy, w10 = dx - dxdy, w01 = dy - dxdy, w11 = dxdy;
    const float4 *q = img + (size_t) iy * w.W + ix;
    const float4 a = q[0], b = q[1], c = q[w.W], d = q[w.W + 1];
    const float gx = a.y * w00 + b.y * w10 + c.y * w01 + d.y * w11;
    const float gy = a.z * w00 + b.z * w10 + c.z * w01 + d.z * w11;
    const double g2 = (double) gx * (double) gx + (double) gy * (double) gy;
    weights[(size_t) p * 8 + k] = (float) sqrt((double) w.cth / ((double) w.cth + g2));
}

// run() prologue on the device: resetOOB on every residual (BA:766-779, DSOResidual.h:81-86), empty Schur tables
__global__ void __launch_bounds__(256) reset_window_kernel(const DevWin w) {
    const size_t tid = blockIdx.x * (size_t) blockDim.x + threadIdx.x, nth = (size_t) gridDim.x * blockDim.x;
    const size_t nT4 = (size_t) w.P * w.N * (T_STRIDE / 4);
    float4 *T0 = reinterpret_cast<float4 *>(w.T[0]), *T1 = reinterpret_cast<float4 *>(w.T[1]);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < nT4; i += nth) { T0[i] = z; T1[i] = z; }
    for (size_t i = tid; i < (size_t) w.R; i += nth) {
        w.r_state[0][i] = RES_IN; w.r_state[1][i] = RES_IN; w.r_energy[0][i] = 0.f; w.r_energy[1][i] = 0.f; w.r_good[0][i] = 0; w.r_good[1][i] = 0;
        w.r_new_state[i] = RES_OUTLIER; w.r_new_energy[i] = 0.f; w.r_new_energy_wo[i] = 0.f; w.r_alive[i] = 1;
        w.r_center[3 * i] = 0.f; w.r_center[3 * i + 1] = 0.f; w.r_center[3 * i + 2] = 0.f;
    }
    for (size_t i = tid; i < (size_t) w.P; i += nth) w.pt_ngood_cur[i] = 0;
    if (w.dbg) for (size_t i = tid; i < (size_t) w.R * DBG_STRIDE; i += nth) w.dbg[i] = 0.f;
}

// Level-0 derivative image straight from the rectified gray image (CaptureImageGenerator::generate, capture/CaptureImage.cpp:249;
// Array2D::gradientImage / gradient, image/Array2D.h:288-294, 314-331): (I, (I[x+1]-I[x-1])*0.5, (I[y+1]-I[y-1])*0.5), all three
// channels zero on the 1-pixel border.  Bit-exact with the reference's GradientImage; a third of the upload volume.
__global__ void gradient_texel_kernel(const float *__restrict__ gray, float4 *__restrict__ dst, const int W, const int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x >= 1 && y >= 1 && x < W - 1 && y < H - 1) {
        const float *p = gray + (size_t) y * W + x;
        t.x = p[0];
        t.y = (p[1] - p[-1]) * 0.5f;
        t.z = (p[W] - p[-W]) * 0.5f;
    }
    dst[(size_t) y * W + x] = t;
}

// AoS (I,dx,dy) 12-byte texels -> float4 texels (one aligned 128-bit load per bilinear tap)
__global__ void repack_image_kernel(const float *__restrict__ src, float4 *__restrict__ dst, const int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

// writes >L2-size bytes: used by the benchmark to evict the window between timed passes
__global__ void l2_flush_kernel(float4 *buf, const size_t n4) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

}  // namespace cmlba
