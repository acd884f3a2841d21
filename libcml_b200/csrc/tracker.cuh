// tracker.cuh -- device code of the coarse tracker (SURVEY.md 8f NEXT #1): pyramid, coarse inverse-depth maps and the
// single-launch coarse-to-fine direct image alignment.
//
// Reference anchors (under /root/reference/src/cml):
//   pyr_gray_kernel / pyr_grad_kernel   capture/CaptureImage.cpp:209-262, image/Array2D.h:388-401 (reduceByTwo), :288-331 (gradientImage)
//   cd_* kernels                        optimization/dso/DSOTracker.cpp:494-725 (makeCoarseDepthL0)
//   track_kernel                        optimization/dso/DSOTracker.cpp:15-246 (optimize), :248-419 (computeResidual), :421-492 (computeHessian)
//
// B200 design.  The reference alternates a residual pass, a Hessian pass over a staged "warped" buffer and an 8x8 solve,
// ~25-60 times per frame, over a few thousand points: microseconds of work per step, so on a GPU the cost is launches and
// round trips, not arithmetic.  track_kernel therefore runs the WHOLE optimisation (all levels, all iterations, solves,
// accept/reject, level repeat, final checks) in one launch: one thread-block cluster per start pose; every CTA walks an
// interleaved share of the level's point list, accumulates energy, counters and the 44 Hessian sums of the candidate pose in
// registers in the same pass (the warped buffer never exists), reduces them with a transposing warp butterfly, and pushes
// its 64 partial sums into every CTA of the cluster through distributed shared memory; one cluster barrier later each CTA
// holds identical totals and thread 0 of each CTA advances an identical copy of the optimiser state machine.  No global
// memory traffic besides the point list and the image taps, no atomics, fixed summation order (deterministic).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "se3.h"

namespace cmltrk {
namespace cg = cooperative_groups;
using cmlba::Pose;

constexpr int MAXL = 6;            // pyramid levels held
constexpr int OPTL = 5;            // levels optimised (DSOTracker.cpp:23-24)
constexpr int TRK_SUMS = 64;       // E, 3 counters, 3 flow sums, pad, 44 Hessian sums, pad
constexpr int TRK_MAX_CLUSTER = 16;
constexpr int PYR_TILE = 32;
constexpr double FIX_ONE = 4294967296.0;   // 2^32: fixed-point unit of the splat maps (integer atomics => order independent)

struct LevelDev {
    int w, h;
    float fx, fy, cx, cy;
    const float4 *pc;      // (u, v, idepth, colour) of the reference keyframe, raster order
    const int *pc_n;
    const float4 *grad;    // (I, dx, dy, 0) of the frame to track
};

struct TrackParams {
    LevelDev lv[OPTL];
    int max_level;
    float huber, cutoff;
    double scale[8];
    int optimize_a, optimize_b;
    double sat_th;
    double ref_tau, ref_a, ref_b, new_tau;
    int has_last;
    double last_rmse[OPTL];
};

struct Candidate {
    double R[9], t[3];     // start refToNew
    double a, b;           // start brightness of the frame to track
};

struct TrackOut {
    double R[9], t[3], a, b;
    double E[OPTL];
    int nT[OPTL], nS[OPTL], nR[OPTL];
    double rep[OPTL];
    double flow[3], rel_aff[2], cov[6];
    int is_correct, sat_ok, iterations, evals;
    long long cyc_advance, cyc_eval, cyc_reduce;      // SM cycles of CTA 0 / thread 0 per phase (where the launch spends its time)
};

// ------------------------------------------------------------------------------------------------ pyramid
// One CTA per 32x32 level-0 tile computes that tile of every coarser level in shared memory: level l pixel = ((a + b) + c) + d) / 4
// of level l-1 (fp32, this order: bit-identical to reduceByTwo).
struct PyrDev {
    int levels;
    int w[MAXL], h[MAXL];
    float *gray[MAXL];
    float4 *grad[MAXL];
};

__global__ void __launch_bounds__(256) pyr_gray_kernel(const PyrDev p) {
    __shared__ float buf[2][PYR_TILE][PYR_TILE + 1];      // ping-pong: level l reads buf[(l - 1) & 1], writes buf[l & 1]
    const int tx0 = blockIdx.x * PYR_TILE, ty0 = blockIdx.y * PYR_TILE, tid = threadIdx.x;
    for (int k = tid; k < PYR_TILE * PYR_TILE; k += 256) {
        const int x = k % PYR_TILE, y = k / PYR_TILE, gx = tx0 + x, gy = ty0 + y;
        buf[0][y][x] = (gx < p.w[0] && gy < p.h[0]) ? p.gray[0][(size_t) gy * p.w[0] + gx] : 0.f;
    }
    __syncthreads();
    int size = PYR_TILE;
    for (int l = 1; l < p.levels; l++) {
        size >>= 1;
        const int ox = tx0 >> l, oy = ty0 >> l;
        float (*src)[PYR_TILE + 1] = buf[(l - 1) & 1];
        float (*dst)[PYR_TILE + 1] = buf[l & 1];
        for (int k = tid; k < size * size; k += 256) {
            const int x = k % size, y = k / size;
            const float v = (((src[2 * y][2 * x] + src[2 * y][2 * x + 1]) + src[2 * y + 1][2 * x]) + src[2 * y + 1][2 * x + 1]) / 4.f;
            dst[y][x] = v;
            if (ox + x < p.w[l] && oy + y < p.h[l]) p.gray[l][(size_t) (oy + y) * p.w[l] + ox + x] = v;
        }
        __syncthreads();
    }
}

// derivative texels (I, dx, dy, 0) of every level; the border ring is zero (gradientImage leaves it untouched)
__global__ void __launch_bounds__(256) pyr_grad_kernel(const PyrDev p) {
    const int l = blockIdx.y;
    const int w = p.w[l], h = p.h[l];
    const float *g = p.gray[l];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < w * h; i += gridDim.x * 256) {
        const int x = i % w, y = i / w;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x > 0 && y > 0 && x < w - 1 && y < h - 1) {
            o.x = g[i];
            o.y = (g[i + 1] - g[i - 1]) * 0.5f;
            o.z = (g[i + w] - g[i - w]) * 0.5f;
        }
        p.grad[l][i] = o;
    }
}

// ------------------------------------------------------------------------------------------------ coarse depth
struct CoarseDev {
    int levels;
    int w[MAXL], h[MAXL];
    long long *mapI[MAXL], *mapW[MAXL];   // fixed-point sums of idepth * weight and weight
    const float *gray[MAXL];              // reference keyframe
    float4 *pc[MAXL];
    int *pc_n;                            // [levels]
    int *row_count[MAXL], *row_offset[MAXL];
    int row_base[MAXL + 1];               // rows of level l = blockIdx range [row_base[l], row_base[l + 1])
    double fx, fy, cx, cy;
};

// first loop of makeCoarseDepthL0 (DSOTracker.cpp:520-553), fp64 like the reference
__global__ void __launch_bounds__(256) cd_project_kernel(const CoarseDev c, const int P, const double *__restrict__ rel /* [F][12] host-to-reference */,
                                                        const int *__restrict__ pt_frame, const float2 *__restrict__ pt_xy, const double *__restrict__ pt_idepth,
                                                        const double *__restrict__ pt_unc) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const double *T = rel + (size_t) pt_frame[i] * 12;
    const float2 xy = pt_xy[i];
    const double id = pt_idepth[i];
    const double px = ((double) xy.x - c.cx) / c.fx, py = ((double) xy.y - c.cy) / c.fy;
    const double qx = T[0] * px + T[1] * py + T[2] + T[9] * id, qy = T[3] * px + T[4] * py + T[5] + T[10] * id, qz = T[6] * px + T[7] * py + T[8] + T[11] * id;
    const double Ku = c.fx * (qx / qz) + c.cx, Kv = c.fy * (qy / qz) + c.cy;
    const double new_id = (1.0 / qz) * id;
    if (!(Ku + 0.5 > -2147483000.0 && Ku + 0.5 < 2147483000.0 && Kv + 0.5 > -2147483000.0 && Kv + 0.5 < 2147483000.0)) return;   // int conversion would be undefined
    const int u = (int) (Ku + 0.5), v = (int) (Kv + 0.5);
    if (u < 0 || u >= c.w[0] || v < 0 || v >= c.h[0]) return;
    const float weight = sqrtf((float) (1e-3 / (pt_unc[i] + 1e-12)));
    const float contrib = (float) (new_id * (double) weight);
    atomicAdd(reinterpret_cast<unsigned long long *>(c.mapI[0] + (size_t) v * c.w[0] + u), (unsigned long long) __double2ll_rn((double) contrib * FIX_ONE));
    atomicAdd(reinterpret_cast<unsigned long long *>(c.mapW[0] + (size_t) v * c.w[0] + u), (unsigned long long) __double2ll_rn((double) weight * FIX_ONE));
}

// 2x2 sums of level l-1 (DSOTracker.cpp:555-585); exact in fixed point
__global__ void __launch_bounds__(256) cd_downsum_kernel(const CoarseDev c, const int l) {
    const int w = c.w[l], h = c.h[l], wm = c.w[l - 1];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < w * h; i += gridDim.x * 256) {
        const int x = i % w, y = i / w;
        const size_t b = (size_t) 2 * x + (size_t) 2 * y * wm;
        c.mapI[l][i] = c.mapI[l - 1][b] + c.mapI[l - 1][b + 1] + c.mapI[l - 1][b + wm] + c.mapI[l - 1][b + wm + 1];
        c.mapW[l][i] = c.mapW[l - 1][b] + c.mapW[l - 1][b + 1] + c.mapW[l - 1][b + wm] + c.mapW[l - 1][b + wm + 1];
    }
}

__device__ __forceinline__ float fix2f(long long v) { return (float) ((double) v * (1.0 / FIX_ONE)); }

// dilation (DSOTracker.cpp:588-668) and normalisation (:671-716) of one interior pixel; true if it becomes a point
__device__ __forceinline__ bool cd_pixel(const CoarseDev &c, const int l, const int x, const int y, float4 &out) {
    const int wl = c.w[l];
    const int i = x + y * wl;
    const long long *I = c.mapI[l], *W = c.mapW[l];
    float wsum = fix2f(W[i]), idsum = fix2f(I[i]);
    if (!(wsum > 0.f)) {
        const int o0 = (l < 2) ? wl + 1 : 1, o1 = (l < 2) ? wl - 1 : wl;     // neighbours in the reference's order: +o0, -o0, +o1, -o1
        const int offs[4] = {o0, -o0, o1, -o1};
        float sum = 0.f, num = 0.f, numn = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float wn = fix2f(W[i + offs[k]]);
            if (wn > 0.f) { sum += fix2f(I[i + offs[k]]); num += wn; numn += 1.f; }
        }
        if (numn > 0.f) { idsum = sum / numn; wsum = num / numn; }
    }
    if (!(wsum > 0.f)) return false;
    const float idepth = idsum / wsum;
    const float col = c.gray[l][i];
    if (!isfinite(col) || !(idepth > 0.f)) return false;
    out = make_float4((float) x, (float) y, idepth, col);
    return true;
}

// rows [2, h-2) of every level: pass 0 counts the points of a row, pass 1 emits them at row_offset (raster order like the reference)
template <int kPass>
__global__ void __launch_bounds__(128) cd_rows_kernel(const CoarseDev c) {
    int l = 0;
    while (l + 1 < c.levels && (int) blockIdx.x >= c.row_base[l + 1]) l++;
    const int y = (int) blockIdx.x - c.row_base[l] + 2;
    const int wl = c.w[l];
    __shared__ int s_warp[4];
    __shared__ int s_run;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_run = kPass ? c.row_offset[l][y] : 0;
    __syncthreads();
    for (int x0 = 2; x0 < wl - 2; x0 += 128) {
        const int x = x0 + tid;
        float4 rec;
        const bool ok = (x < wl - 2) && cd_pixel(c, l, x, y, rec);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        if (kPass) {
            int base = s_run;
            for (int k = 0; k < warp; k++) base += s_warp[k];
            if (ok) c.pc[l][base + __popc(m & ((1u << lane) - 1u))] = rec;
        }
        __syncthreads();
        if (tid == 0) s_run += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        __syncthreads();
    }
    if (!kPass && tid == 0) c.row_count[l][y] = s_run;
}

// exclusive scan of the row counts of level blockIdx.x
__global__ void __launch_bounds__(32) cd_rowscan_kernel(const CoarseDev c) {
    const int l = blockIdx.x, lane = threadIdx.x;
    int run = 0;
    for (int y0 = 2; y0 < c.h[l] - 2; y0 += 32) {
        const int y = y0 + lane;
        const int v = (y < c.h[l] - 2) ? c.row_count[l][y] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (y < c.h[l] - 2) c.row_offset[l][y] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) c.pc_n[l] = run;
}

// ------------------------------------------------------------------------------------------------ tracking
// sum over the warp of 32 per-lane values v[0..31]; lane L returns the total of entry L (31 shuffles instead of 160)
__device__ __forceinline__ float transpose_sum(float *v, const int lane) {
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

struct EvalCmd {
    float RKi[9], Ki[9], t[3];
    float aLL, bLL, b0;
    float cut, maxE;
    int level, exit;
};

struct TrkState {
    EvalCmd cmd;
    Pose cur, cand;
    double a, b, an, bn;
    double Hb[2][72];     // [hsel]: Hessian (64) and gradient (8) of the accepted pose; [hsel ^ 1]: those of the pose just evaluated
    int hsel;
    double inc[8];
    double lambda;
    int it, level, phase, have_repeated, iterations, fail;
    double rep[OPTL];
    double oE[OPTL], nE[OPTL];
    int oT[OPTL], oS[OPTL], oR[OPTL], nT[OPTL], nS[OPTL], nR[OPTL];
    double oFlow[3], nFlow[3];
};

// slot of sum_i w J_a J_b (a <= b, a < 8; J_8 = residual): rows of the upper triangle of the 9x9 accumulator, packed
__host__ __device__ constexpr int h_index(int a, int b) { return 8 + a * 9 - (a * (a - 1)) / 2 + (b - a); }
enum { S_E = 0, S_NT = 1, S_NSAT = 2, S_NROB = 3, S_FT = 4, S_FRT = 5, S_FNUM = 6, S_H = 8 };

// 8x8 LDL^T kept in registers (every index is a compile-time constant after unrolling; only the lower triangle exists).
// Dimensions with active[i] == 0 are decoupled (row/column dropped, x[i] = 0): identical to solving the reference's 7x7 / 6x6
// sub-systems (DSOTracker.cpp:98-121).
struct Ldlt8 {
    double A[36];      // packed lower triangle, row r column c at r (r + 1) / 2 + c; after factor(): unit L below, 1 / D on the diagonal
    __device__ __forceinline__ static constexpr int at(int r, int c) { return r * (r + 1) / 2 + c; }
    __device__ __forceinline__ void factor(const double *H, const double damp, const unsigned active) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c <= r; c++) {
                const bool on = ((active >> r) & 1u) && ((active >> c) & 1u);
                A[at(r, c)] = (r == c) ? (on ? H[r * 9] * damp : 1.0) : (on ? H[r * 8 + c] : 0.0);
            }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double inv = 1.0 / A[at(k, k)];
#pragma unroll
            for (int r = k + 1; r < 8; r++) {
                const double f = A[at(r, k)] * inv;
#pragma unroll
                for (int c = k + 1; c <= r; c++) A[at(r, c)] -= f * A[at(c, k)];
            }
#pragma unroll
            for (int r = k + 1; r < 8; r++) A[at(r, k)] *= inv;
            A[at(k, k)] = inv;
        }
    }
    __device__ __forceinline__ void solve(double *y) const {      // in place
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < r; c++) y[r] -= A[at(r, c)] * y[c];
#pragma unroll
        for (int r = 0; r < 8; r++) y[r] *= A[at(r, r)];
#pragma unroll
        for (int r = 7; r >= 0; r--)
#pragma unroll
            for (int c = r + 1; c < 8; c++) y[r] -= A[at(c, r)] * y[c];
    }
};

__device__ __forceinline__ void exposure_to(const double a0, const double b0, const double t0, const double a1, const double b1, const double t1, double &a, double &b) {
    a = exp(a1 - a0) * t1 / t0;       // Exposure::to (map/Exposure.h)
    b = b1 - a * b0;
}

__device__ void request_eval(TrkState &S, const TrackParams &P, const Pose &T, const double a, const double b, const double cutoff) {
    EvalCmd &c = S.cmd;
    const LevelDev &L = P.lv[S.level];
    c.level = S.level; c.exit = 0;
    const float ifx = 1.f / L.fx, ify = 1.f / L.fy;
    const float Ki[9] = {ifx, 0.f, -L.cx * ifx, 0.f, ify, -L.cy * ify, 0.f, 0.f, 1.f};
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) {
            c.Ki[r * 3 + k] = Ki[r * 3 + k];
            c.RKi[r * 3 + k] = (float) T.R[r * 3] * Ki[k] + (float) T.R[r * 3 + 1] * Ki[3 + k] + (float) T.R[r * 3 + 2] * Ki[6 + k];
        }
    for (int k = 0; k < 3; k++) c.t[k] = (float) T.t[k];
    double al, bl;
    exposure_to(P.ref_a, P.ref_b, P.ref_tau, a, b, P.new_tau, al, bl);
    c.aLL = (float) al; c.bLL = (float) bl; c.b0 = (float) P.ref_b;
    c.cut = (float) cutoff;
    c.maxE = 2.0f * P.huber * c.cut - P.huber * P.huber;
}

// scaled Hessian / gradient entry e (0..63: H row-major, 64..71: g) of the pose just evaluated (computeHessian's epilogue, DSOTracker.cpp:466-491);
// one thread per entry
__device__ __forceinline__ double hessian_entry(const int e, const TrackParams &P, const double *sum) {
    const int nw = (int) (sum[S_NT] - sum[S_NSAT]);
    const double n = (double) ((nw + 3) & ~3);
    const int i = e < 64 ? e >> 3 : e - 64, j = e < 64 ? e & 7 : 8;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const double v = sum[8 + a * 9 - (a * (a - 1)) / 2 + (b - a)] / n * P.scale[i];
    return e < 64 ? v * P.scale[j] : v;
}

__device__ void finish(TrkState &S, const bool converged) {
    S.cmd.exit = 1;
    S.fail = converged ? 0 : 1;
}

// one Gauss-Newton proposal from (H, g, lambda): DSOTracker.cpp:93-160
__device__ void propose(TrkState &S, const TrackParams &P) {
    unsigned active = 0x3fu;
    if (P.optimize_a) active |= 1u << 6;
    if (P.optimize_b) active |= 1u << 7;
    Ldlt8 F;
    const double *H = S.Hb[S.hsel], *g = H + 64;
    F.factor(H, 1.0 + S.lambda, active);
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = ((active >> i) & 1u) ? -g[i] : 0.0;
    F.solve(x);
#pragma unroll
    for (int i = 0; i < 8; i++) S.inc[i] = x[i];
    bool finite = true;
    for (int i = 0; i < 8; i++) finite = finite && isfinite(S.inc[i]);
    if (!finite) { finish(S, false); return; }
    if (S.lambda < 0.001) {
        const double f = sqrt(sqrt(0.001 / S.lambda));
        for (int i = 0; i < 8; i++) S.inc[i] *= f;
    }
    double xs[8];
    for (int i = 0; i < 8; i++) xs[i] = S.inc[i] * P.scale[i];
    const Pose d = cmlba::se3_exp(xs);
    S.cand = cmlba::pose_mul(d, S.cur);
    S.an = S.a + xs[6]; S.bn = S.b + xs[7];
    request_eval(S, P, S.cand, S.an, S.bn, (double) P.cutoff * S.rep[S.level]);
    S.phase = 2;
}

__device__ void begin_level(TrkState &S, const TrackParams &P) {
    S.rep[S.level] = 1.0;
    request_eval(S, P, S.cur, S.a, S.b, (double) P.cutoff * S.rep[S.level]);
    S.phase = 1;
}

__device__ void end_level(TrkState &S, const TrackParams &P) {
    if (P.has_last && S.oE[S.level] / (double) S.oT[S.level] > 1.5 * P.last_rmse[S.level]) { finish(S, false); return; }
    if (S.rep[S.level] > 1.0 && !S.have_repeated) { S.level++; S.have_repeated = 1; }
    S.level--;
    if (S.level < 0) { finish(S, true); return; }
    begin_level(S, P);
}

__device__ __noinline__ void advance(TrkState &S, const TrackParams &P, const double *sum) {
    const int maxIt[OPTL] = {10, 20, 50, 50, 50};
    if (S.phase == 0) { S.level = P.max_level; begin_level(S, P); return; }
    const int lv = S.level;
    const double fl0 = sum[S_FT] / (sum[S_FNUM] + 0.1), fl2 = sum[S_FRT] / (sum[S_FNUM] + 0.1);
    if (S.phase == 1) {
        S.oE[lv] = sum[S_E]; S.oT[lv] = (int) sum[S_NT]; S.oS[lv] = (int) sum[S_NSAT]; S.oR[lv] = (int) sum[S_NROB];
        S.oFlow[0] = fl0; S.oFlow[1] = 0.0; S.oFlow[2] = fl2;
        if (S.oT[lv] < 20) { finish(S, false); return; }
        if ((double) S.oS[lv] / (double) S.oT[lv] > 0.6 && S.rep[lv] < 50.0) {
            S.rep[lv] *= 2.0;
            request_eval(S, P, S.cur, S.a, S.b, (double) P.cutoff * S.rep[lv]);
            return;
        }
        if (S.oT[lv] - S.oS[lv] < 10) { finish(S, false); return; }
        S.hsel ^= 1;        // the Hessian of this evaluation (computed by 72 threads after the reduction) becomes the current one
        S.lambda = 0.01; S.it = 0;
        propose(S, P);
        return;
    }
    // phase 2: a proposal has been evaluated
    S.nE[lv] = sum[S_E]; S.nT[lv] = (int) sum[S_NT]; S.nS[lv] = (int) sum[S_NSAT]; S.nR[lv] = (int) sum[S_NROB];
    S.nFlow[0] = fl0; S.nFlow[1] = 0.0; S.nFlow[2] = fl2;
    S.iterations++;
    const bool accept = (S.nE[lv] / (double) S.nT[lv]) < (S.oE[lv] / (double) S.oT[lv]);
    if (accept) {
        S.hsel ^= 1;
        // `oldResidual = newResidual` copies every level: coarser levels inherit the last tried step there (DSOTracker.cpp:166)
        for (int l = 0; l < OPTL; l++) { S.oE[l] = S.nE[l]; S.oT[l] = S.nT[l]; S.oS[l] = S.nS[l]; S.oR[l] = S.nR[l]; }
        for (int k = 0; k < 3; k++) S.oFlow[k] = S.nFlow[k];
        S.cur = S.cand; S.a = S.an; S.b = S.bn;
        S.lambda *= 0.5;
    } else {
        S.lambda *= 4.0;
    }
    S.it++;
    double n2 = 0.0;
    for (int i = 0; i < 8; i++) n2 += S.inc[i] * S.inc[i];
    if (sqrt(n2) < 1e-3 || S.it >= maxIt[lv]) end_level(S, P);
    else propose(S, P);
}

__device__ __noinline__ void write_out(const TrkState &S, const TrackParams &P, TrackOut &o) {
    for (int i = 0; i < 9; i++) o.R[i] = S.cur.R[i];
    for (int i = 0; i < 3; i++) o.t[i] = S.cur.t[i];
    o.a = S.a; o.b = S.b;
    for (int l = 0; l < OPTL; l++) { o.E[l] = S.oE[l]; o.nT[l] = S.oT[l]; o.nS[l] = S.oS[l]; o.nR[l] = S.oR[l]; o.rep[l] = S.rep[l]; }
    for (int k = 0; k < 3; k++) o.flow[k] = S.oFlow[k];
    o.iterations = S.iterations; o.evals = 0;
    o.is_correct = 0; o.sat_ok = 1; o.rel_aff[0] = o.rel_aff[1] = 0.0;
    for (int k = 0; k < 6; k++) o.cov[k] = 999999.0;
    if (S.fail) return;
    double ra, rb;
    exposure_to(P.ref_a, P.ref_b, P.ref_tau, S.a, S.b, P.new_tau, ra, rb);
    bool good = true;
    if (P.optimize_a) { if (fabs(S.a) > 1.2) good = false; }
    else if (fabs(logf((float) ra)) > 1.5f) good = false;
    if (P.optimize_b) { if (fabs(S.b) > 200.0) good = false; }
    else if (fabsf((float) rb) > 200.f) good = false;
    o.is_correct = good ? 1 : 0;
    o.sat_ok = ((double) S.oS[0] / (double) S.oT[0] > P.sat_th) ? 0 : 1;
    o.rel_aff[0] = ra; o.rel_aff[1] = rb;
    Ldlt8 F;                              // covariance = diag(H^-1)[0:6]
    F.factor(S.Hb[S.hsel], 1.0, 0xffu);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double e[8];
#pragma unroll
        for (int i = 0; i < 8; i++) e[i] = (i == k) ? 1.0 : 0.0;
        F.solve(e);
        o.cov[k] = e[k];
    }
}

template <int TRK_THREADS>
__global__ void __launch_bounds__(TRK_THREADS, 1) track_kernel(const TrackParams P, const Candidate *__restrict__ cands, TrackOut *__restrict__ outs) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int) cluster.num_blocks(), rank = (int) cluster.block_rank();
    const int cand = blockIdx.x / CL;
    constexpr int TRK_WARPS = TRK_THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ TrkState S;
    __shared__ float s_warp[TRK_WARPS][TRK_SUMS];
    __shared__ double s_part[2][TRK_MAX_CLUSTER][TRK_SUMS];
    __shared__ double s_sum[TRK_SUMS];
    __shared__ int s_n[OPTL];          // point counts of the levels (read once: a global load per evaluation is a full L2 round trip)

    if (tid == 0) {
        const Candidate &c = cands[cand];
        for (int i = 0; i < 9; i++) S.cur.R[i] = c.R[i];
        for (int i = 0; i < 3; i++) S.cur.t[i] = c.t[i];
        S.a = c.a; S.b = c.b; S.an = c.a; S.bn = c.b; S.cand = S.cur;
        S.phase = 0; S.have_repeated = 0; S.iterations = 0; S.fail = 0; S.lambda = 0.01; S.it = 0;
        for (int l = 0; l < OPTL; l++) { S.rep[l] = 0.0; S.oE[l] = S.nE[l] = 0.0; S.oT[l] = S.oS[l] = S.oR[l] = S.nT[l] = S.nS[l] = S.nR[l] = 0; }
        for (int k = 0; k < 3; k++) S.oFlow[k] = S.nFlow[k] = 0.0;
        S.hsel = 0;
        for (int i = 0; i < 72; i++) S.Hb[0][i] = S.Hb[1][i] = (i < 64 && i % 9 == 0) ? 1.0 : 0.0;
        for (int i = 0; i < 8; i++) S.inc[i] = 0.0;
    }
    if (tid < TRK_SUMS) s_sum[tid] = 0.0;
    if (tid >= 64 && tid < 64 + OPTL) s_n[tid - 64] = *P.lv[tid - 64].pc_n;
    __syncthreads();
    cluster.sync();      // every CTA of the cluster is resident before anyone writes into its shared memory

    int buf = 0, evals = 0;
    long long cyc_adv = 0, cyc_eval = 0, cyc_red = 0, t_mark = clock64();
    for (;;) {
        if (tid == 0) advance(S, P, s_sum);
        __syncthreads();
        if (tid == 0) { const long long t = clock64(); cyc_adv += t - t_mark; t_mark = t; }
        if (S.cmd.exit) break;
        evals++;

        // ---- evaluate the requested pose on this CTA's share of the level's points (computeResidual + computeHessian in one pass)
        const EvalCmd &c = S.cmd;
        const LevelDev &L = P.lv[c.level];
        const int n = s_n[c.level];
        const float wl3 = (float) (L.w - 3), hl3 = (float) (L.h - 3);
        const float huber = P.huber, base_cut = P.cutoff;
        float acc[TRK_SUMS];
#pragma unroll
        for (int k = 0; k < TRK_SUMS; k++) acc[k] = 0.f;
        // two-stage software pipeline: while point i is reduced into the sums, the taps of point i + stride and the record of
        // point i + 2 stride are in flight (the per-thread chain record -> projection -> taps is ~2 L2 latencies long)
        struct Tap { float4 t00, t10, t01, t11; float u, v, id, dx, dy, col; bool ok; };
        auto stage_a = [&](const float4 p, const int i, const bool in_range, Tap &o) {
            const float x = p.x, y = p.y, id = p.z, refColor = p.w;
            o.ok = false;
            if (!in_range || !isfinite(refColor)) return;
            const float rx = c.RKi[0] * x + c.RKi[1] * y + c.RKi[2], ry = c.RKi[3] * x + c.RKi[4] * y + c.RKi[5], rz = c.RKi[6] * x + c.RKi[7] * y + c.RKi[8];
            const float ptx = rx + c.t[0] * id, pty = ry + c.t[1] * id, ptz = rz + c.t[2] * id;
            const float u = ptx / ptz, v = pty / ptz;
            const float Ku = L.fx * u + L.cx, Kv = L.fy * v + L.cy;
            const float new_id = id / ptz;
            if (c.level == 0 && (i & 31) == 0) {      // flow indicators (DSOTracker.cpp:315-344)
                const float kx = c.Ki[0] * x + c.Ki[1] * y + c.Ki[2], ky = c.Ki[3] * x + c.Ki[4] * y + c.Ki[5], kz = c.Ki[6] * x + c.Ki[7] * y + c.Ki[8];
                const float ax = kx + c.t[0] * id, ay = ky + c.t[1] * id, az = kz + c.t[2] * id;
                const float bx = kx - c.t[0] * id, by = ky - c.t[1] * id, bz = kz - c.t[2] * id;
                const float cx3 = rx - c.t[0] * id, cy3 = ry - c.t[1] * id, cz3 = rz - c.t[2] * id;
                const float KuT = L.fx * (ax / az) + L.cx, KvT = L.fy * (ay / az) + L.cy;
                const float KuT2 = L.fx * (bx / bz) + L.cx, KvT2 = L.fy * (by / bz) + L.cy;
                const float Ku3 = L.fx * (cx3 / cz3) + L.cx, Kv3 = L.fy * (cy3 / cz3) + L.cy;
                acc[S_FT] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
                acc[S_FT] += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
                acc[S_FRT] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
                acc[S_FRT] += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
                acc[S_FNUM] += 2.f;
            }
            if (!(Ku > 2.f && Kv > 2.f && Ku < wl3 && Kv < hl3 && new_id > 0.f)) return;
            const int ix = (int) Ku, iy = (int) Kv;
            const float4 *g = L.grad + (size_t) iy * L.w + ix;
            o.t00 = g[0]; o.t10 = g[1]; o.t01 = g[L.w]; o.t11 = g[L.w + 1];
            o.u = u; o.v = v; o.id = new_id; o.dx = Ku - (float) ix; o.dy = Kv - (float) iy; o.col = refColor;
            o.ok = true;
        };
        auto stage_b = [&](const Tap &q) {
            if (!q.ok) return;
            const float dxdy = q.dx * q.dy;
            const float w00 = 1.f - q.dx - q.dy + dxdy, w10 = q.dx - dxdy, w01 = q.dy - dxdy;
            const float hI = dxdy * q.t11.x + w01 * q.t01.x + w10 * q.t10.x + w00 * q.t00.x;
            const float hx = dxdy * q.t11.y + w01 * q.t01.y + w10 * q.t10.y + w00 * q.t00.y;
            const float hy = dxdy * q.t11.z + w01 * q.t01.z + w10 * q.t10.z + w00 * q.t00.z;
            if (!(isfinite(hI) && isfinite(hx) && isfinite(hy))) return;
            const float r = hI - (c.aLL * q.col + c.bLL);
            const float ar = fabsf(r);
            const float hw = ar < huber ? 1.f : huber / ar;
            acc[S_NT] += 1.f;
            if (ar <= base_cut) acc[S_NROB] += 1.f;
            if (ar > c.cut) {
                acc[S_E] += c.maxE;
                acc[S_NSAT] += 1.f;
                return;
            }
            acc[S_E] += hw * r * r * (2.f - hw);
            const float gx = hx * L.fx, gy = hy * L.fy, u = q.u, v = q.v;
            float J[9];
            J[0] = q.id * gx;
            J[1] = q.id * gy;
            J[2] = 0.f - (q.id * (u * gx + v * gy));
            J[3] = 0.f - ((u * v * gx) + gy * (1.f + v * v));
            J[4] = (u * v * gy) + (gx * (1.f + u * u));
            J[5] = u * gy - v * gx;
            J[6] = c.aLL * (c.b0 - q.col);
            J[7] = -1.f;
            J[8] = r;
#pragma unroll
            for (int a = 0; a < 8; a++) {
                const float jw = J[a] * hw;
#pragma unroll
                for (int b = a; b < 9; b++) acc[h_index(a, b)] += jw * J[b];
            }
        };
        const int i0 = rank * TRK_THREADS + tid, stride = CL * TRK_THREADS;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        Tap cur;
        stage_a(i0 < n ? L.pc[i0] : zero4, i0, i0 < n, cur);
        float4 p_next = (i0 + stride < n) ? L.pc[i0 + stride] : zero4;
        for (int i = i0; i < n; i += stride) {
            const int j = i + stride;
            const float4 p_next2 = (j + stride < n) ? L.pc[j + stride] : zero4;
            Tap nxt;
            stage_a(p_next, j, j < n, nxt);
            stage_b(cur);
            cur = nxt; p_next = p_next2;
        }
        if (tid == 0) { const long long t = clock64(); cyc_eval += t - t_mark; t_mark = t; }
        // ---- CTA partial: transposing butterfly inside each warp, fp64 across warps (fixed order)
        s_warp[warp][lane] = transpose_sum(acc, lane);
        s_warp[warp][32 + lane] = transpose_sum(acc + 32, lane);
        __syncthreads();
        if (tid < TRK_SUMS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < TRK_WARPS; w++) s += (double) s_warp[w][tid];
            for (int r = 0; r < CL; r++) *cluster.map_shared_rank(&s_part[buf][rank][tid], r) = s;     // push into every CTA of the cluster
        }
        cluster.sync();
        if (tid < TRK_SUMS) {
            double s = 0.0;
            for (int r = 0; r < CL; r++) s += s_part[buf][r][tid];
            s_sum[tid] = s;
        }
        buf ^= 1;
        __syncthreads();
        if (tid < 72) S.Hb[S.hsel ^ 1][tid] = hessian_entry(tid, P, s_sum);
        __syncthreads();
        if (tid == 0) { const long long t = clock64(); cyc_red += t - t_mark; t_mark = t; }
    }
    if (rank == 0 && tid == 0) {
        write_out(S, P, outs[cand]);
        outs[cand].evals = evals; outs[cand].cyc_advance = cyc_adv; outs[cand].cyc_eval = cyc_eval; outs[cand].cyc_reduce = cyc_red;
    }
}

}  // namespace cmltrk
