// linearize.cuh -- the sampling kernel of the photometric-BA hot path and the device-side tile binning that feeds it.
//
// Reference behaviour restated on the device (file:line under /root/reference/src/cml):
//   linearize_tile_kernel   optimization/dso/DSOBundleAdjustment.cpp:62-316 (linearize) + :2051-2093 (applyRes, as a double-buffered
//                           candidate) + :1568-1599 (fixLinearization bookkeeping); image/Array2D.h:265-286 (bilinear taps).  The sums of
//                           addToHessianTop (:1648-1779, MatrixAccumulators.h:776-937) are taken from its Jacobian records by accumulate_kernel.
//
// Design (B200): the 8-pixel pattern of a residual reads a <= 6x6 texel footprint of the TARGET image.  Residuals are sorted on the
// device, once per run(), by (target frame, 64x32 tile that holds the centre projection); inside a tile they are interleaved by the
// shared-memory bank group of their centre texel.  A persistent CTA per SM (12 warps, no producer warp) walks a contiguous range of that
// order.  The 70x38-texel boxes (tile + halo 3) of the tiles the CTA needs are streamed through a 4-stage shared-memory ring with TMA
// (cp.async.bulk.tensor.2d + mbarrier complete_tx): two boxes up front, the next ones issued by the warp that first sees a box landed /
// whose arrival frees a stage.  One residual per lane, 32 consecutive residuals per pass; every bilinear tap comes from the staged tile.  A
// lane whose footprint left its box (pose drift since the binning, strong warps) or whose tile is not in the ring reads its taps from
// global memory through the same generic pointer, so correctness never depends on the binning; sparse windows (few residuals per tile)
// run without staging altogether (DevWin::tma_on = 0).
//
// Arithmetic: centre projection in fp64 (the reference's scalar_t); the 7 other pattern pixels as fp32 offsets from it,
//   q_i - q_c = f (d_xy - (P_xy/P_z) d_z) / (P_z + d_z),  d = sx A + sy B,  A = R[:,0]/fx, B = R[:,1]/fy,
// added to the centre kept as a float pair (hi, lo): |error| < 1.5e-6 px before the final rounding to float, i.e. the sample
// position differs from the reference's by at most one float ulp in a few percent of the pixels (measured).  A residual with
// a pixel closer than 1e-3 px to the in-bounds limits is re-projected exactly in fp64, so the OOB decision is the reference's.
//
// Output: the Jacobian record of every good residual (x[10] y[10] JIdx2[3] ...; the reference's efsJ, DSOResidual.h:22-69) goes to
// rj[candidate] in the HOST's bin-major residual order; addToHessianTop (the 13x13 blocks) is evaluated from these records by
// accumulate_kernel (kernels.cuh), bin by bin, in a fixed order: results are bitwise reproducible.
#pragma once
#include <cuda.h>

#include "dev_types.h"

namespace cmlba {

constexpr int LT_TILE_TABLE = 24;     // tile descriptors / user counts of a CTA that travel in its cta_info record and are kept in shared memory (more tiles: read from global memory)
constexpr int LT_INIT_BOXES = 2;      // boxes a CTA issues up front; the rest of the ring is filled as those land (bulk traffic in flight delays every other request of the SM)
constexpr int LT_INFO_INTS = 64;      // ints per cta_info record: q0 q1 t_first - | jd[4] | users[4] | - [4] | jd table [LT_TILE_TABLE] | users table [LT_TILE_TABLE]
struct alignas(64) TileMaps { CUtensorMap m[MAXF]; };   // one 2-D tensor map per window frame: rows of W float4 texels seen as 2W 8-byte elements

constexpr int LT_STAGE_REC = 32;       // Jacobian records a warp stages in shared memory at a time (two rounds per pass) for its line-coalesced stores
// CW warps, ST ring stages
__host__ __device__ __forceinline__ size_t lt_smem_bytes(int N, int CW, int ST) {
    return (size_t) ST * LT_STAGE_STRIDE + (size_t) 2 * N * sizeof(PairPre) + 2 * ST * sizeof(unsigned long long) + MAXF * sizeof(float) + 8 * sizeof(int) + LT_TILE_TABLE * 2 * sizeof(int) + (size_t) CW * LT_STAGE_REC * RJ_STRIDE * sizeof(float);
}

// ---- mbarrier / TMA (PTX ISA 8.x; SASS: SYNCS.*, UTMALDG)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_n(unsigned long long *bar, uint32_t n) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(n) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
constexpr int LT_EMPTY_COUNT = 1 << 12;   // arrival count of a stage's "empty" barrier: the producer tops the users of a tile up to this number
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, uint32_t parity) {      // non-blocking
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a tile that never lands (bad tensor map) degrades the warp to global-memory taps instead of hanging the device
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, uint32_t parity) {
#pragma unroll 1
    for (int i = 0; i < (1 << 22); i++) if (mbar_try_wait(bar, parity)) return true;   // every failed try_wait has already slept for the hardware's time limit
    return false;
}
// A consumer may run several tiles ahead of the producer, and a parity wait can only tell the current phase of a barrier from the
// previous one.  So the producer tags every stage with the tile it is loading (plain shared-memory store before the TMA is issued); a
// consumer first sees its tile's tag, then waits for the phase.  Its own pending arrival keeps the stage from moving on meanwhile.
__device__ __forceinline__ bool lt_wait_tile(const int *tag_, unsigned long long *full, const int s, const int q, const uint32_t parity) {
    const volatile int *tag = tag_;
    if (tag[s] != q) {
#pragma unroll 1
        for (int i = 0; tag[s] != q; i++) { if (i > (1 << 22)) return false; __nanosleep(64); }
    }
    return mbar_wait(full + s, parity);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x),
                 "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y) {     // the box goes to L2 only (SASS: UTMAPF)
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ float rcp_nr(float d) {      // MUFU.RCP + one Newton step: relative error ~1e-7 after rounding
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(r) : "f"(d));
    const float e = fmaf(-d, r, 1.f);
    return fmaf(r, e, r);
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_excl_scan_256(int v, int *total) {
    __shared__ int s_w[8];
    __shared__ int s_tot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (threadIdx.x == 0) { int run = 0; for (int k = 0; k < 8; k++) { const int t = s_w[k]; s_w[k] = run; run += t; } s_tot = run; }
    __syncthreads();
    const int res = inc - v + s_w[wid];
    *total = s_tot;
    __syncthreads();
    return res;
}

// ------------------------------------------------------------------------------------------------
// Tile binning, step 1: key of every residual (host order) = ((target * n_tiles + tile) * N + host), tile from the centre projection
// at the current state; histogram with integer atomics (order-independent); the last CTA scans the histogram into bin_offs,
// numbers the non-empty (target, tile) jobs and re-zeroes the histogram.
__global__ void __launch_bounds__(256) bin_count_kernel(const DevWin w) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < w.R) {
        const int p = w.r_point[i], h = w.r_host[i], t = w.r_target[i];
        const PairPre &pp = w.pairs[h * w.N + t];
        const double rho = w.pt_idepth[p];
        const double kx = ((double) w.pt_x[p] - w.cx) * w.fxi, ky = ((double) w.pt_y[p] - w.cy) * w.fyi;
        const double P0 = pp.R[0] * kx + pp.R[1] * ky + pp.R[2] + pp.t[0] * rho;
        const double P1 = pp.R[3] * kx + pp.R[4] * ky + pp.R[5] + pp.t[1] * rho;
        const double P2 = pp.R[6] * kx + pp.R[7] * ky + pp.R[8] + pp.t[2] * rho;
        const double iz = 1.0 / P2;
        double Ku = P0 * iz * w.fx + w.cx, Kv = P1 * iz * w.fy + w.cy;
        int tx = 0, ty = 0;
        if (Ku == Ku && Kv == Kv) {          // not NaN; infinities clamp
            Ku = fmin(fmax(Ku, 0.0), (double) (w.W - 1)); Kv = fmin(fmax(Kv, 0.0), (double) (w.H - 1));
            tx = (int) Ku / LT_TILE_W; ty = (int) Kv / LT_TILE_H;
        }
        const int key = ((t * w.n_tiles) + ty * w.tiles_x + tx) * w.N + h;
        w.bin_key[i] = key;
        // shared-memory bank group (16-byte units, 8 per 128-byte row of banks) of the centre texel inside the tile's staged box
        w.bin_g[i] = (uint8_t) ((((int) Kv - (ty * LT_TILE_H - LT_HALO)) * LT_BOX_W + ((int) Ku - (tx * LT_TILE_W - LT_HALO))) & 7);
        atomicAdd(w.bin_hist + key, 1);
    }
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(w.bin_ticket, 1) == (int) gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) w.bin_ticket[0] = 0;
    // groups g = (target, tile) of N keys (one per host): exclusive scan of the counts (bin_offs) and numbering of the non-empty groups (tile jobs)
    const int N = w.N, G = N * w.n_tiles;
    int carry_c = 0, carry_j = 0;
    for (int g0 = 0; g0 < G; g0 += 256) {
        const int g = g0 + threadIdx.x;
        int vals[MAXF];
        int c = 0;
#pragma unroll
        for (int h = 0; h < MAXF; h++) { vals[h] = (g < G && h < N) ? __ldcg(w.bin_hist + (size_t) g * N + h) : 0; c += vals[h]; }
        int tc, tj;
        const int ec = block_excl_scan_256(c, &tc), ej = block_excl_scan_256(c > 0 ? 1 : 0, &tj);
        if (g < G) {
            int run = carry_c + ec;
#pragma unroll
            for (int h = 0; h < MAXF; h++) if (h < N) { w.bin_offs[(size_t) g * N + h] = run; run += vals[h]; }
            if (c > 0) {
                const int job = carry_j + ej, t = g / w.n_tiles, tile = g - t * w.n_tiles;
                w.job_desc[job] = (uint32_t) t | ((uint32_t) (tile % w.tiles_x) << 4) | ((uint32_t) (tile / w.tiles_x) << 16);
                w.job_of_tile[g] = job;
                w.job_begin[job] = carry_c + ec;             // first sorted residual of the job
            } else w.job_of_tile[g] = -1;
        }
        carry_c += tc; carry_j += tj;
    }
    if (threadIdx.x == 0) w.job_begin[carry_j] = carry_c;    // = R
}

// step 2: stable scatter into the sorted order.  One CTA per (target, host) bin: its residuals are contiguous in host order and it owns
// every key of the bin.  Warp k takes the k-th eighth of the bin; per-warp tile counts (pass 1) are prefixed over the warps on top of
// bin_offs, then every warp ranks its residuals inside their (target, tile, host) group by host order (pass 2) -- deterministic,
// no atomics.
__global__ void __launch_bounds__(256) bin_scatter_kernel(const DevWin w) {
    extern __shared__ int s_run[];           // [8 warps][n_tiles]
    const int bin = blockIdx.x, N = w.N, t = bin / N, h = bin - t * N, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nt = w.n_tiles;
    const int b0 = w.res_bin_begin[bin], b1 = w.res_bin_begin[bin + 1];
    if (b0 >= b1) return;
    const int per = (((b1 - b0 + 7) >> 3) + 31) & ~31;
    const int a0 = min(b0 + wid * per, b1), a1 = min(a0 + per, b1);
    for (int k = threadIdx.x; k < 8 * nt; k += 256) s_run[k] = 0;
    __syncthreads();
    int *mine = s_run + wid * nt;
    for (int base = a0; base < a1; base += 32) {           // pass 1: how many residuals of my slice fall into every tile
        const int i = base + lane;
        const bool valid = i < a1;
        const int tile = valid ? (w.bin_key[i] / N) % nt : -1 - lane;      // invalid lanes match nobody
        const unsigned m = __match_any_sync(0xffffffffu, tile);
        if (valid && (m & ((1u << lane) - 1u)) == 0) mine[tile] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
    for (int tile = threadIdx.x; tile < nt; tile += 256) {
        int run = w.bin_offs[((size_t) t * nt + tile) * N + h];
        for (int k = 0; k < 8; k++) { const int c = s_run[k * nt + tile]; s_run[k * nt + tile] = run; run += c; }
    }
    __syncthreads();
    for (int base = a0; base < a1; base += 32) {           // pass 2: scatter
        const int i = base + lane;
        const bool valid = i < a1;
        const int tile = valid ? (w.bin_key[i] / N) % nt : -1 - lane;
        const int rp = valid ? w.r_point[i] : 0;
        const unsigned m = __match_any_sync(0xffffffffu, tile);
        const int rank = __popc(m & ((1u << lane) - 1u));
        int pos = 0;
        if (valid) pos = mine[tile] + rank;
        __syncwarp();
        if (valid && rank == 0) mine[tile] += __popc(m);
        __syncwarp();
        if (valid) {
            w.r_pht[pos] = (uint32_t) rp | ((uint32_t) h << 24) | ((uint32_t) t << 28);
            w.r_job[pos] = w.job_of_tile[t * nt + tile];
            w.r_src[pos] = i;
        }
    }
}

// step 2a: inside every tile job the residuals are re-ordered so that consecutive lanes cycle through the eight bank groups of their centre
// texel.  All lanes of a warp apply (nearly) the same pattern offsets, so the 32 taps of eight consecutive lanes then fall into eight
// different 16-byte bank groups: the LDS.128 of a quarter warp is conflict-free instead of ~2.6-way conflicted (random positions).
// Order inside a group = sorted order (deterministic).  One CTA per (target, tile); jobs larger than LT_IL_MAX keep their order.
constexpr int LT_IL_MAX = 1024;
__global__ void __launch_bounds__(256) bin_interleave_kernel(const DevWin w) {
    const int job = w.job_of_tile[blockIdx.x];
    if (job < 0) return;
    const int jb = w.job_begin[job], n = w.job_begin[job + 1] - jb;
    if (n < 16 || n > LT_IL_MAX) return;
    __shared__ uint32_t s_pht[LT_IL_MAX];
    __shared__ int s_src[LT_IL_MAX];
    __shared__ short s_rank[LT_IL_MAX];
    __shared__ uint8_t s_g[LT_IL_MAX];
    __shared__ int s_cnt[8];
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += 256) { const int src = w.r_src[jb + i]; s_pht[i] = w.r_pht[jb + i]; s_src[i] = src; s_g[i] = w.bin_g[src]; }
    __syncthreads();
    const int i0 = tid * 4;                          // blocked assignment keeps the sorted order inside a group
    for (int g = 0; g < 8; g++) {
        int c = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) c += (i0 + k < n && s_g[i0 + k] == g) ? 1 : 0;
        int tot;
        int run = block_excl_scan_256(c, &tot);
#pragma unroll
        for (int k = 0; k < 4; k++) if (i0 + k < n && s_g[i0 + k] == g) s_rank[i0 + k] = (short) run++;
        if (tid == 0) s_cnt[g] = tot;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int g = s_g[i], j = s_rank[i];
        int pos = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { const int cq = s_cnt[q]; pos += min(cq, j) + ((q < g && cq > j) ? 1 : 0); }
        w.r_pht[jb + pos] = s_pht[i]; w.r_src[jb + pos] = s_src[i];
    }
}

// one box into a ring stage, as a real call: the call sites sit in register-critical code and run once per tile
__device__ __noinline__ void lt_issue_box(unsigned char *stage, const CUtensorMap *maps, const uint32_t jd, const int users, unsigned long long *full_s, unsigned long long *empty_s) {
    const int t = (int) (jd & 15u), tx = (int) ((jd >> 4) & 0xfffu), ty = (int) (jd >> 16);
    mbar_arrive_n(empty_s, (uint32_t) (LT_EMPTY_COUNT - users));
    mbar_expect_tx(full_s, LT_TILE_BYTES);
    tma_load_2d(stage, maps + t, (tx * LT_TILE_W - LT_HALO) * 2, ty * LT_TILE_H - LT_HALO, full_s);
}

// step 2b: the per-point constants of every residual (pixel, reference colours, gradient weights), copied into the sorted residual order as
// five float4 columns r_pt4[k][R]: the sampling kernel reads them with fully coalesced 16-byte loads (20 LSU wavefronts per warp pass)
// instead of gathering seven arrays by point index (~250 wavefronts: the load/store unit, not HBM, was what its first phase waited on)
__global__ void __launch_bounds__(256) bin_pack_kernel(const DevWin w) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= w.R) return;
    const int p = (int) (w.r_pht[r] & 0xffffffu);
    const float4 *col = reinterpret_cast<const float4 *>(w.pt_colors + (size_t) p * 8), *wt = reinterpret_cast<const float4 *>(w.pt_weights + (size_t) p * 8);
    w.r_pt4[r] = make_float4(w.pt_x[p], w.pt_y[p], 0.f, 0.f);
    w.r_pt4[(size_t) w.R + r] = col[0]; w.r_pt4[(size_t) 2 * w.R + r] = col[1];
    w.r_pt4[(size_t) 3 * w.R + r] = wt[0]; w.r_pt4[(size_t) 4 * w.R + r] = wt[1];
}

// step 3: per-CTA records of linearize_tile_kernel (grid = lt_grid persistent CTAs over contiguous ranges of warp passes)
__global__ void __launch_bounds__(256) bin_finish_kernel(const DevWin w) {
    const int per_cta = (w.n_chunks + w.lt_grid - 1) / w.lt_grid;
    for (int b = threadIdx.x; b < w.lt_grid; b += 256) {
        const int c0 = b * per_cta, c1 = min(c0 + per_cta, w.n_chunks);
        int *info = w.cta_info + (size_t) b * LT_INFO_INTS;
        for (int k = 0; k < LT_INFO_INTS; k++) info[k] = 0;
        if (c0 >= c1) continue;
        const int q0 = w.r_job[c0 * 32], q1 = w.r_job[min(c1 * 32, w.R) - 1];
        info[0] = q0; info[1] = q1; info[2] = (int) (w.r_pht[c0 * 32] >> 28);
        for (int i = 0; i < LT_TILE_TABLE && q0 + i <= q1; i++) {
            const int jb = w.job_begin[q0 + i], je = w.job_begin[q0 + i + 1];
            const int jd = (int) w.job_desc[q0 + i], users = min((je - 1) >> 5, c1 - 1) - max(jb >> 5, c0) + 1;
            if (i < 4) { info[4 + i] = jd; info[8 + i] = users; }
            info[16 + i] = jd; info[16 + LT_TILE_TABLE + i] = users;
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <bool kDump, int LT_CWARPS, int LT_STAGES>
__global__ void __launch_bounds__(LT_CWARPS * 32, 1) linearize_tile_kernel(const DevWin w, const __grid_constant__ TileMaps tm, const int fix, const int respect_done) {
    pdl_enter();
    Ctrl *ctrl = w.ctrl;
    extern __shared__ __align__(1024) unsigned char lt_smem[];
    unsigned char *ring = lt_smem;                                                                  // [LT_STAGES][LT_BOX_H][LT_BOX_W] float4
    PairPre *s_pairs = reinterpret_cast<PairPre *>(lt_smem + (size_t) LT_STAGES * LT_STAGE_STRIDE); // [2 targets][N hosts]
    unsigned long long *full = reinterpret_cast<unsigned long long *>(s_pairs + 2 * w.N);
    unsigned long long *empty = full + LT_STAGES;
    float *s_th = reinterpret_cast<float *>(empty + LT_STAGES);                                     // [N] frameEnergyTH
    int *s_tag = reinterpret_cast<int *>(s_th + MAXF);                                              // [LT_STAGES] tile in (or on its way into) every stage
    unsigned long long *pairs_bar = reinterpret_cast<unsigned long long *>(s_tag + 4);              // the staged constants are in place
    int *s_jd = s_tag + 8, *s_users = s_jd + LT_TILE_TABLE;                                         // descriptors / user counts of the CTA's first LT_TILE_TABLE tiles
    float4 *s_stage = reinterpret_cast<float4 *>(s_users + LT_TILE_TABLE) + (size_t) (threadIdx.x >> 5) * (LT_STAGE_REC * RJ_STRIDE / 4);   // [LT_CWARPS][LT_STAGE_REC][9] float4
    const int N = w.N, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (w.n_chunks + (int) gridDim.x - 1) / (int) gridDim.x;
    const int c0 = blockIdx.x * per, c1 = min(c0 + per, w.n_chunks);
    if (c0 >= c1) return;
    // ---- ONE exposed memory latency before the first boxes can be issued and the first point records requested: everything below depends
    // on the launch geometry only -- the control block, the CTA's record (tile range, descriptors and user counts of its first
    // LT_TILE_TABLE tiles; written by bin_finish_kernel) and the per-residual scalars of every warp's first pass (both copies of the
    // double-buffered state / energy: which one is current is part of the same round trip).
    const int4 *infop = reinterpret_cast<const int4 *>(w.cta_info + (size_t) blockIdx.x * LT_INFO_INTS);
    const int done_ld = ctrl->done, cur = ctrl->cur, fin_ld = respect_done == 2 ? ctrl->final_done : 0;
    const int4 info = __ldg(infop);
    int tab_jd = 0, tab_users = 0;
    if ((int) threadIdx.x < LT_TILE_TABLE) { tab_jd = __ldg(w.cta_info + (size_t) blockIdx.x * LT_INFO_INTS + 16 + threadIdx.x); tab_users = __ldg(w.cta_info + (size_t) blockIdx.x * LT_INFO_INTS + 16 + LT_TILE_TABLE + threadIdx.x); }
    uint32_t pht = 0; int job = 0, src = 0; uint8_t alive_ld = 0, st = RES_OOB, nst = 0; float e_old = 0.f, ne = 0.f;
    {
        const int rr = (c0 + warp < c1 ? c0 + warp : c0) * 32 + lane, rl = rr < w.R ? rr : w.R - 1;
        pht = __ldg(w.r_pht + rl); job = __ldg(w.r_job + rl); src = __ldg(w.r_src + rl);
        alive_ld = w.r_alive[rl]; nst = w.r_new_state[rl]; ne = w.r_new_energy[rl];
        const uint8_t st0 = w.r_state[0][rl], st1 = w.r_state[1][rl];
        const float e0 = w.r_energy[0][rl], e1 = w.r_energy[1][rl];
        st = rr < w.R ? (cur ? st1 : st0) : (uint8_t) RES_OOB; e_old = cur ? e1 : e0;
    }
    const int nxt = cur ^ 1;
    if (respect_done && launch_skipped(respect_done, done_ld, fin_ld)) return;
    // the per-residual scalars of every later pass are requested at the bottom of the pass before it
    auto load_headers = [&](const int cc) {
        const int rr = cc * 32 + lane, rl = rr < w.R ? rr : w.R - 1;
        pht = __ldg(w.r_pht + rl); job = __ldg(w.r_job + rl); src = __ldg(w.r_src + rl);
        alive_ld = w.r_alive[rl]; st = rr < w.R ? w.r_state[cur][rl] : (uint8_t) RES_OOB; e_old = w.r_energy[cur][rl];
        nst = w.r_new_state[rl]; ne = w.r_new_energy[rl];
    };
    // development trace: stamp k of this warp (SM clock); compiled to a predicated-off store in normal runs
    long long *trace = (w.lt_mode & 2) ? w.lt_trace + ((size_t) blockIdx.x * 16 + warp) * 32 : nullptr;
#define LT_STAMPK(k) do { if (trace && (k) < 28 && lane == (__ffs(__activemask()) - 1)) trace[(k)] = clock64(); } while (0)
    LT_STAMPK(0);
    if (trace && lane == 0) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); trace[30] = gt; }
    const int r_first = c0 * 32, r_last = min(c1 * 32, w.R) - 1;
    const int q0 = info.x, q1 = info.y, t_first = info.z, n_tiles_cta = q1 - q0 + 1;
    if (trace && q0 >= 0) LT_STAMPK(22);             // the CTA record has arrived
    // There is no producer warp.  A stage goes back to "empty" when every warp pass that overlaps its tile (the passes are consecutive, so
    // their number follows from the tile's residual range) has arrived on the stage's barrier; the warp whose arrival completes that
    // phase issues the box of the tile that comes LT_STAGES later into the same stage (one winner: compare-and-swap on the stage tag),
    // topping the new tile's user count up to the barrier's fixed arrival count.
    auto issue = [&](const int i, const uint32_t jd_, const int users) {      // tile q0 + i, by one thread; the stage tag is already set
        const int s = i % LT_STAGES;
        const int t = (int) (jd_ & 15u), tx = (int) ((jd_ >> 4) & 0xfffu), ty = (int) (jd_ >> 16);
        mbar_arrive_n(empty + s, (uint32_t) (LT_EMPTY_COUNT - users));
        mbar_expect_tx(full + s, LT_TILE_BYTES);
        tma_load_2d(ring + (size_t) s * LT_STAGE_STRIDE, &tm.m[t], (tx * LT_TILE_W - LT_HALO) * 2, ty * LT_TILE_H - LT_HALO, full + s);
    };
    auto tile_meta = [&](const int i, uint32_t &jd, int &users) {
        if (i < LT_TILE_TABLE) { jd = (uint32_t) ((volatile int *) s_jd)[i]; users = ((volatile int *) s_users)[i]; return; }
        jd = __ldg(w.job_desc + q0 + i);
        const int jb = __ldg(w.job_begin + q0 + i), je = __ldg(w.job_begin + q0 + i + 1);
        users = min((je - 1) >> 5, c1 - 1) - max(jb >> 5, c0) + 1;
    };
    // lane 0 of a warp, after its pass is done with tile q0 + k: arrive; if that completed the stage's phase, bring in tile q0 + k + LT_STAGES
    auto release = [&](const int k) {
        const int s = k % LT_STAGES;
        mbar_arrive(empty + s);
        if (k + LT_STAGES < n_tiles_cta && mbar_test(empty + s, (uint32_t) ((k / LT_STAGES) & 1)) && atomicCAS(s_tag + s, q0 + k, q0 + k + LT_STAGES) == q0 + k) {
            uint32_t jd; int users;
            tile_meta(k + LT_STAGES, jd, users);
            issue(k + LT_STAGES, jd, users);
        }
    };
    // Only LT_INIT_BOXES boxes are issued up front: a box in flight sits in front of every other request of the SM (measured: with four
    // 42 KB boxes outstanding the first per-residual loads of the kernel took 3 us longer).  Tile k + LT_INIT_BOXES of the initial ring
    // window is issued by whoever first sees tile k landed (one winner per tile: compare-and-swap on s_pump).
    int *s_pump = s_tag + 6;                         // [2]
    auto pump = [&](const int k) {
        const int i = k + LT_INIT_BOXES;
        if (i < LT_STAGES && i < n_tiles_cta && atomicCAS(s_pump + (i - LT_INIT_BOXES), 0, 1) == 0)
            lt_issue_box(ring + (size_t) i * LT_STAGE_STRIDE, tm.m, (uint32_t) ((volatile int *) s_jd)[i], ((volatile int *) s_users)[i], full + i, empty + i);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, LT_EMPTY_COUNT); s_tag[s] = q0 + s; }
        s_pump[0] = 0; s_pump[1] = 0;
        mbar_init(pairs_bar, LT_CWARPS * 32);
        mbar_fence_init();
        if (w.tma_on) {
            const int4 jd4 = __ldg(infop + 1), us4 = __ldg(infop + 2);
            const int init = (w.lt_mode & 16) ? LT_STAGES : LT_INIT_BOXES;     // development: 16 = the whole ring up front
            if (0 < n_tiles_cta && 0 < init) issue(0, (uint32_t) jd4.x, us4.x);
            if (1 < n_tiles_cta && 1 < init) issue(1, (uint32_t) jd4.y, us4.y);
            if (2 < n_tiles_cta && 2 < init) { s_pump[0] = 1; issue(2, (uint32_t) jd4.z, us4.z); }
            if (3 < n_tiles_cta && 3 < init) { s_pump[1] = 1; issue(3, (uint32_t) jd4.w, us4.w); }
        }
    }
    if ((int) threadIdx.x < LT_TILE_TABLE) { s_jd[threadIdx.x] = tab_jd; s_users[threadIdx.x] = tab_users; }
    LT_STAMPK(23);                                   // (warp 0: barriers initialised, first boxes issued)
    __syncthreads();                                 // the barriers and the tile table are set up (the first boxes are on their way)
    // ---- second round trip: the pair constants (staged in shared memory by all threads) and, issued before the staging stores can
    // stall on them, the point records of every warp's first pass
    LT_STAMPK(24);
    constexpr int PW = sizeof(PairPre) / 8, PAIR_IT = (2 * MAXF * PW + LT_CWARPS * 32 - 1) / (LT_CWARPS * 32);
    double pair_v[PAIR_IT];
#pragma unroll
    for (int it = 0; it < PAIR_IT; it++) {
        const int i = threadIdx.x + it * LT_CWARPS * 32;
        const int rec = i / PW, k = i - rec * PW, tt = rec / N, h = rec - tt * N, t = min(t_first + tt, N - 1);
        pair_v[it] = i < 2 * N * PW ? reinterpret_cast<const double *>(w.pairs + h * N + t)[k] : 0.0;
    }
    const float th_v = (int) threadIdx.x < N ? w.frames[threadIdx.x].energy_th : 0.f;
    // the point record of a pass (requested right after its per-residual scalars have arrived)
    uint32_t jd = 0, pht_next = 0; double rho = 0.0; float xcf = 0.f, ycf = 0.f; float4 c0v, c1v, w0v, w1v;
    auto load_point = [&](const int cc) {
        const int pp_ = (int) (pht & 0xffffffu), rl = min(cc * 32 + lane, w.R - 1);
        jd = __ldg(w.job_desc + job);
        rho = w.pt_idepth[pp_];                      // the only per-point value that changes between passes: gathered
        const float4 xy = __ldg(w.r_pt4 + rl);
        xcf = xy.x; ycf = xy.y;
        c0v = __ldg(w.r_pt4 + (size_t) w.R + rl); c1v = __ldg(w.r_pt4 + (size_t) 2 * w.R + rl);
        w0v = __ldg(w.r_pt4 + (size_t) 3 * w.R + rl); w1v = __ldg(w.r_pt4 + (size_t) 4 * w.R + rl);
        pht_next = (cc + LT_CWARPS < c1) ? __ldg(w.r_pht + min((cc + LT_CWARPS) * 32 + lane, w.R - 1)) : pht;     // for the L2 prefetch of the pass after
    };
    if (c0 + warp < c1) {     // first pass: the point record is on its way to L2 while the pair constants are staged (no registers held across the prologue)
        const int pn = (int) (pht & 0xffffffu);
        prefetch_l2(w.pt_idepth + pn); prefetch_l2(w.job_desc + job);
        if (lane < 20) prefetch_l2(w.r_pt4 + (size_t) (lane >> 2) * w.R + (c0 + warp) * 32 + (lane & 3) * 8);      // 5 columns x 4 lines
    }
    {
        double *dst = reinterpret_cast<double *>(s_pairs);
#pragma unroll
        for (int it = 0; it < PAIR_IT; it++) { const int i = threadIdx.x + it * LT_CWARPS * 32; if (i < 2 * N * PW) dst[i] = pair_v[it]; }
        if ((int) threadIdx.x < N) s_th[threadIdx.x] = th_v;
        mbar_arrive(pairs_bar);                      // release: the stores above are visible to whoever sees the phase complete
    }
    LT_STAMPK(25);                                   // pair constants staged
    // ---- consumer warps.  The per-residual arrays of the CTA's range are contiguous: one L2 prefetch per 128-byte line up front (the point
    // records of a pass are prefetched during the pass before it, below), so that only a warp's first pass waits on HBM for its inputs.
    {
        const int first_line = (r_first >> 5) + LT_CWARPS, n_lines = (r_last >> 5) - first_line + 1;      // the first pass of every warp has its own loads in flight
        for (int k = threadIdx.x; k < 7 * n_lines; k += LT_CWARPS * 32) {
            const int arr = k / n_lines, r = (first_line + (k - arr * n_lines)) << 5;
            const void *ptr = arr == 0 ? (const void *) (w.r_pht + r) : arr == 1 ? (const void *) (w.r_job + r) : arr == 2 ? (const void *) (w.r_energy[cur] + r)
                            : arr == 3 ? (const void *) (w.r_new_energy + r) : arr == 4 ? (const void *) (w.r_src + r)
                            : arr == 5 ? (const void *) (w.r_state[cur] + r) : (const void *) (w.r_new_state + r);
            prefetch_l2(ptr);
        }
    }
    LT_STAMPK(26);                                   // prefetches issued
    bool pairs_ready = false;
    // ---- consumer warps
    bool tma_ok = w.tma_on != 0;
    const double Wm2 = (double) ((float) w.W - 2.f), Hm2 = (double) ((float) w.H - 2.f);
    for (int c = c0 + warp; c < c1; c += LT_CWARPS, c < c1 ? load_headers(c) : (void) 0) {
        const int r = c * 32 + lane;
        const bool in_chunk = r < w.R;
        load_point(c);
        if (c + LT_CWARPS < c1 && lane < 8) {        // the per-residual scalars of this warp's NEXT pass (one line per array): into L1 now, loaded at the bottom of this pass
            const int rn = (c + LT_CWARPS) * 32;
            const void *ptr = lane == 0 ? (const void *) (w.r_pht + rn) : lane == 1 ? (const void *) (w.r_job + rn) : lane == 2 ? (const void *) (w.r_src + rn)
                            : lane == 3 ? (const void *) (w.r_alive + rn) : lane == 4 ? (const void *) (w.r_state[cur] + rn) : lane == 5 ? (const void *) (w.r_energy[cur] + rn)
                            : lane == 6 ? (const void *) (w.r_new_state + rn) : (const void *) (w.r_new_energy + rn);
            prefetch_l1(ptr);
        }
        const int p = (int) (pht & 0xffffffu), h = (int) ((pht >> 24) & 15u), t = (int) (pht >> 28);
        const bool valid = in_chunk && alive_ld;
        const double xc = (double) xcf, yc = (double) ycf;
        const int tr_b = 1 + 7 * ((c - c0 - warp) / LT_CWARPS);
        LT_STAMPK(tr_b);                             // pass start
        if (!pairs_ready) { mbar_wait(pairs_bar, 0); pairs_ready = true; }
        const PairPre *ppp = (t - t_first < 2) ? s_pairs + (t - t_first) * N + h : w.pairs + h * N + t;
        const PairPre &pp = *ppp;
        double ret = 0.0;
        uint8_t st_out = st;
        float e_out = e_old, neo = -1.f;             // state_NewEnergyWithOutlier = -1 (BA:66)
        bool good = false, sample = false;
        float qx[8], qy[8];
        double Pc0 = 0, Pc1 = 0, Pc2 = 1, Kuc = 0, Kvc = 0;
        if (valid) {
            ret = (double) e_old;                    // every early exit returns state_energy
            if (st != RES_OOB) {
                const double R0 = pp.R[0], R1 = pp.R[1], R2 = pp.R[2], R3 = pp.R[3], R4 = pp.R[4], R5 = pp.R[5], R6 = pp.R[6], R7 = pp.R[7], R8 = pp.R[8];
                const double tx = pp.t[0] * rho, ty = pp.t[1] * rho, tz = pp.t[2] * rho;
                {   // centre in fp64 (BA:107-118)
                    const double kx = (xc - w.cx) * w.fxi, ky = (yc - w.cy) * w.fyi;
                    Pc0 = R0 * kx + R1 * ky + R2 + tx;
                    Pc1 = R3 * kx + R4 * ky + R5 + ty;
                    Pc2 = R6 * kx + R7 * ky + R8 + tz;
                }
                const double iP2 = 1.0 / Pc2;        // one reciprocal instead of two divisions (<= 1 ulp of fp64 before the cast to float)
                const double un = Pc0 * iP2, vn = Pc1 * iP2;
                Kuc = un * w.fx + w.cx; Kvc = vn * w.fy + w.cy;
                const bool cin = Kuc >= 2.0 && Kvc >= 2.0 && Kuc < Wm2 && Kvc < Hm2;
                if (cin) {                           // setCenterProjectedTo (BA:131)
                    w.r_center[(size_t) r * 3 + 0] = (float) Kuc; w.r_center[(size_t) r * 3 + 1] = (float) Kvc; w.r_center[(size_t) r * 3 + 2] = (float) ((double) (float) iP2 * rho);
                }
                // pattern pixels as fp32 offsets from the centre
                const float kh_x = (float) Kuc, kh_y = (float) Kvc;
                const float kl_x = (float) (Kuc - (double) kh_x), kl_y = (float) (Kvc - (double) kh_y);
                const float uf = (float) un, vf = (float) vn, pz = (float) Pc2, fxf = (float) w.fx, fyf = (float) w.fy;
                const float a0 = pp.Af[0], a1 = pp.Af[1], a2 = pp.Af[2], b0_ = pp.Bf[0], b1_ = pp.Bf[1], b2_ = pp.Bf[2];
                bool all_in = cin, any_out = !cin;
                const float lo_in = 2.f + 1e-3f, hx_in = (float) Wm2 - 1e-3f, hy_in = (float) Hm2 - 1e-3f, lo_out = 2.f - 1e-3f, hx_out = (float) Wm2 + 1e-3f, hy_out = (float) Hm2 + 1e-3f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float sx = (float) pat_sx(i), sy = (float) pat_sy(i);     // compile-time constants after unrolling
                    const float d0 = sx * a0 + sy * b0_, d1 = sx * a1 + sy * b1_, d2 = sx * a2 + sy * b2_;
                    const float rc = rcp_nr(pz + d2);
                    const float dx = (fmaf(-uf, d2, d0) * rc) * fxf, dy = (fmaf(-vf, d2, d1) * rc) * fyf;
                    const float x = i == 4 ? kh_x : kh_x + (kl_x + dx), y = i == 4 ? kh_y : kh_y + (kl_y + dy);
                    qx[i] = x; qy[i] = y;
                    all_in = all_in && (x >= lo_in && y >= lo_in && x < hx_in && y < hy_in);
                    any_out = any_out || (x < lo_out || y < lo_out || x >= hx_out || y >= hy_out);
                }
                bool inb = all_in;
                if ((!all_in && !any_out) || w.lt_exact) {
                    // a pixel within 1e-3 px of the limits (or a non-finite offset): the reference's own fp64 projection decides (BA:197-212)
                    inb = true;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const double kx = (xc + (double) c_sx[i] - w.cx) * w.fxi, ky = (yc + (double) c_sy[i] - w.cy) * w.fyi;
                        const double P0 = R0 * kx + R1 * ky + R2 + tx, P1 = R3 * kx + R4 * ky + R5 + ty, P2 = R6 * kx + R7 * ky + R8 + tz;
                        const double iz = 1.0 / P2;
                        const double Ku = (P0 * iz) * w.fx + w.cx, Kv = (P1 * iz) * w.fy + w.cy;
                        inb = inb && (Ku >= 2.0 && Kv >= 2.0 && Ku < Wm2 && Kv < Hm2);
                        qx[i] = (float) Ku; qy[i] = (float) Kv;
                    }
                }
                if (!inb) nst = RES_OOB;             // setNewState(OOB) (BA:116, 210); state_NewEnergy keeps its old value
                else sample = true;
            }
        }
        LT_STAMPK(tr_b + 1);                         // projection done
        // ---- the tiles of this pass: [q_lo, q_last]; the first LT_STAGES of them can be in the ring together
        const int q_lo = __shfl_sync(0xffffffffu, job, 0), q_last = __shfl_sync(0xffffffffu, job, 31);
        const int q_hi = min(q_last, q_lo + LT_STAGES - 1);
        if (tma_ok) {
            bool ok = true;
            for (int q = q_lo; q <= q_hi; q++) {
                const int k = q - q0;
                ok = ok && lt_wait_tile(s_tag, full, k % LT_STAGES, q, (uint32_t) ((k / LT_STAGES) & 1));
                if (ok && k < LT_INIT_BOXES && lane == 0) pump(k);
            }
            tma_ok = __all_sync(0xffffffffu, ok);
        }
        LT_STAMPK(tr_b + 2);                         // tiles landed
        if ((w.lt_mode & 1) == 1) {      // development: ring protocol only
            __syncwarp();
            if (tma_ok && lane == 0) for (int q = q_lo; q <= q_hi; q++) release(q - q0);
            if (tma_ok) for (int q = q_hi + 1; q <= q_last; q++) { const int k = q - q0; if (!lt_wait_tile(s_tag, full, k % LT_STAGES, q, (uint32_t) ((k / LT_STAGES) & 1))) break; __syncwarp(); if (lane == 0) release(k); }
            continue;
        }
        // where this lane's taps come from: its staged tile, or the image itself
        bool use_smem = tma_ok && job <= q_hi;
        const int box_x = (int) ((jd >> 4) & 0xfffu) * LT_TILE_W - LT_HALO, box_y = (int) (jd >> 16) * LT_TILE_H - LT_HALO;
        // ---- taps: from the staged tile when the whole footprint is inside its box, else from the image (image/Array2D.h:265-286)
        int ox = 0, oy = 0, pitch = w.W;
        const float4 *tbase = w.img[t];
        if (sample && use_smem) {
            int lx = (int) qx[0], hx = lx, ly = (int) qy[0], hy = ly;
#pragma unroll
            for (int i = 1; i < 8; i++) { const int ix = (int) qx[i], iy = (int) qy[i]; lx = min(lx, ix); hx = max(hx, ix); ly = min(ly, iy); hy = max(hy, iy); }
            use_smem = lx >= box_x && ly >= box_y && hx + 1 < box_x + LT_BOX_W && hy + 1 < box_y + LT_BOX_H;
            if (use_smem) { ox = box_x; oy = box_y; pitch = LT_BOX_W; tbase = reinterpret_cast<const float4 *>(ring + (size_t) ((job - q0) % LT_STAGES) * LT_STAGE_STRIDE); }
        }
        if (trace) {     // development: lanes that sample / lanes that fall back to the image (slots 28, 29 of the warp's trace row)
            const unsigned ms = __ballot_sync(0xffffffffu, sample), mf = __ballot_sync(0xffffffffu, sample && !use_smem);
            if (lane == 0) { trace[28] += __popc(ms); trace[29] += __popc(mf); }
        }
        if ((w.lt_mode & 4) && !use_smem) sample = false;     // development: what the image fallback costs (results are wrong in this mode)
        // ---- per-residual sampling + Jacobians
        float rec[RJ_STRIDE];
        float trow[T_STRIDE];
#pragma unroll
        for (int k = 0; k < RJ_STRIDE; k++) rec[k] = 0.f;
#pragma unroll
        for (int k = 0; k < T_STRIDE; k++) trow[k] = 0.f;
        float sI[8], sgx[8], sgy[8];
        if (sample) {
#pragma unroll
            for (int i = 0; i < 8; i++) {      // all 32 taps first (the tile can go back to the producer right after)
                const int ix = (int) qx[i], iy = (int) qy[i];
                const float4 *tp = tbase + ((iy - oy) * pitch + (ix - ox));
                const float4 t00 = tp[0], t10 = tp[1], t01 = tp[pitch], t11 = tp[pitch + 1];
                const float dx = qx[i] - (float) ix, dy = qy[i] - (float) iy;
                const float dxdy = dx * dy;
                const float w00 = 1.f - dx - dy + dxdy, w10 = dx - dxdy, w01 = dy - dxdy, w11 = dxdy;
                sI[i] = t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11;
                sgx[i] = t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11;
                sgy[i] = t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11;
            }
        }
        LT_STAMPK(tr_b + 3);                         // taps done
        if (c + LT_CWARPS < c1) {                     // the point records of this warp's next pass: on their way to L2 while this one computes
            const int pn = (int) (pht_next & 0xffffffu);
            prefetch_l2(w.pt_idepth + pn);
            if (lane < 20) prefetch_l2(w.r_pt4 + (size_t) (lane >> 2) * w.R + min((c + LT_CWARPS) * 32 + (lane & 3) * 8, w.R - 1));
        }
        // the taps are in registers: this pass is done with its tiles.  Tiles beyond the ring window (very sparse windows only; their
        // lanes read global memory) still count this pass as a user: wait for them and arrive, one by one.
        __syncwarp();
        if (tma_ok) {
            if (lane == 0) for (int q = q_lo; q <= q_hi; q++) release(q - q0);
            for (int q = q_hi + 1; q <= q_last; q++) {
                const int k = q - q0;
                if (!lt_wait_tile(s_tag, full, k % LT_STAGES, q, (uint32_t) ((k / LT_STAGES) & 1))) { tma_ok = false; break; }
                __syncwarp();
                if (lane == 0) release(k);
            }
        }
        if (sample) {
            const float col[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
            const float wts[8] = {w0v.x, w0v.y, w0v.z, w0v.w, w1v.x, w1v.y, w1v.z, w1v.w};
            const float b0 = pp.b0;
            const float sqrt_cth = sqrtf(w.cth);
            float J00 = 0, J11 = 0, J10 = 0, A00 = 0, A01 = 0, A10 = 0, A11 = 0, B00 = 0, B01 = 0, B11 = 0, wJI2 = 0, E = 0;
            float JIr0 = 0, JIr1 = 0, Jabr0 = 0, Jabr1 = 0, rr = 0;
            bool finite = true;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float I = sI[i], gx = sgx[i], gy = sgy[i];
                finite = finite && isfinite(I) && isfinite(gx) && isfinite(gy);
                const float refReal = (float) (pp.a * (double) col[i] + pp.b);     // exposureTransition (BA:229)
                const float res = I - refReal;
                const float ar = fabsf(res);
                // MUFU-based division / rsqrt / sqrt (<= 2 ulp, ~2e-7 relative; the parity tolerance is 1e-4)
                float hw = ar < w.huber ? 1.f : __fdividef(w.huber, ar);           // BA:233
                float wg = sqrt_cth * rsqrtf(w.cth + (gx * gx + gy * gy));         // BA:234  sqrt(c / (c + |grad|^2))
                wg = 0.5f * (wg + wts[i]);                                         // BA:235
                E += wg * wg * hw * res * res * (2.f - hw);                        // BA:237
                if (hw < 1.f) hw = __fsqrt_rn(hw);
                hw = hw * wg;
                const float h1 = gx * hw, h2 = gy * hw, drdA = I - b0;
                const float rF = res * hw;
                const float ja = (w.optA ? drdA * hw : 0.f), jb = (w.optB ? hw : 0.f);   // BA:273-278 (zeroed after the sums below)
                J00 += h1 * h1; J11 += h2 * h2; J10 += h1 * h2;
                A00 += drdA * hw * h1; A01 += drdA * hw * h2; A10 += hw * h1; A11 += hw * h2;
                B00 += drdA * drdA * hw * hw; B01 += drdA * hw * hw; B11 += hw * hw;
                wJI2 += hw * hw * (h1 * h1 + h2 * h2);
                JIr0 += rF * h1; JIr1 += rF * h2; Jabr0 += rF * ja; Jabr1 += rF * jb; rr += rF * rF;   // BA:1722-1729
                if (kDump) {
                    float *d = w.dbg + (size_t) r * DBG_STRIDE;
                    d[i] = rF; d[8 + i] = h1; d[16 + i] = h2; d[24 + i] = ja; d[32 + i] = jb;
                }
            }
            LT_STAMPK(tr_b + 4);                         // photometric sums done
            if (!finite) {
                // BA:220-223 sets the *committed* state to OOB.  (The reference leaves a stale isActiveAndIsGoodNEW
                // behind in that case; only reachable with NaN/Inf texels, we clear it.)
                st_out = RES_OOB;
            } else if (!isfinite(E)) {
                nst = RES_OOB;               // BA:297-300
            } else {
                neo = E;
                const float th = fmaxf(s_th[h], s_th[t]);
                if (E > th || wJI2 < 2.f) { E = th; nst = RES_OUTLIER; } else nst = RES_IN;   // BA:303-311
                ne = E;
                ret = (double) E;
                if (nst == RES_IN) {
                    // ---- geometric Jacobians at the FEJ point (BA:120-188); note u,v are the UN-normalised P.xy (BA:121-122)
                    const float drescale = (float) (1.0 / Pc2);
                    const float new_idepth = (float) ((double) drescale * rho);
                    const float u = (float) Pc0, v = (float) Pc1;
                    const float fxf = (float) w.fx, fyf = (float) w.fy;
                    const double ud = (double) u, vd = (double) v, dr = (double) drescale;
                    const double klx = (xc - w.cx) * w.fxi, kly = (yc - w.cy) * w.fyi;      // KliP
                    const float Jpdd0 = (float) (dr * (pp.t0[0] - pp.t0[2] * ud) * (double) fxf);
                    const float Jpdd1 = (float) (dr * (pp.t0[1] - pp.t0[2] * vd) * (double) fyf);
                    double dCx[4], dCy[4];
                    dCx[2] = dr * (pp.R0[6] * ud - pp.R0[0]);
                    dCx[3] = (double) (fxf * drescale) * (pp.R0[7] * ud - pp.R0[1]) / (double) fyf;
                    dCx[0] = klx * dCx[2]; dCx[1] = kly * dCx[3];
                    dCy[2] = (double) (fyf * drescale) * (pp.R0[6] * vd - pp.R0[3]) / (double) fxf;
                    dCy[3] = dr * (pp.R0[7] * vd - pp.R0[4]);
                    dCy[0] = klx * dCy[2]; dCy[1] = kly * dCy[3];
                    const double sF = (double) w.scaleF, sC = (double) w.scaleC;
                    dCx[0] = (dCx[0] + ud) * sF; dCx[1] *= sF; dCx[2] = (dCx[2] + 1.0) * sC; dCx[3] *= sC;
                    dCy[0] *= sF; dCy[1] = (dCy[1] + vd) * sF; dCy[2] *= sC; dCy[3] = (dCy[3] + 1.0) * sC;
                    // record: x = [Jpdc_x | Jpdxi_x], y = [Jpdc_y | Jpdxi_y]
                    rec[0] = (float) dCx[0]; rec[1] = (float) dCx[1]; rec[2] = (float) dCx[2]; rec[3] = (float) dCx[3];
                    rec[4] = new_idepth * fxf; rec[5] = 0.f; rec[6] = -new_idepth * u * fxf; rec[7] = -u * v * fxf; rec[8] = (1.f + u * u) * fxf; rec[9] = -v * fxf;
                    rec[10] = (float) dCy[0]; rec[11] = (float) dCy[1]; rec[12] = (float) dCy[2]; rec[13] = (float) dCy[3];
                    rec[14] = 0.f; rec[15] = new_idepth * fyf; rec[16] = -new_idepth * v * fyf; rec[17] = -(1.f + v * v) * fyf; rec[18] = u * v * fyf; rec[19] = u * fyf;
                    if (w.marg_mode) {
                        // MARGINALIZED accumulation (BA:1686-1690): the residual vector is res_toZeroF = resF - [JI*Jp Jab]*delta
                        // (fixLinearization, BA:2210-2238).  Its moments follow from the sums above; JabF is zeroed for a fixed a / b.
                        const float *dp = w.pair_delta + (size_t) (h * N + t) * 8;
                        const float dF = (float) (rho - (double) w.pt_idepth_zero[p]);
                        float jx = Jpdd0 * dF, jy = Jpdd1 * dF;
#pragma unroll
                        for (int k = 0; k < 6; k++) { jx += rec[4 + k] * dp[k]; jy += rec[14 + k] * dp[k]; }
                        const float da = dp[6], db = dp[7];
                        const float a00 = w.optA ? A00 : 0.f, a01 = w.optA ? A01 : 0.f, a10 = w.optB ? A10 : 0.f, a11 = w.optB ? A11 : 0.f;
                        const float b00 = w.optA ? B00 : 0.f, b01 = (w.optA && w.optB) ? B01 : 0.f, b11 = w.optB ? B11 : 0.f;
                        const float cross = JIr0 * jx + JIr1 * jy + Jabr0 * da + Jabr1 * db;
                        const float gx_ = J00 * jx + J10 * jy + a00 * da + a10 * db, gy_ = J10 * jx + J11 * jy + a01 * da + a11 * db;
                        const float ga = a00 * jx + a01 * jy + b00 * da + b01 * db, gb = a10 * jx + a11 * jy + b01 * da + b11 * db;
                        rr = rr - 2.f * cross + (jx * gx_ + jy * gy_ + da * ga + db * gb);
                        JIr0 -= gx_; JIr1 -= gy_; Jabr0 -= ga; Jabr1 -= gb;
                    }
                    rec[20] = J00; rec[21] = J10; rec[22] = J11;                    // JIdx2
                    rec[23] = A00; rec[24] = A10; rec[25] = JIr0;                   // x-multipliers of columns a, b, r (BA:1740-1745)
                    rec[26] = A01; rec[27] = A11; rec[28] = JIr1;                   // y-multipliers
                    rec[29] = B00; rec[30] = B01; rec[31] = Jabr0; rec[32] = B11; rec[33] = Jabr1; rec[34] = rr;   // BA:1736-1738
                    // applyRes (BA:2066-2080) and the per-point sums of addToHessianTop (BA:1747-1750)
                    const float v0 = J00 * Jpdd0 + J10 * Jpdd1, v1 = J10 * Jpdd0 + J11 * Jpdd1;
#pragma unroll
                    for (int k = 0; k < 6; k++) trow[k] = rec[4 + k] * v0 + rec[14 + k] * v1;
                    trow[6] = A00 * Jpdd0 + A01 * Jpdd1;
                    trow[7] = A10 * Jpdd0 + A11 * Jpdd1;
                    trow[8] = JIr0 * Jpdd0 + JIr1 * Jpdd1;                          // bd
                    trow[9] = v0 * Jpdd0 + v1 * Jpdd1;                              // Hdd
#pragma unroll
                    for (int k = 0; k < 4; k++) trow[10 + k] = rec[k] * v0 + rec[10 + k] * v1;   // Hcd
                    trow[14] = 1.f;
                    good = true;
                    if (kDump) {
                        float *d = w.dbg + (size_t) r * DBG_STRIDE;
                        d[40] = Jpdd0; d[41] = Jpdd1; d[42] = J00; d[43] = J10; d[44] = J11;
                        d[45] = A00; d[46] = A01; d[47] = A10; d[48] = A11; d[49] = B00; d[50] = B01; d[51] = B11;
                    }
                    if (fix) {   // BA:1571-1592: relative baseline, numGoodResiduals
                        const double Rk0 = pp.R[0] * klx + pp.R[1] * kly + pp.R[2], Rk1 = pp.R[3] * klx + pp.R[4] * kly + pp.R[5], Rk2 = pp.R[6] * klx + pp.R[7] * kly + pp.R[8];
                        const double ix_ = (Rk0 / Rk2) * w.fx + w.cx, iy_ = (Rk1 / Rk2) * w.fy + w.cy;
                        const double ddx = ix_ - Kuc, ddy = iy_ - Kvc;
                        const float relBS = (float) (0.01 * sqrt(ddx * ddx + ddy * ddy));
                        atomicMax(reinterpret_cast<int *>(w.pt_max_rel_bs + p), __float_as_int(relBS));
                        atomicAdd(w.pt_num_good + p, 1);
                    }
                }
            }
        }
        LT_STAMPK(tr_b + 5);                         // classification, Jacobians done
        if (w.lt_mode & 8) { const double es = warp_sum_d(ret); if (lane == 0) w.energy_part[c] = es; continue; }     // development: no per-residual stores
        if (valid) {
            // applyRes (BA:2051-2093), as the candidate that becomes current when the step is accepted
            if (st != RES_OOB && st_out != RES_OOB) { st_out = nst; e_out = ne; }
            w.r_new_state[r] = nst; w.r_new_energy[r] = ne; w.r_new_energy_wo[r] = neo;
            w.r_state[nxt][r] = st_out; w.r_energy[nxt][r] = e_out; w.r_good[nxt][r] = good ? 1 : 0;
            if (fix && !good) {                      // BA:1595-1598, 1623-1640: non-good residuals are deleted
                w.r_alive[r] = 0;
                atomicAdd(&ctrl->num_dropped, 1);
            }
        }
        if (fix && in_chunk) {                       // final states in the host's residual order (finish_run reads these)
            w.fin_state[src] = st_out; w.fin_energy[src] = e_out; w.fin_alive[src] = (valid && good) ? 1 : 0;
        }
        // ---- the Jacobian record (the reference's efsJ) in the host's residual order: what addToHessianTop is evaluated from
        // (accumulate_kernel).  rec[35] = 1 marks a good residual; the record of any other residual is all zeros.  A record is 144
        // contiguous bytes but the records of a warp are scattered: the warp stages 16 records at a time in shared memory and writes
        // them out with 9 consecutive lanes per record (4-5 lines per store instruction instead of 32).
        {   // the Schur row of (point, target): 64 contiguous bytes per lane, staged and written out by four consecutive lanes per row
            const int row_or_none = valid ? p * N + t : -1;
            float4 *mine = s_stage + lane * (T_STRIDE / 4 + 1);       // +1: rows 80 bytes apart, conflict-free for the quarter-warp phases of a 16-byte store
#pragma unroll
            for (int k = 0; k < T_STRIDE / 4; k++) mine[k] = make_float4(trow[4 * k], trow[4 * k + 1], trow[4 * k + 2], trow[4 * k + 3]);
            __syncwarp();
#pragma unroll 2
            for (int j = 0; j < T_STRIDE / 4; j++) {
                const int f = lane + 32 * j, m = f >> 2, k = f & 3;
                const int rowm = __shfl_sync(0xffffffffu, row_or_none, m);
                if (rowm >= 0) reinterpret_cast<float4 *>(w.T[nxt] + (size_t) rowm * T_STRIDE)[k] = s_stage[m * (T_STRIDE / 4 + 1) + k];
            }
            __syncwarp();
        }
#ifndef LT_NO_COOP
        {
            const int src_or_none = in_chunk ? src : -1;
            float4 *mine = s_stage + lane * (RJ_STRIDE / 4);
#pragma unroll
            for (int k = 0; k < RJ_STRIDE / 4 - 1; k++) mine[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
            mine[RJ_STRIDE / 4 - 1] = make_float4(rec[32], rec[33], rec[34], good ? 1.f : 0.f);
            __syncwarp();
#pragma unroll 3
            for (int j = 0; j < RJ_STRIDE / 4; j++) {
                const int f = lane + 32 * j, m = f / (RJ_STRIDE / 4), k = f - m * (RJ_STRIDE / 4);
                const int srcm = __shfl_sync(0xffffffffu, src_or_none, m);
                if (srcm >= 0) reinterpret_cast<float4 *>(w.rj[nxt] + (size_t) srcm * RJ_STRIDE)[k] = s_stage[f];
            }
            __syncwarp();
        }
#else
        if (in_chunk) {
            float4 *rj4 = reinterpret_cast<float4 *>(w.rj[nxt] + (size_t) src * RJ_STRIDE);
#pragma unroll
            for (int k = 0; k < RJ_STRIDE / 4 - 1; k++) rj4[k] = make_float4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
            rj4[RJ_STRIDE / 4 - 1] = make_float4(rec[32], rec[33], rec[34], good ? 1.f : 0.f);
        }
#endif
        LT_STAMPK(tr_b + 6);                         // pass done
        // chunk energy (fp64, fixed order)
        const double es = warp_sum_d(ret);
        if (lane == 0) w.energy_part[c] = es;
    }
    if (trace && lane == 0) { long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); trace[31] = gt; }
#undef LT_STAMPK
}

}  // namespace cmlba
