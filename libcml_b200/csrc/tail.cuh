// tail.cuh -- the reduced camera system in ONE launch: Schur complement of the inverse depths (addToHessianSC, BA:1880-1937), stitching of
// the active and Schur parts through the adjoints (stitchDoubleTop BA:1781-1878, stitchDoubleSC BA:1939-2043) and the assembly of
// sys = [HA | bA | H_sc | b_sc].  Replaces the schur -> stitch_pair -> assemble chain (three dependent launches, partial sums through
// global memory) whenever a cluster of >= N CTAs can be scheduled; the three kernels stay as the fallback.
//
// One thread-block cluster per host frame i; CTA rank r of the cluster owns frame r.
//   phase 1  every CTA takes every CS-th chunk of <= 64 points of host i and accumulates the rank update of the augmented per-point
//            vectors z = s [JpJdF over all targets | Hcd | bdSum] (see schur_kernel) in register tiles, then parks the (8N+5)^2
//            partial in its shared memory;
//   phase 2  CTA r sums rows 8r..8r+7 of the CS partials through distributed shared memory (fp64, rank order), adds the 13x13 block of
//            bin (i -> r) from the sampling kernel's partial blocks and writes every block product of the pair (i, r) to its st_out slot
//            (rank i: the calibration block of host i);
//   phase 3  the cluster that finishes last (one ticket per cluster) gathers all slots into sys, its CTAs sharing the elements; with the
//            peer-memory exchange on, it pushes them to every rank and publishes the epoch.
// Fixed summation orders throughout: bitwise reproducible.
#pragma once
#include <cooperative_groups.h>

namespace cmlba {
namespace cg = cooperative_groups;

constexpr int TAIL_THREADS = 512;
constexpr int TAIL_TPT = 2;          // 4x4 tiles of the rank update per thread (2N(2N+1)/2 + 4N tiles: 168 at N = 8, 592 at N = 16)

__host__ __device__ __forceinline__ size_t tail_smem_bytes(int N) {
    const size_t NB = 8 * (size_t) N;
    const size_t phase1 = sizeof(float) * ((size_t) SC_CHUNK * N * T_STRIDE + (size_t) SC_CHUNK * (NB + SCZ_PAD));
    const size_t phase2 = sizeof(double) * (8 * NB + 40 + (size_t) N * 64 + NB + ACC_N + 128);
    const size_t part = sizeof(float) * (((NB * NB + NB * 5 + 20) + 3) & ~(size_t) 3);
    return (phase1 > phase2 ? phase1 : phase2) + part;
}

// element e of sys from the st_out slots (the body of assemble_kernel, one element per call)
// sum_k!=a slot(a,k)[o_row] + sum_k!=a slot(k,a)[o_col], all loads issued before the (fixed-order) adds
__device__ __forceinline__ double sum_slots(const double *st, const int N, const int S, const int a, const int o_row, const int o_col) {
    return gather_seq(st, N, a * N * S + o_row, S, a, -1) + gather_seq(st, N, a * S + o_col, N * S, a, -1);
}
__device__ __forceinline__ void assemble_element(const DevWin &w, const bool p2p, const int e) {
    const int N = w.N, n = w.n, nn = n * n, S = st_stride(N);
    const double *st = w.st_out;
#define SLOT(i, j) (st + (size_t) ((i) * N + (j)) * S)
    const bool schur = e >= nn + n;
    const int q = schur ? e - nn - n : e;
    double v = 0.0;
    if (q < nn) {
        int r = q / n, c = q - r * n;
        if (r < 4 && c >= 4) { const int t = r; r = c; c = t; }          // calibration rows mirror the columns
        if (r < 4) {                                                      // (C,C)
            if (!schur) {                                                 // HA[C,C]: over all ordered pairs
                for (int k = 0; k < N * N; k++) { const int i = k / N, j = k - i * N; if (i != j) v += __ldcg(SLOT(i, j) + ST_A_CC + r * 4 + c); }
            } else for (int i = 0; i < N; i++) v += __ldcg(SLOT(i, i) + r * 4 + c);
        } else if (c < 4) {                                               // (frame a, C)
            const int a = (r - 4) >> 3, rr = (r - 4) & 7;
            const int o_i = (schur ? ST_S_IC : ST_A_IC) + rr * 4 + c, o_t = (schur ? ST_S_JC : ST_A_TC) + rr * 4 + c;
            v = sum_slots(st, N, S, a, o_i, o_t);
        } else {
            const int a = (r - 4) >> 3, rr = (r - 4) & 7, b = (c - 4) >> 3, cc = (c - 4) & 7;
            if (a == b) {
                if (!schur) v = sum_slots(st, N, S, a, ST_A_II + rr * 8 + cc, ST_A_TT + rr * 8 + cc);
                else v = sum_slots(st, N, S, a, ST_S_II + rr * 8 + cc, ST_S_JK + a * 64 + rr * 8 + cc);
            } else {
                const int lo = a < b ? a : b, hi = a < b ? b : a, rl = a < b ? rr : cc, rh = a < b ? cc : rr;   // element (lo rl, hi rh)
                if (!schur) v = __ldcg(SLOT(lo, hi) + ST_A_IT + rl * 8 + rh) + __ldcg(SLOT(hi, lo) + ST_A_IT + rh * 8 + rl);
                else {
                    const double x1 = __ldcg(SLOT(hi, lo) + ST_S_JI + rl * 8 + rh), x2 = __ldcg(SLOT(lo, hi) + ST_S_JI + rh * 8 + rl);
                    for (int k = 0; k < N; k++) if (k != lo && k != hi) v += __ldcg(SLOT(k, lo) + ST_S_JK + hi * 64 + rl * 8 + rh);
                    v += x1;
                    v += x2;
                }
            }
        }
    } else {
        const int r = q - nn;
        if (r < 4) {
            if (!schur) { for (int k = 0; k < N * N; k++) { const int i = k / N, j = k - i * N; if (i != j) v += __ldcg(SLOT(i, j) + ST_BA_C + r); } }
            else for (int i = 0; i < N; i++) v += __ldcg(SLOT(i, i) + 16 + r);
        } else {
            const int a = (r - 4) >> 3, rr = (r - 4) & 7;
            const int o_i = (schur ? ST_BS_I : ST_BA_I) + rr, o_t = (schur ? ST_BS_J : ST_BA_T) + rr;
            v = sum_slots(st, N, S, a, o_i, o_t);
        }
    }
#undef SLOT
    sys_emit(w, p2p, e, v);
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) tail_kernel(const DevWin w, const int respect_done) {
    pdl_enter();
    if (respect_done && w.ctrl->done) return;          // uniform over the grid
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int) cluster.num_blocks(), rank = (int) cluster.block_rank();
    extern __shared__ __align__(16) unsigned char tail_smem[];
    const int N = w.N, NB = 8 * N, ZS = NB + SCZ_PAD, tid = threadIdx.x, i = blockIdx.x / CS;
    const int cur = w.ctrl->cur;
    const size_t phase1 = sizeof(float) * ((size_t) SC_CHUNK * N * T_STRIDE + (size_t) SC_CHUNK * ZS);
    const size_t phase2 = sizeof(double) * ((size_t) 8 * NB + 40 + (size_t) N * 64 + NB + ACC_N + 128);
    float *part = reinterpret_cast<float *>(tail_smem + (phase1 > phase2 ? phase1 : phase2));      // this CTA's Schur partial, layout of sc_part
    float *sT = reinterpret_cast<float *>(tail_smem);                  // [SC_CHUNK][N][T_STRIDE] raw Schur rows
    float *sZ = sT + SC_CHUNK * N * T_STRIDE;                          // [SC_CHUNK][ZS]          augmented, scaled
    const int ntr = 2 * N, ntri = ntr * (ntr + 1) / 2, ntiles = ntri + 2 * ntr;
    // ---- phase 1
    float acc[TAIL_TPT][4][4];
    int ttr[TAIL_TPT], ttc[TAIL_TPT];
#pragma unroll
    for (int u = 0; u < TAIL_TPT; u++) {
        const int id = tid + u * TAIL_THREADS;
        int tr = 0, tc = 0;
        if (id < ntri) {
            tr = (int) ((sqrtf(8.f * (float) id + 1.f) - 1.f) * 0.5f);
            while (tr * (tr + 1) / 2 > id) tr--;
            while ((tr + 1) * (tr + 2) / 2 <= id) tr++;
            tc = id - tr * (tr + 1) / 2;
        } else if (id < ntiles) { tr = (id - ntri) >> 1; tc = ntr + ((id - ntri) & 1); }
        ttr[u] = tr; ttc[u] = tc;
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[u][a][b] = 0.f;
    }
    float hcc = 0.f;                                    // threads 480..499: Hcc (4x4), bc (4)
    const int cb = w.host_chunk_begin[i], ce = w.host_chunk_begin[i + 1];
    for (int c = cb + rank; c < ce; c += CS) {
        const int begin = w.sc_chunk_begin[c], cnt = w.sc_chunk_count[c];
        {   // stage the rows of the chunk's points (contiguous in T)
            const float4 *src = reinterpret_cast<const float4 *>(w.T[cur] + (size_t) begin * N * T_STRIDE);
            float4 *dst = reinterpret_cast<float4 *>(sT);
            const int n4 = cnt * N * (T_STRIDE / 4), tot4 = SC_CHUNK * N * (T_STRIDE / 4);
            for (int k = tid; k < tot4; k += TAIL_THREADS) dst[k] = k < n4 ? __ldg(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (tid < SC_CHUNK) {   // per point sums (BA:1895-1907)
            float sh = 0.f, hc0 = 0.f, hc1 = 0.f, hc2 = 0.f, hc3 = 0.f, bs = 0.f;
            if (tid < cnt) {
                const int p = begin + tid;
                float Hdd = 0.f, bd = 0.f, h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f; int ng = 0;
                for (int t = 0; t < N; t++) {
                    const float4 a = *reinterpret_cast<const float4 *>(sT + (tid * N + t) * T_STRIDE + 8);    // bd Hdd Hcd0 Hcd1
                    const float4 b = *reinterpret_cast<const float4 *>(sT + (tid * N + t) * T_STRIDE + 12);   // Hcd2 Hcd3 good pad
                    bd += a.x; Hdd += a.y; h0 += a.z; h1 += a.w; h2 += b.x; h3 += b.y; ng += (b.z != 0.f);
                }
                const float priorF = w.pt_priorF[p];
                float idh = 0.f, bdSum = 0.f, hd = 0.f;
                if (ng > 0) {
                    float Hs = Hdd + priorF;
                    if (Hs < 1e-10f) Hs = 1e-10f;
                    idh = Hs;
                    hd = (float) (1.0 / (double) Hs);
                    const float deltaF = (float) (w.pt_idepth[p] - (double) w.pt_idepth_zero[p]);
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
            sT[tid * N * T_STRIDE + 15] = sh;            // pad slot of the point's first row carries the scale to the next step
        }
        __syncthreads();
        for (int k = tid; k < SC_CHUNK * N * 2; k += TAIL_THREADS) {      // z_u = s * JpJdF
            const int row = k >> 1, half = k & 1, pnt = row / N, t = row - pnt * N;
            const float sh = sT[pnt * N * T_STRIDE + 15];
            float4 v = *reinterpret_cast<const float4 *>(sT + row * T_STRIDE + half * 4);
            v.x *= sh; v.y *= sh; v.z *= sh; v.w *= sh;
            *reinterpret_cast<float4 *>(sZ + pnt * ZS + t * 8 + half * 4) = v;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < TAIL_TPT; u++) {
            if (tid + u * TAIL_THREADS < ntiles) {
                const float *pa = sZ + ttr[u] * 4, *pb = sZ + ttc[u] * 4;
#pragma unroll 4
                for (int pnt = 0; pnt < SC_CHUNK; pnt++) {
                    const float4 a = *reinterpret_cast<const float4 *>(pa + pnt * ZS);
                    const float4 b = *reinterpret_cast<const float4 *>(pb + pnt * ZS);
                    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) acc[u][x][y] += av[x] * bv[y];
                }
            }
        }
        if (tid >= 480 && tid < 500) {       // Hcc (4x4), bc (4)
            const int k = tid - 480;
            const int x = k < 16 ? (k >> 2) : (k - 16), y = k < 16 ? (k & 3) : 4;
            for (int pnt = 0; pnt < SC_CHUNK; pnt++) hcc += sZ[pnt * ZS + NB + x] * sZ[pnt * ZS + NB + y];
        }
        __syncthreads();
    }
    {   // park the partial (layout of sc_part: D[NB][NB] | E[NB][4] | EB[NB] | Hcc[16] | bc[4])
        float *oE = part + NB * NB, *oEB = oE + NB * 4, *oHcc = oEB + NB, *obc = oHcc + 16;
#pragma unroll
        for (int u = 0; u < TAIL_TPT; u++) {
            if (tid + u * TAIL_THREADS < ntiles) {
                const int tr = ttr[u], tc = ttc[u];
                if (tc < ntr) {
#pragma unroll
                    for (int x = 0; x < 4; x++) *reinterpret_cast<float4 *>(part + (size_t) (tr * 4 + x) * NB + tc * 4) = make_float4(acc[u][x][0], acc[u][x][1], acc[u][x][2], acc[u][x][3]);
                    if (tc < tr) {
#pragma unroll
                        for (int y = 0; y < 4; y++) *reinterpret_cast<float4 *>(part + (size_t) (tc * 4 + y) * NB + tr * 4) = make_float4(acc[u][0][y], acc[u][1][y], acc[u][2][y], acc[u][3][y]);
                    }
                } else if (tc == ntr) {
#pragma unroll
                    for (int x = 0; x < 4; x++) *reinterpret_cast<float4 *>(oE + (tr * 4 + x) * 4) = make_float4(acc[u][x][0], acc[u][x][1], acc[u][x][2], acc[u][x][3]);
                } else {
#pragma unroll
                    for (int x = 0; x < 4; x++) oEB[tr * 4 + x] = acc[u][x][0];
                }
            }
        }
        if (tid >= 480 && tid < 500) { const int k = tid - 480; if (k < 16) oHcc[k] = hcc; else obc[k - 16] = hcc; }
    }
    cluster.sync();                                     // every partial of host i is in place
    // ---- phase 2: pair (host i, frame j = rank)
    const int j = rank;
    double *smd = reinterpret_cast<double *>(tail_smem);
    if (j < N) {
        double *out = w.st_out + (size_t) (i * N + j) * st_stride(N);
        if (i == j) {   // calibration block of host i's Schur complement (BA:1908-1909, 2026-2027)
            if (tid < 20) {
                double s = 0.0;
                for (int k = 0; k < CS; k++) s += (double) cluster.map_shared_rank(part, k)[NB * NB + NB * 5 + tid];
                out[tid] = s;
            }
        } else {
            double *Dj = smd;              // [8][NB]  rows of frame j of D_i
            double *Ej = Dj + 8 * NB;      // [8][4]   (Dj, Ej, EBj contiguous: filled by one loop)
            double *EBj = Ej + 32;         // [8]
            double *G = EBj + 8;           // [N][8][8] AH_ik
            double *atd = G + N * 64;      // [N][8]   diag(AT_ik)
            double *A = atd + NB;          // [ACC_N]  packed 13x13 block of bin (i -> j)
            double *Y = A + ACC_N;         // [8][8]   sum_k D_jk AH_ik^T
            double *M = Y + 64;            // [8][8]   AH_ij A8
            for (int e = tid; e < 8 * NB + 40; e += TAIL_THREADS) {
                int off;
                if (e < 8 * NB) off = j * 8 * NB + e;
                else if (e < 8 * NB + 32) off = NB * NB + j * 32 + (e - 8 * NB);
                else off = NB * NB + NB * 4 + j * 8 + (e - 8 * NB - 32);
                double s = 0.0;
                for (int k = 0; k < CS; k++) s += (double) cluster.map_shared_rank(part, k)[off];
                Dj[e] = s;                 // Dj, Ej, EBj are contiguous
            }
            if (tid < ACC_N) {   // 13x13 block of bin (i -> j): the slices of accumulate_kernel
                const float *src = w.acc_bin + (size_t) (j * N + i) * ACC_SLICES * ACC_N + tid;
                double a = 0.0;
                for (int sl = 0; sl < ACC_SLICES; sl++) a += (double) __ldcg(src + sl * ACC_N);
                A[tid] = a;
            }
            for (int e = tid; e < N * 64; e += TAIL_THREADS) G[e] = w.AH[(size_t) (i * N) * 64 + e];
            for (int e = tid; e < NB; e += TAIL_THREADS) atd[e] = w.AT[((size_t) (i * N + (e >> 3))) * 64 + (e & 7) * 9];
            __syncthreads();
            const double *AHj = G + j * 64, *atj = atd + j * 8;
            {
                const int r = (tid >> 3) & 7, c = tid & 7;
                double s = 0.0;
                if (tid < 64) { for (int k = 0; k < N; k++) for (int m = 0; m < 8; m++) s += Dj[r * NB + k * 8 + m] * G[k * 64 + c * 8 + m]; Y[tid] = s; }
                else if (tid < 128) { for (int m = 0; m < 8; m++) s += AHj[r * 8 + m] * A[acc_index(4 + m, 4 + c)]; M[tid - 64] = s; }
            }
            __syncthreads();
            const int tot = st_stride(N);
            for (int e = tid; e < tot; e += TAIL_THREADS) {
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
    }
    // ---- phase 3: the last cluster assembles sys
    __threadfence();                                    // this CTA's slot is visible device-wide before the ticket is taken
    cluster.sync();                                     // (also: nobody reads the parked partials any more)
    __shared__ int s_last;
    if (rank == 0 && tid == 0) {
        const int ticket = atomicAdd(&w.ctrl->asm_done_count, 1);
        const int last = ticket == (int) gridDim.x / CS - 1;
        if (last) w.ctrl->asm_done_count = 0;
        s_last = last;
    }
    cluster.sync();
    const int is_last = *cluster.map_shared_rank(&s_last, 0);
    cluster.sync();                                     // rank 0 must not leave while its flag is being read
    if (!is_last) return;
    __threadfence();
    const bool p2p = w.p2p_on && w.world > 1;
    const int nelem = 2 * w.n * w.n + 2 * w.n;
    for (int e = rank * TAIL_THREADS + tid; e < nelem; e += CS * TAIL_THREADS) assemble_element(w, p2p, e);
    if (p2p) {                                           // every element of this rank's system has been pushed: publish the epoch
        __threadfence_system();
        cluster.sync();
        if (rank == 0 && tid == 0) { __threadfence_system(); p2p_signal(w); }
    }
}

}  // namespace cmlba
