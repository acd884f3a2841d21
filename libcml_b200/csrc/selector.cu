// selector.cu -- DSO pixel selector on the device and the C ABI of include/cmlsel.h (SURVEY.md 8f NEXT #4, PixelSelector part).
//
// Reference anchors: /root/reference/src/cml/features/corner/PixelSelector.cpp
//   sel_hist_kernel / sel_smooth_kernel   :30-118 (computeHistQuantil, makeHists)
//   sel_block_kernel                      :217-365 (select)
//   Selector::make_maps                   :121-213 (makeMaps: potential adaptation, recursion, random sub-sampling)
//   Selector::compute                     :367-384
//
// B200 design.  select() is a sequential sweep: the random direction of every block is pattern[n2], n2 = the number of level-1
// selections made so far.  Inside a 4*pot block everything else is local, so one warp evaluates one 4*pot block exactly (the sticky -2
// flags reduce to nesting rules, the strict > comparisons to first-in-traversal-order arg-max), starting from the n2 of its block; those
// starting values are an exclusive prefix sum of the per-block selection counts, which themselves depend (rarely: only when a candidate's gradient is exactly orthogonal to the
// drawn direction) on the directions.  The host iterates simulate -> scan until the counts reproduce themselves; blocks before the first
// disagreement are already final, so the iteration converges and the result is exactly the sequential one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmlsel.h"

namespace cmlsel {

struct SelDev {
    int w, h, w32, h32;
    const float4 *t0, *t1, *t2;     // texel levels 0..2
    int w1, w2;
    const unsigned char *pattern;
    float *ths, *ths_smoothed;      // flat, (w32 * h32 + 100) entries like the reference's arrays
    float *map;                     // [h][w]
    int *hits, *start;              // per 4*pot block: level-1 selections of the last simulation, exclusive prefix
    int *counters;                  // [0..2] n2 n3 n4, [3] blocks whose count changed
    int *row_count, *row_offset;    // raster compaction of the map
    int2 *list;                     // (pixel index, type) in raster order
    int *list_n;
};

__constant__ float c_dir[16][2] = {{0.f, 1.0000f}, {0.3827f, 0.9239f}, {0.1951f, 0.9808f}, {0.9239f, 0.3827f}, {0.7071f, 0.7071f}, {0.3827f, -0.9239f}, {0.8315f, 0.5556f},
                                   {0.8315f, -0.5556f}, {0.5556f, -0.8315f}, {0.9808f, 0.1951f}, {0.9239f, -0.3827f}, {0.7071f, -0.7071f}, {0.5556f, 0.8315f}, {0.9808f, -0.1951f},
                                   {1.0000f, 0.0000f}, {0.1951f, -0.9808f}};

// one CTA per 32x32 block: histogram of int(sqrt(weighted gradient norm)), threshold = median bin + 7
__global__ void __launch_bounds__(256) sel_hist_kernel(const SelDev s) {
    __shared__ int hist[50];
    const int bx = blockIdx.x, by = blockIdx.y, tid = threadIdx.x;
    if (tid < 50) hist[tid] = 0;
    __syncthreads();
    for (int k = tid; k < 1024; k += 256) {
        const int it = (k & 31) + 32 * bx, jt = (k >> 5) + 32 * by;
        if (it > s.w - 2 || jt > s.h - 2 || it < 1 || jt < 1) continue;
        const float gf = sqrtf(s.t0[(size_t) jt * s.w + it].w);
        int g = (int) gf;
        if (g > 48) g = 48;
        if (g >= 0 && isfinite(gf)) { atomicAdd(&hist[g + 1], 1); atomicAdd(&hist[0], 1); }
    }
    __syncthreads();
    if (tid == 0) {
        int th = (int) lroundf((float) hist[0] * 0.5f), q = 90;
        for (int i = 0; i < 90; i++) { th -= (i + 1 < 50) ? hist[i + 1] : 0; if (th < 0) { q = i; break; } }
        s.ths[bx + by * s.w32] = (float) q + 7.f;
    }
}

__global__ void sel_smooth_kernel(const SelDev s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.w32 * s.h32) return;
    const int x = i % s.w32, y = i / s.w32, w32 = s.w32, h32 = s.h32;
    float sum = 0, num = 0;
    if (x > 0) {
        if (y > 0) { num++; sum += s.ths[x - 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += s.ths[x - 1 + (y + 1) * w32]; }
        num++; sum += s.ths[x - 1 + y * w32];
    }
    if (x < w32 - 1) {
        if (y > 0) { num++; sum += s.ths[x + 1 + (y - 1) * w32]; }
        if (y < h32 - 1) { num++; sum += s.ths[x + 1 + (y + 1) * w32]; }
        num++; sum += s.ths[x + 1 + y * w32];
    }
    if (y > 0) { num++; sum += s.ths[x + (y - 1) * w32]; }
    if (y < h32 - 1) { num++; sum += s.ths[x + (y + 1) * w32]; }
    num++; sum += s.ths[x + y * w32];
    s.ths_smoothed[i] = (sum / num) * (sum / num);
}

// One WARP per 4*pot block.  The sticky -2 flags of the reference's sweep reduce
// to: a pot block selects its best level-1 candidate; a 2*pot block selects a level-2 point only if none of its pot blocks selected; the
// 4*pot block selects a level-3 point only if nothing else was selected in it; "best" = largest |grad . dir| with the FIRST pixel in the
// reference's traversal order winning ties (strict > in the sequential code).  The lanes stride the pixels of a pot block, keep running
// bests with strict >, and the warp combines (value, traversal position) by shuffles.
struct Best { float v; int pos, idx; };
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, b.v, off);
        const int op = __shfl_xor_sync(0xffffffffu, b.pos, off), oi = __shfl_xor_sync(0xffffffffu, b.idx, off);
        if (oi >= 0 && (b.idx < 0 || ov > b.v || (ov == b.v && op < b.pos))) { b.v = ov; b.pos = op; b.idx = oi; }
    }
    return b;
}
__global__ void __launch_bounds__(128) sel_block_warp_kernel(const SelDev s, const int pot, const float thFactor, const int nbx, const int nblocks) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= nblocks) return;
    const int w = s.w, h = s.h;
    const int x4 = (b % nbx) * 4 * pot, y4 = (b / nbx) * 4 * pot;
    const float dw1 = 0.75f, dw2 = dw1 * dw1;
    int n2 = s.start[b], c2 = 0, c3 = 0, c4 = 0;
    const int my3 = min(4 * pot, h - y4), mx3 = min(4 * pot, w - x4);
    const int d4 = s.pattern[n2] & 0xF;
    Best b4{0.f, 0, -1};
    int seq = 0;
    for (int y3 = 0; y3 < my3; y3 += 2 * pot) for (int x3 = 0; x3 < mx3; x3 += 2 * pot) {
        const int x34 = x3 + x4, y34 = y3 + y4;
        const int my2 = min(2 * pot, h - y34), mx2 = min(2 * pot, w - x34);
        const int d3 = s.pattern[n2] & 0xF;
        Best b3{0.f, 0, -1};
        int c2_here = 0;
        for (int y2 = 0; y2 < my2; y2 += pot) for (int x2 = 0; x2 < mx2; x2 += pot, seq++) {
            const int x234 = x2 + x34, y234 = y2 + y34;
            const int my1 = min(pot, h - y234), mx1 = min(pot, w - x234);
            const int d2 = s.pattern[n2] & 0xF;
            Best b2{0.f, 0, -1};
            for (int k = lane; k < my1 * mx1; k += 32) {
                const int y1 = k / mx1, x1 = k - y1 * mx1;
                const int xf = x1 + x234, yf = y1 + y234, idx = xf + w * yf;
                if (xf < 4 || xf >= w - 5 || yf < 4 || yf > h - 4) continue;
                const float th0 = s.ths_smoothed[(xf >> 5) + (yf >> 5) * s.w32];
                const float th1 = th0 * dw1, th2 = th1 * dw2;
                const float4 t = s.t0[idx];
                const int pos = (seq << 16) | k;
                if (t.w > th0 * thFactor) {
                    const float dn = fabsf(__fadd_rn(__fmul_rn(t.y, c_dir[d2][0]), __fmul_rn(t.z, c_dir[d2][1])));
                    if (dn > b2.v) { b2.v = dn; b2.pos = pos; b2.idx = idx; }
                }
                const float ag1 = s.t1[(size_t) (int) ((float) yf * 0.5f + 0.25f) * s.w1 + (int) ((float) xf * 0.5f + 0.25f)].w;
                if (ag1 > th1 * thFactor) {
                    const float dn = fabsf(__fadd_rn(__fmul_rn(t.y, c_dir[d3][0]), __fmul_rn(t.z, c_dir[d3][1])));
                    if (dn > b3.v) { b3.v = dn; b3.pos = pos; b3.idx = idx; }
                }
                const float ag2 = s.t2[(size_t) (int) ((double) ((float) yf * 0.25f) + 0.125) * s.w2 + (int) ((double) ((float) xf * 0.25f) + 0.125)].w;
                if (ag2 > th2 * thFactor) {
                    const float dn = fabsf(__fadd_rn(__fmul_rn(t.y, c_dir[d4][0]), __fmul_rn(t.z, c_dir[d4][1])));
                    if (dn > b4.v) { b4.v = dn; b4.pos = pos; b4.idx = idx; }
                }
            }
            b2 = warp_best(b2);
            if (b2.idx > 0) { if (lane == 0) s.map[b2.idx] = 1.f; n2++; c2++; c2_here++; }
        }
        b3 = warp_best(b3);
        if (c2_here == 0 && b3.idx > 0) { if (lane == 0) s.map[b3.idx] = 2.f; c3++; }
    }
    b4 = warp_best(b4);
    if (c2 == 0 && c3 == 0 && b4.idx > 0) { if (lane == 0) s.map[b4.idx] = 4.f; c4++; }
    if (lane != 0) return;
    if (c2 != s.hits[b]) { s.hits[b] = c2; atomicAdd(s.counters + 3, 1); }
    if (c2) atomicAdd(s.counters + 0, c2);
    if (c3) atomicAdd(s.counters + 1, c3);
    if (c4) atomicAdd(s.counters + 2, c4);
}

// exclusive prefix sum, one CTA (n <= a few 100k)
__global__ void __launch_bounds__(1024) sel_scan_kernel(const int *__restrict__ in, int *__restrict__ out, const int n) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? in[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane], ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, ti, d); if (lane >= d) ti += u; }
            warp_tot[lane] = ti - t;
        }
        __syncthreads();
        const int excl = carry + warp_tot[wid] + inc - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
}

// raster compaction of the non-zero map entries: per-row counts, (scan), then emit; optional random sub-sampling (makeMaps :186-201):
// the rn-th non-zero entry in raster order dies when pattern[rn] > charTH
__global__ void __launch_bounds__(128) sel_rowcount_kernel(const SelDev s) {
    const int y = blockIdx.x, tid = threadIdx.x;
    int c = 0;
    for (int x = tid; x < s.w; x += 128) c += s.map[(size_t) y * s.w + x] != 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    __shared__ int part[4];
    if ((tid & 31) == 0) part[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) s.row_count[y] = part[0] + part[1] + part[2] + part[3];
}
template <bool kKill>
__global__ void __launch_bounds__(128) sel_rowemit_kernel(const SelDev s, const int charTH) {
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int s_warp[4];
    __shared__ int s_run;
    if (tid == 0) s_run = s.row_offset[y];
    __syncthreads();
    for (int x0 = 0; x0 < s.w; x0 += 128) {
        const int x = x0 + tid;
        const float v = x < s.w ? s.map[(size_t) y * s.w + x] : 0.f;
        const bool nz = v != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int rn = s_run + __popc(m & ((1u << lane) - 1u));
        for (int k = 0; k < wid; k++) rn += s_warp[k];
        if (nz) {
            if (kKill) { if ((int) s.pattern[rn] > charTH) s.map[(size_t) y * s.w + x] = 0.f; }
            else s.list[rn] = make_int2(y * s.w + x, (int) v);
        }
        __syncthreads();
        if (tid == 0) s_run += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        __syncthreads();
    }
    if (!kKill && tid == 0 && y == s.h - 1) *s.list_n = s_run;
}

static thread_local std::string g_create_error;

#define SCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return CMLSEL_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

struct Selector {
    int device = 0, w = 0, h = 0, w32 = 0, h32 = 0, potential = 3;
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<unsigned char> pattern;
    unsigned char *d_pattern = nullptr;
    float *d_ths = nullptr, *d_sm = nullptr, *d_map = nullptr;
    int *d_hits = nullptr, *d_start = nullptr, *d_counters = nullptr, *d_row_count = nullptr, *d_row_offset = nullptr, *d_list_n = nullptr;
    int2 *d_list = nullptr;
    int *h_pin = nullptr;
    size_t blocks_cap = 0;
    long launches = 0;

    ~Selector() {
        void *v[] = {d_pattern, d_ths, d_sm, d_map, d_hits, d_start, d_counters, d_row_count, d_row_offset, d_list_n, d_list};
        for (void *p : v) if (p) cudaFree(p);
        if (h_pin) cudaFreeHost(h_pin);
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }

    int create(int dev, int W, int H) {
        if (W < 64 || H < 64) { error = "image too small for the 32-pixel blocks of the selector"; return CMLSEL_ERR_ARG; }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) { error = "no CUDA device " + std::to_string(dev) + " (the pixel selector has no CPU path)"; return CMLSEL_ERR_CUDA; }
        device = dev; w = W; h = H; w32 = W / 32; h32 = H / 32;
        SCK(cudaSetDevice(dev));
        SCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SCK(cudaEventCreate(&ev0)); SCK(cudaEventCreate(&ev1));
        pattern.resize((size_t) W * H);
        uint32_t state = 777;      // PixelSelector::myRand
        for (size_t i = 0; i < pattern.size(); i++) { state = state * 1664525u + 1013904223u; pattern[i] = (unsigned char) (state >> 24); }
        const size_t px = (size_t) W * H, nth = (size_t) w32 * h32 + 100;
        SCK(cudaMalloc(&d_pattern, px)); SCK(cudaMemcpy(d_pattern, pattern.data(), px, cudaMemcpyHostToDevice));
        SCK(cudaMalloc(&d_ths, nth * 4)); SCK(cudaMalloc(&d_sm, nth * 4)); SCK(cudaMalloc(&d_map, px * 4));
        SCK(cudaMalloc(&d_counters, 32)); SCK(cudaMalloc(&d_row_count, (size_t) H * 4)); SCK(cudaMalloc(&d_row_offset, (size_t) H * 4)); SCK(cudaMalloc(&d_list_n, 4));
        SCK(cudaMalloc(&d_list, px * sizeof(int2)));
        blocks_cap = ((size_t) (W + 3) / 4) * ((size_t) (H + 3) / 4);       // pot = 1 is the finest grid
        SCK(cudaMalloc(&d_hits, blocks_cap * 4)); SCK(cudaMalloc(&d_start, blocks_cap * 4));
        SCK(cudaHostAlloc((void **) &h_pin, 64, cudaHostAllocDefault));
        return CMLSEL_OK;
    }

    SelDev dev_args(const void *const *tex) const {
        SelDev s{};
        s.w = w; s.h = h; s.w32 = w32; s.h32 = h32;
        s.t0 = (const float4 *) tex[0]; s.t1 = (const float4 *) tex[1]; s.t2 = (const float4 *) tex[2];
        s.w1 = w / 2; s.w2 = (w / 2) / 2;
        s.pattern = d_pattern; s.ths = d_ths; s.ths_smoothed = d_sm; s.map = d_map; s.hits = d_hits; s.start = d_start; s.counters = d_counters;
        s.row_count = d_row_count; s.row_offset = d_row_offset; s.list = d_list; s.list_n = d_list_n;
        return s;
    }

    // select(): iterate simulate -> scan to the fixed point; n[3] = (n2, n3, n4)
    int select(const SelDev &s, int pot, float thf, int n[3]) {
        const int nbx = (w + 4 * pot - 1) / (4 * pot), nby = (h + 4 * pot - 1) / (4 * pot), nb = nbx * nby;
        SCK(cudaMemsetAsync(d_hits, 0, (size_t) nb * 4, stream));
        SCK(cudaMemsetAsync(d_start, 0, (size_t) nb * 4, stream));
        auto pass = [&]() -> cudaError_t {
            cudaError_t e = cudaMemsetAsync(d_map, 0, (size_t) w * h * 4, stream);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_counters, 0, 32, stream);
            if (e != cudaSuccess) return e;
            sel_block_warp_kernel<<<(nb + 3) / 4, 128, 0, stream>>>(s, pot, thf, nbx, nb); launches++;
            return cudaGetLastError();
        };
        SCK(pass());                                                 // pass 1 (all starts 0) only produces the per-block counts: no read-back needed
        for (int it = 0; it < nb + 2; it++) {
            sel_scan_kernel<<<1, 1024, 0, stream>>>(d_hits, d_start, nb); launches++;
            SCK(pass());
            SCK(cudaMemcpyAsync(h_pin, d_counters, 16, cudaMemcpyDeviceToHost, stream));
            SCK(cudaStreamSynchronize(stream));
            n[0] = h_pin[0]; n[1] = h_pin[1]; n[2] = h_pin[2];
            if (h_pin[3] == 0) return CMLSEL_OK;                 // every block reproduced the count its start value was built from
        }
        error = "select() did not reach its fixed point";
        return CMLSEL_ERR_STATE;
    }

    int make_maps(const SelDev &s, float density, int recursions, float thf, int *num_out) {
        sel_hist_kernel<<<dim3(w32, h32), 256, 0, stream>>>(s); launches++;
        sel_smooth_kernel<<<(w32 * h32 + 127) / 128, 128, 0, stream>>>(s); launches++;
        int n[3];
        int rc = select(s, potential, thf, n);
        if (rc) return rc;
        const float numHave = (float) (n[0] + n[1] + n[2]), numWant = density;
        const float quotia = numWant / numHave;
        const float K = numHave * (float) ((potential + 1) * (potential + 1));
        int ideal = (int) sqrtf(K / numWant) - 1;
        if (ideal < 1) ideal = 1;
        if (recursions > 0 && quotia > 1.25f && potential > 1) {
            if (ideal >= potential) ideal = potential - 1;
            potential = ideal;
            return make_maps(s, density, recursions - 1, thf, num_out);
        } else if (recursions > 0 && quotia < 0.25f) {
            if (ideal <= potential) ideal = potential + 1;
            potential = ideal;
            return make_maps(s, density, recursions - 1, thf, num_out);
        }
        if (quotia < 0.95f) {
            const int charTH = (int) (unsigned char) (255.0f * quotia);
            sel_rowcount_kernel<<<h, 128, 0, stream>>>(s); launches++;
            sel_scan_kernel<<<1, 1024, 0, stream>>>(d_row_count, d_row_offset, h); launches++;
            sel_rowemit_kernel<true><<<h, 128, 0, stream>>>(s, charTH); launches++;
        }
        potential = ideal;
        *num_out = (int) numHave;
        return CMLSEL_OK;
    }

    int compute(const void *const *tex, float density, int recursions, float thf, int capacity, float *xy, float *types, int32_t *count, float *gpu_ms) {
        if (!tex || !tex[0] || !tex[1] || !tex[2] || !count || capacity < 0 || (capacity > 0 && (!xy || !types))) { error = "NULL argument"; return CMLSEL_ERR_ARG; }
        if (!(density > 0)) { error = "density must be positive"; return CMLSEL_ERR_ARG; }
        SCK(cudaSetDevice(device));
        const SelDev s = dev_args(tex);
        SCK(cudaMemsetAsync(d_ths, 0, ((size_t) w32 * h32 + 100) * 4, stream));
        SCK(cudaMemsetAsync(d_sm, 0, ((size_t) w32 * h32 + 100) * 4, stream));
        SCK(cudaEventRecord(ev0, stream));
        int have = 0;
        int rc = make_maps(s, density, recursions, thf, &have);
        if (rc) return rc;
        sel_rowcount_kernel<<<h, 128, 0, stream>>>(s); launches++;
        sel_scan_kernel<<<1, 1024, 0, stream>>>(d_row_count, d_row_offset, h); launches++;
        sel_rowemit_kernel<false><<<h, 128, 0, stream>>>(s, 0); launches++;
        SCK(cudaEventRecord(ev1, stream));
        SCK(cudaGetLastError());
        SCK(cudaMemcpyAsync(h_pin, d_list_n, 4, cudaMemcpyDeviceToHost, stream));
        SCK(cudaStreamSynchronize(stream));
        const int nsel = h_pin[0];
        std::vector<int2> list((size_t) nsel);
        if (nsel) SCK(cudaMemcpy(list.data(), d_list, (size_t) nsel * sizeof(int2), cudaMemcpyDeviceToHost));
        // compute(): 32-pixel border, derivative finite (a non-finite derivative makes the weighted norm non-finite, which never exceeds a threshold),
        // emission order x outer / y inner
        std::vector<int2> keep;
        keep.reserve(list.size());
        for (const int2 &e : list) { const int x = e.x % w, y = e.x / w; if (x >= 32 && x < w - 32 && y >= 32 && y < h - 32) keep.push_back(make_int2(x * h + y, e.y)); }
        std::sort(keep.begin(), keep.end(), [](const int2 &a, const int2 &b) { return a.x < b.x; });
        *count = (int32_t) keep.size();
        for (int i = 0; i < (int) keep.size() && i < capacity; i++) { xy[2 * i] = (float) (keep[i].x / h); xy[2 * i + 1] = (float) (keep[i].x % h); types[i] = (float) keep[i].y; }
        if (gpu_ms) SCK(cudaEventElapsedTime(gpu_ms, ev0, ev1));
        return CMLSEL_OK;
    }

    int64_t read(const char *name, void *dst, int64_t cap) {
        const std::string n(name);
        const void *src = nullptr; size_t bytes = 0;
        if (n == "ths") { src = d_ths; bytes = (size_t) w32 * h32 * 4; }
        else if (n == "ths_smoothed") { src = d_sm; bytes = (size_t) w32 * h32 * 4; }
        else if (n == "map") { src = d_map; bytes = (size_t) w * h * 4; }
        else { error = "unknown buffer " + n; return CMLSEL_ERR_ARG; }
        if ((int64_t) bytes > cap) { error = "buffer too small"; return CMLSEL_ERR_ARG; }
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { error = "device read failed"; return CMLSEL_ERR_CUDA; }
        return (int64_t) bytes;
    }
};

}  // namespace cmlsel

using cmlsel::Selector;
#define SH(h) reinterpret_cast<Selector *>(h)

extern "C" {

int cmlsel_create(int device, int width, int height, cmlsel_handle *out) {
    if (!out) { cmlsel::g_create_error = "out is NULL"; return CMLSEL_ERR_ARG; }
    *out = nullptr;
    Selector *s = new Selector();
    const int rc = s->create(device, width, height);
    if (rc) { cmlsel::g_create_error = s->error; delete s; return rc; }
    *out = reinterpret_cast<cmlsel_handle>(s);
    return CMLSEL_OK;
}
void cmlsel_destroy(cmlsel_handle h) { delete SH(h); }
const char *cmlsel_last_error(cmlsel_handle h) { return h ? SH(h)->error.c_str() : cmlsel::g_create_error.c_str(); }
int cmlsel_set_potential(cmlsel_handle h, int potential) { if (!h || potential < 1) return CMLSEL_ERR_ARG; SH(h)->potential = potential; return CMLSEL_OK; }
int cmlsel_get_potential(cmlsel_handle h) { return h ? SH(h)->potential : CMLSEL_ERR_ARG; }
int cmlsel_compute(cmlsel_handle h, const void *const *d_texels, float density, int recursions_left, float th_factor, int capacity, float *corners_xy, float *types, int32_t *count,
                   float *gpu_ms) {
    return h ? SH(h)->compute(d_texels, density, recursions_left, th_factor, capacity, corners_xy, types, count, gpu_ms) : CMLSEL_ERR_ARG;
}
int64_t cmlsel_read(cmlsel_handle h, const char *name, void *dst, int64_t capacity) { return (h && name && dst) ? SH(h)->read(name, dst, capacity) : CMLSEL_ERR_ARG; }

}  // extern "C"
