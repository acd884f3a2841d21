// fast.cu -- FAST-9 corner detector on the device and the C ABI of include/cmlfast.h (SURVEY.md 8f NEXT #4, first unit of the ORB extractor).
//
// Reference anchors: /root/reference/src/cml/features/corner/FAST.cpp
//   fast_score_kernel   fast9_detect :2985-5913 and fast9_corner_score :15-2944 (generated decision trees of one predicate: 9 contiguous circle
//                       pixels all > p + b or all < p - b; the score is the bisection's largest b in [threshold, 255] that still detects)
//   fast_nms_* kernels  nonmax_suppression :5921-6033 (a corner survives iff no 8-neighbour corner has a score >= its own), raster order
//
// B200 design: the reference walks a 3 000-line decision tree per pixel and a pointer-chasing row scan for the suppression.  Here every pixel
// builds the two 16-bit "brighter" / "darker" ring masks and tests "9 contiguous bits" with 8 shifts-and-ands on the doubled mask; the score map
// makes the suppression a 3x3 stencil, and the survivors are compacted in raster order by row counts + scan + row emit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmlfast.h"

namespace cmlfast {

__device__ __forceinline__ bool arc9(unsigned m) {          // m: 16 ring bits; true if 9 circularly contiguous bits are set
    m |= m << 16;
    unsigned t = m & (m >> 1);
    t &= t >> 2;            // runs of 4
    t &= t >> 4;            // runs of 8
    t &= m >> 8;            // runs of 9
    return (t & 0xffffu) != 0u;
}

__device__ __forceinline__ bool is_corner(const int (&ring)[16], const int c, const int b) {
    unsigned br = 0, dk = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) { br |= (unsigned) (ring[k] > c + b) << k; dk |= (unsigned) (ring[k] < c - b) << k; }
    return arc9(br) || arc9(dk);
}

// score map: -1 = not a corner at `threshold`, else the bisection result of fast9_corner_score
__global__ void __launch_bounds__(256) fast_score_kernel(const uint8_t *__restrict__ img, int *__restrict__ score, const int w, const int h, const int threshold) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    int s = -1;
    if (x >= 3 && y >= 3 && x < w - 3 && y < h - 3) {
        const uint8_t *p = img + (size_t) y * w + x;
        const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}, dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};     // make_offsets
        int ring[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ring[k] = p[dy[k] * w + dx[k]];
        const int c = *p;
        if (is_corner(ring, c, threshold)) {
            int bmin = threshold, bmax = 255, b = (bmax + bmin) / 2;
            for (;;) {
                if (is_corner(ring, c, b)) bmin = b; else bmax = b;
                if (bmin == bmax - 1 || bmin == bmax) break;
                b = (bmin + bmax) / 2;
            }
            s = bmin;
        }
    }
    score[(size_t) y * w + x] = s;
}

__device__ __forceinline__ bool survives(const int *__restrict__ score, const int w, const int h, const int x, const int y) {
    const int s = score[(size_t) y * w + x];
    if (s < 0) return false;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            const int xx = x + dx, yy = y + dy;
            if (xx >= 0 && yy >= 0 && xx < w && yy < h && score[(size_t) yy * w + xx] >= s) return false;     // non-corners hold -1
        }
    return true;
}

__global__ void __launch_bounds__(128) fast_rowcount_kernel(const int *__restrict__ score, int *__restrict__ row_count, const int w, const int h) {
    const int y = blockIdx.x, tid = threadIdx.x;
    int c = 0;
    for (int x = tid; x < w; x += 128) c += survives(score, w, h, x, y);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    __shared__ int part[4];
    if ((tid & 31) == 0) part[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) row_count[y] = part[0] + part[1] + part[2] + part[3];
}

__global__ void __launch_bounds__(1024) fast_scan_kernel(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ total, const int n) {
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
    if (tid == 0) *total = carry;
}

__global__ void __launch_bounds__(128) fast_rowemit_kernel(const int *__restrict__ score, const int *__restrict__ row_offset, int4 *__restrict__ list, const int w, const int h,
                                                          const int capacity) {
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int s_warp[4];
    __shared__ int s_run;
    if (tid == 0) s_run = row_offset[y];
    __syncthreads();
    for (int x0 = 0; x0 < w; x0 += 128) {
        const int x = x0 + tid;
        const bool keep = x < w && survives(score, w, h, x, y);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[wid] = __popc(m);
        __syncthreads();
        int k = s_run + __popc(m & ((1u << lane) - 1u));
        for (int q = 0; q < wid; q++) k += s_warp[q];
        if (keep && k < capacity) list[k] = make_int4(x, y, score[(size_t) y * w + x], 0);
        __syncthreads();
        if (tid == 0) s_run += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        __syncthreads();
    }
}

static thread_local std::string g_create_error;

#define FCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return CMLFAST_ERR_CUDA;                                                               \
        }                                                                                          \
    } while (0)

struct Fast {
    int device = 0, maxw = 0, maxh = 0;
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *d_img = nullptr, *h_img = nullptr;
    int *d_score = nullptr, *d_row_count = nullptr, *d_row_offset = nullptr, *d_total = nullptr, *h_total = nullptr;
    int4 *d_list = nullptr;
    size_t list_cap = 0;

    ~Fast() {
        void *v[] = {d_img, d_score, d_row_count, d_row_offset, d_total, d_list};
        for (void *p : v) if (p) cudaFree(p);
        if (h_img) cudaFreeHost(h_img); if (h_total) cudaFreeHost(h_total);
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    int create(int dev, int W, int H) {
        if (W < 7 || H < 7) { error = "image smaller than the 7x7 FAST footprint"; return CMLFAST_ERR_ARG; }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) { error = "no CUDA device " + std::to_string(dev) + " (the FAST detector has no CPU path)"; return CMLFAST_ERR_CUDA; }
        device = dev; maxw = W; maxh = H;
        FCK(cudaSetDevice(dev));
        FCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        FCK(cudaEventCreate(&ev0)); FCK(cudaEventCreate(&ev1));
        const size_t px = (size_t) W * H;
        FCK(cudaMalloc(&d_img, px)); FCK(cudaHostAlloc((void **) &h_img, px, cudaHostAllocDefault));
        FCK(cudaMalloc(&d_score, px * 4)); FCK(cudaMalloc(&d_row_count, (size_t) H * 4)); FCK(cudaMalloc(&d_row_offset, (size_t) H * 4)); FCK(cudaMalloc(&d_total, 4));
        FCK(cudaHostAlloc((void **) &h_total, 4, cudaHostAllocDefault));
        return CMLFAST_OK;
    }
    int compute(const uint8_t *img, int w, int h, int th, int capacity, int32_t *xy, int32_t *scores, int32_t *count, float *gpu_ms) {
        if (!img || !count || w < 7 || h < 7 || w > maxw || h > maxh || (size_t) w * h > (size_t) maxw * maxh || capacity < 0 || (capacity > 0 && (!xy || !scores))) {
            error = "bad arguments (NULL pointer or image larger than the handle's)"; return CMLFAST_ERR_ARG;
        }
        FCK(cudaSetDevice(device));
        const size_t px = (size_t) w * h;
        if ((size_t) capacity > list_cap) {
            if (d_list) cudaFree(d_list);
            d_list = nullptr; list_cap = 0;
            FCK(cudaMalloc(&d_list, (size_t) capacity * sizeof(int4)));
            list_cap = (size_t) capacity;
        }
        memcpy(h_img, img, px);
        FCK(cudaMemcpyAsync(d_img, h_img, px, cudaMemcpyHostToDevice, stream));
        FCK(cudaEventRecord(ev0, stream));
        fast_score_kernel<<<dim3((w + 31) / 32, (h + 7) / 8), 256, 0, stream>>>(d_img, d_score, w, h, th);
        fast_rowcount_kernel<<<h, 128, 0, stream>>>(d_score, d_row_count, w, h);
        fast_scan_kernel<<<1, 1024, 0, stream>>>(d_row_count, d_row_offset, d_total, h);
        if (capacity > 0) fast_rowemit_kernel<<<h, 128, 0, stream>>>(d_score, d_row_offset, d_list, w, h, capacity);
        FCK(cudaEventRecord(ev1, stream));
        FCK(cudaGetLastError());
        FCK(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, stream));
        FCK(cudaStreamSynchronize(stream));
        const int n = *h_total, m = std::min(n, capacity);
        *count = n;
        if (m > 0) {
            std::vector<int4> list((size_t) m);
            FCK(cudaMemcpy(list.data(), d_list, (size_t) m * sizeof(int4), cudaMemcpyDeviceToHost));
            for (int i = 0; i < m; i++) { xy[2 * i] = list[i].x; xy[2 * i + 1] = list[i].y; scores[i] = list[i].z; }
        }
        if (gpu_ms) FCK(cudaEventElapsedTime(gpu_ms, ev0, ev1));
        return CMLFAST_OK;
    }
};

}  // namespace cmlfast

using cmlfast::Fast;

extern "C" {

int cmlfast_create(int device, int max_width, int max_height, cmlfast_handle *out) {
    if (!out) { cmlfast::g_create_error = "out is NULL"; return CMLFAST_ERR_ARG; }
    *out = nullptr;
    Fast *f = new Fast();
    const int rc = f->create(device, max_width, max_height);
    if (rc) { cmlfast::g_create_error = f->error; delete f; return rc; }
    *out = reinterpret_cast<cmlfast_handle>(f);
    return CMLFAST_OK;
}
void cmlfast_destroy(cmlfast_handle h) { delete reinterpret_cast<Fast *>(h); }
const char *cmlfast_last_error(cmlfast_handle h) { return h ? reinterpret_cast<Fast *>(h)->error.c_str() : cmlfast::g_create_error.c_str(); }
int cmlfast_compute(cmlfast_handle h, const uint8_t *image, int width, int height, int threshold, int capacity, int32_t *xy, int32_t *scores, int32_t *count, float *gpu_ms) {
    return h ? reinterpret_cast<Fast *>(h)->compute(image, width, height, threshold, capacity, xy, scores, count, gpu_ms) : CMLFAST_ERR_ARG;
}

}  // extern "C"
