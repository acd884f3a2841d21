// imgprep.cu -- image preparation on the device and the C ABI of include/cmlimg.h (SURVEY.md 8f NEXT #3).
//
// Reference anchors (under /root/reference/src/cml):
//   prep_level0_kernel   capture/CaptureImage.cpp:137-202 (LUT, inverse vignette, removeDistortion), image/LookupTable.h:99-104,
//                        map/InternalCalibration.h:404-437, image/Array2D.h:242-263 (interpolate); pyramid :228-236 + Array2D.h:388-401
//   prep_texel_kernel    image/Array2D.h:288-331 (gradientImage), image/Array2DProxy.h:198-226 (WeightedGradientImageProxy)
//   Prep::set_photometric image/LookupTable.h:112-131 (computeInverse)
//
// B200 design: the reference makes four full-image passes on the host (LUT, vignette, remap, then per level reduce / gradient / weight).
// Here the photometric correction is applied per TAP inside the remap gather (no corrected intermediate image exists), the same CTA
// reduces its 32x32 tile to every coarser level in shared memory, and one second kernel writes all levels' texels.  Streaming,
// HBM-bound: ~43 algorithmic bytes per rectified pixel (raw 4 + vignette 4 + map 8 + gray levels 5.3 + texels 21.3).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmlimg.h"

namespace cmlimg {

constexpr int MAXL = CMLIMG_MAX_LEVELS;
constexpr int TILE = 32;

struct PrepDev {
    int levels, in_w, in_h;
    int w[MAXL], h[MAXL];
    const void *raw;               // float or uint8 sensor image (template parameter of prep_level0_kernel)
    const float *vig;              // may be null
    const float2 *map;             // may be null (identity)
    const float *lut;              // 256 values, null = identity
    const float *inv;              // 256 inverse values
    float *gray[MAXL];
    float4 *texel[MAXL];
};

// GrayLookupTable::operator()(float): uint8 truncation, uint8 wrap of the upper index
__device__ __forceinline__ float lut_apply(const float *s_lut, const float v) {
    const unsigned i0 = (unsigned) (int) v & 255u, i1 = (i0 + 1u) & 255u;
    const float f = v - (float) i0;
    return __fadd_rn(__fmul_rn(s_lut[i0], 1.0f - f), __fmul_rn(s_lut[i1], f));
}

template <typename TIn>
__global__ void __launch_bounds__(256) prep_level0_kernel(const PrepDev p) {
    __shared__ float s_lut[256];
    __shared__ float buf[2][TILE][TILE + 1];
    const int tid = threadIdx.x;
    const bool has_lut = p.lut != nullptr;
    if (has_lut) s_lut[tid] = p.lut[tid];
    __syncthreads();
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    // 4 rectified pixels per thread, phase by phase (map entries, then all 32 tap loads, then arithmetic) so that the loads of the four
    // pixels are in flight together instead of one dependent chain per pixel
    constexpr int PPT = TILE * TILE / 256;
    const TIn *__restrict__ raw = static_cast<const TIn *>(p.raw);
    const float *__restrict__ vig = p.vig;
    float2 m[PPT];
    bool live[PPT], fin[PPT];
#pragma unroll
    for (int q = 0; q < PPT; q++) {
        const int k = tid + 256 * q, x = k % TILE, y = k / TILE, gx = tx0 + x, gy = ty0 + y;
        live[q] = gx < p.w[0] && gy < p.h[0];
        m[q] = make_float2((float) gx, (float) gy);
        if (live[q] && p.map) m[q] = p.map[(size_t) gy * p.w[0] + gx];
        fin[q] = live[q] && isfinite(m[q].x);
    }
    float r[PPT][4], g[PPT][4];
#pragma unroll
    for (int q = 0; q < PPT; q++) {
        const int ix = fin[q] ? (int) m[q].x : 0, iy = fin[q] ? (int) m[q].y : 0;
        const size_t i0 = (size_t) iy * p.in_w + ix;
        const bool four = fin[q] && p.map != nullptr;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const size_t i = i0 + (t & 1) + (size_t) (t >> 1) * p.in_w;
            const bool need = t == 0 ? fin[q] : four;
            r[q][t] = need ? (float) raw[i] : 0.f;
            g[q][t] = (need && vig) ? vig[i] : 1.f;
        }
    }
#pragma unroll
    for (int q = 0; q < PPT; q++) {
        const int k = tid + 256 * q, x = k % TILE, y = k / TILE, gx = tx0 + x, gy = ty0 + y;
        float t4[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            float v = r[q][t];
            if (has_lut) v = lut_apply(s_lut, v);
            if (vig) v = __fmul_rn(v, g[q][t]);
            t4[t] = v;
        }
        float v = 0.f;
        if (fin[q]) {
            if (p.map) {
                const int ix = (int) m[q].x, iy = (int) m[q].y;
                const float dx = m[q].x - (float) ix, dy = m[q].y - (float) iy, dxdy = dx * dy;
                v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t4[0], 1.f - dx - dy + dxdy), __fmul_rn(t4[1], dx - dxdy)), __fmul_rn(t4[2], dy - dxdy)), __fmul_rn(t4[3], dxdy));
            } else v = t4[0];
        }
        if (live[q]) p.gray[0][(size_t) gy * p.w[0] + gx] = v;
        buf[0][y][x] = v;
    }
    __syncthreads();
    int size = TILE;
    for (int l = 1; l < p.levels && size > 1; l++) {       // levels past the tile's depth (l >= 6) are finished by prep_tail_kernel
        size >>= 1;
        const int ox = tx0 >> l, oy = ty0 >> l;
        float (*src)[TILE + 1] = buf[(l - 1) & 1];
        float (*dst)[TILE + 1] = buf[l & 1];
        for (int k = tid; k < size * size; k += 256) {
            const int x = k % size, y = k / size;
            const float v = (((src[2 * y][2 * x] + src[2 * y][2 * x + 1]) + src[2 * y + 1][2 * x]) + src[2 * y + 1][2 * x + 1]) / 4.f;
            dst[y][x] = v;
            if (ox + x < p.w[l] && oy + y < p.h[l]) p.gray[l][(size_t) (oy + y) * p.w[l] + ox + x] = v;
        }
        __syncthreads();
    }
}

// levels deeper than a 32x32 tile reaches (>= 6): plain 2x2 means from the level above, one small launch per level
__global__ void __launch_bounds__(256) prep_tail_kernel(const PrepDev p, const int l) {
    const int w = p.w[l], h = p.h[l], wm = p.w[l - 1];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < w * h; i += gridDim.x * 256) {
        const int x = i % w, y = i / w;
        const float *s = p.gray[l - 1] + (size_t) 2 * y * wm + 2 * x;
        p.gray[l][i] = (((s[0] + s[1]) + s[wm]) + s[wm + 1]) / 4.f;
    }
}

// texel (I, dx, dy, weighted gradient norm) of every level; border ring (0, 0, 0, 0) like gradientImage
__global__ void __launch_bounds__(256) prep_texel_kernel(const PrepDev p) {
    __shared__ float s_inv[256];
    s_inv[threadIdx.x] = p.inv[threadIdx.x];
    __syncthreads();
    const int l = blockIdx.y;
    const int w = p.w[l], h = p.h[l];
    const float *__restrict__ g = p.gray[l];
    for (int y = blockIdx.x; y < h; y += gridDim.x) {           // a row per CTA step: no integer division per pixel
        const bool yin = y > 0 && y < h - 1;
        const float *__restrict__ row = g + (size_t) y * w;
        float4 *__restrict__ orow = p.texel[l] + (size_t) y * w;
        for (int x = threadIdx.x; x < w; x += 256) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yin && x > 0 && x < w - 1) {
                o.x = row[x];
                o.y = (row[x + 1] - row[x - 1]) * 0.5f;
                o.z = (row[x + w] - row[x - w]) * 0.5f;
            }
            int c = (int) lroundf(o.x);
            c = c < 5 ? 5 : (c > 250 ? 250 : c);
            const float gw = s_inv[c + 1] - s_inv[c];
            o.w = __fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(o.y, o.y), __fmul_rn(o.z, o.z)), gw), gw);
            orow[x] = o;
        }
    }
}

__global__ void __launch_bounds__(256) flush_kernel(float4 *buf, const size_t n) {
    for (size_t i = (size_t) blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t) gridDim.x * 256) buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

static thread_local std::string g_create_error;

#define ICK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return CMLIMG_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

struct Prep {
    int device = 0, in_w = 0, in_h = 0, L = 0;
    int w[MAXL]{}, h[MAXL]{};
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float *d_raw = nullptr, *d_vig = nullptr, *d_lut = nullptr, *d_inv = nullptr, *h_raw = nullptr;
    float2 *d_map = nullptr;
    char *d_block = nullptr;
    float *gray[MAXL]{}; float4 *texel[MAXL]{};
    float4 *d_flush = nullptr;
    bool has_lut = false, has_vig = false, has_map = false, prepared = false, raw_u8 = false;

    ~Prep() {
        void *v[] = {d_raw, d_vig, d_lut, d_inv, d_map, d_block, d_flush};
        for (void *p : v) if (p) cudaFree(p);
        if (h_raw) cudaFreeHost(h_raw);
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }

    int create(int dev, int iw, int ih, int ow, int oh, int levels) {
        if (iw < 16 || ih < 16 || ow < 16 || oh < 16) { error = "bad image size"; return CMLIMG_ERR_ARG; }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) { error = "no CUDA device " + std::to_string(dev) + " (the image preparation has no CPU path)"; return CMLIMG_ERR_CUDA; }
        device = dev; in_w = iw; in_h = ih;
        ICK(cudaSetDevice(dev));
        if (levels <= 0) {         // CaptureImage.cpp:39-72
            levels = 0;
            double sx = ow, sy = oh;
            for (;;) { if (sx * sy <= 625.0 && levels >= 5) break; levels++; sx /= 2; sy /= 2; if (levels >= 32) break; }
        }
        L = std::min(levels, MAXL);
        for (int l = 0; l < L; l++) { w[l] = l ? w[l - 1] / 2 : ow; h[l] = l ? h[l - 1] / 2 : oh; if (w[l] < 1 || h[l] < 1) { L = l; break; } }
        ICK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ICK(cudaEventCreate(&ev0)); ICK(cudaEventCreate(&ev1));
        const size_t ipx = (size_t) iw * ih;
        ICK(cudaMalloc(&d_raw, ipx * 4)); ICK(cudaMalloc(&d_vig, ipx * 4)); ICK(cudaMalloc(&d_lut, 1024)); ICK(cudaMalloc(&d_inv, 1024));
        ICK(cudaMalloc(&d_map, (size_t) ow * oh * 8));
        ICK(cudaHostAlloc((void **) &h_raw, ipx * 4, cudaHostAllocDefault));
        size_t off = 0, og[MAXL], ot[MAXL];
        for (int l = 0; l < L; l++) { og[l] = off; off += (((size_t) w[l] * h[l] * 4) + 255) & ~(size_t) 255; ot[l] = off; off += (((size_t) w[l] * h[l] * 16) + 255) & ~(size_t) 255; }
        ICK(cudaMalloc(&d_block, off));
        ICK(cudaMemset(d_block, 0, off));
        for (int l = 0; l < L; l++) { gray[l] = (float *) (d_block + og[l]); texel[l] = (float4 *) (d_block + ot[l]); }
        float ident[256];
        for (int i = 0; i < 256; i++) ident[i] = (float) i;
        ICK(cudaMemcpy(d_inv, ident, 1024, cudaMemcpyHostToDevice));
        return CMLIMG_OK;
    }

    int set_photometric(const float *lut, const float *vig) {
        ICK(cudaSetDevice(device));
        ICK(cudaStreamSynchronize(stream));
        float inv[256];
        for (int i = 0; i < 256; i++) inv[i] = (float) i;
        if (lut) {                  // GrayLookupTable::computeInverse (LookupTable.h:112-131), fp32 like the reference
            for (int i = 1; i < 255; i++)
                for (int s = 1; s < 255; s++)
                    if (lut[s] <= i && lut[s + 1] >= i) { inv[i] = s + (i - lut[s]) / (lut[s + 1] - lut[s]); break; }
            inv[0] = 0; inv[255] = 255;
            ICK(cudaMemcpy(d_lut, lut, 1024, cudaMemcpyHostToDevice));
        }
        ICK(cudaMemcpy(d_inv, inv, 1024, cudaMemcpyHostToDevice));
        if (vig) ICK(cudaMemcpy(d_vig, vig, (size_t) in_w * in_h * 4, cudaMemcpyHostToDevice));
        has_lut = lut != nullptr; has_vig = vig != nullptr;
        return CMLIMG_OK;
    }
    int set_map(const float *map) {
        ICK(cudaSetDevice(device));
        ICK(cudaStreamSynchronize(stream));
        if (!map && (in_w != w[0] || in_h != h[0])) { error = "without an undistortion map the input and output sizes must agree"; return CMLIMG_ERR_ARG; }
        if (map) {
            // the bilinear taps of every finite entry must lie inside the input (the reference's map builder guarantees x < w - 1, y < h - 1)
            const size_t n = (size_t) w[0] * h[0];
            for (size_t i = 0; i < n; i++) {
                const float x = map[2 * i], y = map[2 * i + 1];
                if (std::isfinite(x) && !(x >= 0 && y >= 0 && x < in_w - 1 && y < in_h - 1)) { error = "undistortion map points outside the input image"; return CMLIMG_ERR_ARG; }
            }
            ICK(cudaMemcpy(d_map, map, n * 8, cudaMemcpyHostToDevice));
        }
        has_map = map != nullptr;
        return CMLIMG_OK;
    }

    PrepDev dev_args() const {
        PrepDev p{};
        p.levels = L; p.in_w = in_w; p.in_h = in_h;
        for (int l = 0; l < L; l++) { p.w[l] = w[l]; p.h[l] = h[l]; p.gray[l] = gray[l]; p.texel[l] = texel[l]; }
        p.raw = d_raw; p.vig = has_vig ? d_vig : nullptr; p.map = has_map ? d_map : nullptr; p.lut = has_lut ? d_lut : nullptr; p.inv = d_inv;
        return p;
    }
    void launch() {
        const PrepDev p = dev_args();
        const dim3 tiles((w[0] + TILE - 1) / TILE, (h[0] + TILE - 1) / TILE);
        if (raw_u8) prep_level0_kernel<uint8_t><<<tiles, 256, 0, stream>>>(p);
        else prep_level0_kernel<float><<<tiles, 256, 0, stream>>>(p);
        for (int l = 6; l < L; l++) prep_tail_kernel<<<std::max(1, (w[l] * h[l] + 255) / 256), 256, 0, stream>>>(p, l);
        prep_texel_kernel<<<dim3(std::min(2368, h[0]), L), 256, 0, stream>>>(p);
    }

    int prepare(const void *raw, bool u8, float *gpu_ms) {
        if (!raw) { error = "raw is NULL"; return CMLIMG_ERR_ARG; }
        if (!has_map && (in_w != w[0] || in_h != h[0])) { error = "no undistortion map set"; return CMLIMG_ERR_STATE; }
        ICK(cudaSetDevice(device));
        const size_t bytes = (size_t) in_w * in_h * (u8 ? 1 : 4);
        raw_u8 = u8;
        if (raw != (const void *) h_raw) { ICK(cudaStreamSynchronize(stream)); memcpy(h_raw, raw, bytes); }
        ICK(cudaMemcpyAsync(d_raw, h_raw, bytes, cudaMemcpyHostToDevice, stream));
        ICK(cudaEventRecord(ev0, stream));
        launch();
        ICK(cudaEventRecord(ev1, stream));
        ICK(cudaGetLastError());
        ICK(cudaStreamSynchronize(stream));
        if (gpu_ms) ICK(cudaEventElapsedTime(gpu_ms, ev0, ev1));
        prepared = true;
        return CMLIMG_OK;
    }

    int bench(int repeats, int flush, float *ms_out) {
        if (!prepared) { error = "cmlimg_prepare first"; return CMLIMG_ERR_STATE; }
        ICK(cudaSetDevice(device));
        const size_t fb = (size_t) 384 << 20;
        if (flush && !d_flush) ICK(cudaMalloc(&d_flush, fb));
        double tot = 0;
        for (int i = 0; i < repeats; i++) {
            if (flush) flush_kernel<<<1184, 256, 0, stream>>>(d_flush, fb / 16);
            ICK(cudaEventRecord(ev0, stream));
            launch();
            ICK(cudaEventRecord(ev1, stream));
            ICK(cudaStreamSynchronize(stream));
            float ms = 0; ICK(cudaEventElapsedTime(&ms, ev0, ev1)); tot += ms;
        }
        ICK(cudaGetLastError());
        *ms_out = (float) (tot / std::max(repeats, 1));
        return CMLIMG_OK;
    }

    const void *ptr(const std::string &n, size_t *bytes) {
        if (n.size() == 5 && n.compare(0, 4, "gray") == 0 && isdigit(n[4]) && n[4] - '0' < L) { const int l = n[4] - '0'; if (bytes) *bytes = (size_t) w[l] * h[l] * 4; return gray[l]; }
        if (n.size() == 6 && n.compare(0, 5, "texel") == 0 && isdigit(n[5]) && n[5] - '0' < L) { const int l = n[5] - '0'; if (bytes) *bytes = (size_t) w[l] * h[l] * 16; return texel[l]; }
        return nullptr;
    }
};

}  // namespace cmlimg

using cmlimg::Prep;
#define IH(h) reinterpret_cast<Prep *>(h)

extern "C" {

int cmlimg_create(int device, int in_width, int in_height, int out_width, int out_height, int levels, cmlimg_handle *out) {
    if (!out) { cmlimg::g_create_error = "out is NULL"; return CMLIMG_ERR_ARG; }
    *out = nullptr;
    Prep *p = new Prep();
    const int rc = p->create(device, in_width, in_height, out_width, out_height, levels);
    if (rc) { cmlimg::g_create_error = p->error; delete p; return rc; }
    *out = reinterpret_cast<cmlimg_handle>(p);
    return CMLIMG_OK;
}
void cmlimg_destroy(cmlimg_handle h) { delete IH(h); }
const char *cmlimg_last_error(cmlimg_handle h) { return h ? IH(h)->error.c_str() : cmlimg::g_create_error.c_str(); }
int cmlimg_set_photometric(cmlimg_handle h, const float *lut, const float *inv_vignette) { return h ? IH(h)->set_photometric(lut, inv_vignette) : CMLIMG_ERR_ARG; }
int cmlimg_set_undistort_map(cmlimg_handle h, const float *map) { return h ? IH(h)->set_map(map) : CMLIMG_ERR_ARG; }
float *cmlimg_input_buffer(cmlimg_handle h) { return h ? IH(h)->h_raw : nullptr; }
int cmlimg_prepare(cmlimg_handle h, const float *raw, float *gpu_ms) { return h ? IH(h)->prepare(raw, false, gpu_ms) : CMLIMG_ERR_ARG; }
int cmlimg_prepare_u8(cmlimg_handle h, const uint8_t *raw, float *gpu_ms) { return h ? IH(h)->prepare(raw, true, gpu_ms) : CMLIMG_ERR_ARG; }
int cmlimg_levels(cmlimg_handle h, int32_t *num_levels, int32_t *wh) {
    if (!h || !num_levels) return CMLIMG_ERR_ARG;
    *num_levels = IH(h)->L;
    if (wh) for (int l = 0; l < IH(h)->L; l++) { wh[2 * l] = IH(h)->w[l]; wh[2 * l + 1] = IH(h)->h[l]; }
    return CMLIMG_OK;
}
int64_t cmlimg_read(cmlimg_handle h, const char *name, void *dst, int64_t capacity) {
    if (!h || !name || !dst) return CMLIMG_ERR_ARG;
    Prep *p = IH(h);
    size_t bytes = 0;
    const void *src = p->ptr(name, &bytes);
    if (!src) { p->error = std::string("unknown buffer ") + name; return CMLIMG_ERR_ARG; }
    if ((int64_t) bytes > capacity) { p->error = "buffer too small"; return CMLIMG_ERR_ARG; }
    if (cudaSetDevice(p->device) != cudaSuccess || cudaStreamSynchronize(p->stream) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        p->error = "device read failed"; return CMLIMG_ERR_CUDA;
    }
    return (int64_t) bytes;
}
const void *cmlimg_device_ptr(cmlimg_handle h, const char *name) { return (h && name) ? IH(h)->ptr(name, nullptr) : nullptr; }
int cmlimg_bench(cmlimg_handle h, int repeats, int flush_l2, float *ms_per_frame) { return (h && ms_per_frame && repeats > 0) ? IH(h)->bench(repeats, flush_l2, ms_per_frame) : CMLIMG_ERR_ARG; }

}  // extern "C"
