// Development microbenchmark (not product code): throughput of TMA box loads of float4 texel tiles for several tensor-map
// shapes, ring depths and box sizes.  One producer thread per CTA (148 CTAs) streams `per_cta` boxes through `depth` smem slots.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_bench tools/tma_bench.cu -lcudart
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, uint32_t parity) {
    for (int i = 0; i < (1 << 22); i++) if (mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void tma2(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma3(void *dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void bulk1(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Params {
    int kind;        // 0: 2-D tensor map, 1: 3-D chunked tensor map, 2: 1-D bulk copies per row
    int box_w, box_h;   // texels
    int depth, per_cta, splits;   // splits: the box is fetched as `splits` row bands (separate TMA ops on the same barrier)
    int W, H, tiles_x, tiles_y;
    const float4 *img;
};

__global__ void __launch_bounds__(32, 1) stream_kernel(const __grid_constant__ CUtensorMap map, const Params p, int *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned long long full[8];
    if (threadIdx.x != 0) return;
    const int tile_bytes = p.box_w * p.box_h * 16;
    for (int s = 0; s < p.depth; s++) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    const int ntiles = p.tiles_x * p.tiles_y;
    int bad = 0;
    for (int i = 0; i < p.per_cta + p.depth; i++) {
        const int s = i % p.depth, u = i / p.depth;
        if (u > 0) bad |= !mbar_wait(full + s, (u - 1) & 1);        // the previous use of the slot has landed
        if (i >= p.per_cta) continue;
        const int tile = (blockIdx.x * p.per_cta + i) % ntiles, tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int x0 = tx * 64, y0 = ty * 32;
        unsigned char *dst = smem + (size_t) s * tile_bytes;
        mbar_expect_tx(full + s, tile_bytes);
        if (p.kind == 0) {
            const int band = p.box_h / p.splits;
            for (int b = 0; b < p.splits; b++) tma2(dst + (size_t) b * band * p.box_w * 16, &map, x0 * 2, y0 + b * band, full + s);
        } else if (p.kind == 1) {
            const int band = p.box_h / p.splits;
            for (int b = 0; b < p.splits; b++) tma3(dst + (size_t) b * band * p.box_w * 16, &map, 0, x0 / 8, y0 + b * band, full + s);
        } else {
            for (int r = 0; r < p.box_h; r++) bulk1(dst + (size_t) r * p.box_w * 16, p.img + (size_t) min(y0 + r, p.H - 1) * p.W + min(x0, p.W - p.box_w), p.box_w * 16, full + s);
        }
    }
    if (bad) sink[0] = 1;
}

__global__ void flush_kernel(float4 *buf, size_t n4) { for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) buf[i] = make_float4(1, 2, 3, 4); }

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int W = 640, H = 480;
    float4 *img; CK(cudaMalloc(&img, (size_t) W * H * 16 * 8));
    CK(cudaMemset(img, 0, (size_t) W * H * 16 * 8));
    float4 *fl; const size_t flb = (size_t) 384 << 20; CK(cudaMalloc(&fl, flb));
    int *sink; CK(cudaMalloc(&sink, 4)); CK(cudaMemset(sink, 0, 4));
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn) fp;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { const char *name; int kind, bw, bh, depth, splits, swz, l2, elem; };
    // the image stack is one tall image of 8*H rows
    std::vector<Cfg> cfgs = {
        {"2D u64 72x40 d3", 0, 72, 40, 3, 1, 0, 1, 8}, {"2D u64 72x40 d4", 0, 72, 40, 4, 1, 0, 1, 8}, {"2D u64 72x40 d4 L2-256", 0, 72, 40, 4, 1, 0, 2, 8}, {"2D u64 72x40 d4 split5", 0, 72, 40, 4, 5, 0, 1, 8},
        {"2D u64 72x40 d4 split10", 0, 72, 40, 4, 10, 0, 1, 8}, {"2D u64 72x40 d2", 0, 72, 40, 2, 1, 0, 1, 8}, {"2D u64 72x40 d1", 0, 72, 40, 1, 1, 0, 1, 8},
        {"2D f32 64x40 d4", 0, 64, 40, 4, 1, 0, 1, 4}, {"2D u64 40x40 d6", 0, 40, 40, 6, 1, 0, 1, 8}, {"2D u64 72x20 d8", 0, 72, 20, 8, 1, 0, 1, 8}, {"2D u64 128x40 d2", 0, 128, 40, 2, 1, 0, 1, 8},
        {"3D u64 chunk128B 80x40 d4", 1, 80, 40, 4, 1, 0, 1, 8}, {"3D u64 chunk128B 80x40 d4 swz128", 1, 80, 40, 4, 1, 3, 1, 8}, {"3D u64 chunk128B 80x40 d4 split5 swz128", 1, 80, 40, 4, 5, 3, 1, 8},
        {"bulk rows 72x40 d4", 2, 72, 40, 4, 1, 0, 1, 8}, {"bulk rows 72x40 d2", 2, 72, 40, 2, 1, 0, 1, 8},
    };
    printf("%-42s %10s %10s %12s %12s\n", "config", "warm_us", "cold_us", "warm_GB/s", "cold_GB/s");
    for (auto &c : cfgs) {
        CUtensorMap map;
        CUresult rc = CUDA_SUCCESS;
        const CUtensorMapL2promotion l2 = c.l2 == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        const CUtensorMapSwizzle sw = c.swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
        if (c.kind == 0) {
            const int epp = 16 / c.elem;   // elements per texel
            cuuint64_t gdim[2] = {(cuuint64_t) epp * W, (cuuint64_t) H * 8}; cuuint64_t gstr[1] = {(cuuint64_t) W * 16};
            cuuint32_t box[2] = {(cuuint32_t) (epp * c.bw), (cuuint32_t) (c.bh / c.splits)}; cuuint32_t es[2] = {1, 1};
            rc = enc(&map, c.elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else if (c.kind == 1) {
            cuuint64_t gdim[3] = {16, (cuuint64_t) W / 8, (cuuint64_t) H * 8}; cuuint64_t gstr[2] = {128, (cuuint64_t) W * 16};
            cuuint32_t box[3] = {16, (cuuint32_t) (c.bw / 8), (cuuint32_t) (c.bh / c.splits)}; cuuint32_t es[3] = {1, 1, 1};
            rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (rc != CUDA_SUCCESS) { printf("%-42s encode failed (%d)\n", c.name, (int) rc); continue; }
        Params p; p.kind = c.kind; p.box_w = c.bw; p.box_h = c.bh; p.depth = c.depth; p.per_cta = 9; p.splits = c.splits; p.W = W; p.H = H * 8; p.tiles_x = W / 64; p.tiles_y = H * 8 / 32; p.img = img;
        if (c.kind == 2) p.splits = 1;
        const size_t smem = (size_t) c.depth * c.bw * c.bh * 16;
        if (smem > 216 * 1024) { printf("%-42s smem too large\n", c.name); continue; }
        const double bytes = 148.0 * p.per_cta * c.bw * c.bh * 16;
        float tw = 0, tc = 0;
        for (int it = 0; it < 3; it++) stream_kernel<<<148, 32, smem>>>(map, p, sink);
        CK(cudaDeviceSynchronize());
        const int reps = 20;
        cudaEventRecord(e0); for (int it = 0; it < reps; it++) stream_kernel<<<148, 32, smem>>>(map, p, sink); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&tw, e0, e1); tw /= reps;
        for (int it = 0; it < 5; it++) {
            flush_kernel<<<148 * 8, 256>>>(fl, flb / 16);
            cudaEventRecord(e0); stream_kernel<<<148, 32, smem>>>(map, p, sink); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float t; cudaEventElapsedTime(&t, e0, e1); tc += t / 5;
        }
        printf("%-42s %10.2f %10.2f %12.1f %12.1f\n", c.name, tw * 1e3, tc * 1e3, bytes / (tw * 1e-3) / 1e9, bytes / (tc * 1e-3) / 1e9);
    }
    int hs = 0; CK(cudaMemcpy(&hs, sink, 4, cudaMemcpyDeviceToHost));
    printf("timeouts: %d\n", hs);
    return 0;
}
