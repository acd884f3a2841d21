// tracer.cu -- host side of the immature-point tracer and the C ABI of include/cmltrc.h (SURVEY.md 8f NEXT #2).
//
// Host-side reference anchors (under /root/reference/src/cml/optimization/dso):
//   Tracer::pair_table   DSOTracer.cpp:605-606 (hostToFrame_KRKi, hostToFrame_Kt), :431-432 (hostToTarget, exposure transition)
//   Tracer::window_order Map::getGroupFrames -> OrderedSet<PFrame, Comparator>: newest frame first (types.h:996-1012)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmltrc.h"
#include "se3.h"
#include "tracer.cuh"

namespace cmltrc {

static thread_local std::string g_create_error;
using cmlba::Pose;

#define RCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return CMLTRC_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

// Point ids are handed out once and never reused; the device arrays are indexed by SLOT.  Removed points leave dead slots behind, which
// compact_kernel squeezes out when more than a quarter of the slots are dead (Tracer::maybe_compact), so device memory and the cost of a
// trace pass follow the LIVE points, not every point ever created.
__global__ void compact_kernel(const PointsDev src, const PointsDev dst, const int n_live, const int *__restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_live) return;
    const int o = perm[i];
    dst.host[i] = src.host[o]; dst.xy[i] = src.xy[o]; dst.status[i] = src.status[o];
    dst.idmin[i] = src.idmin[o]; dst.idmax[i] = src.idmax[o]; dst.u[i] = src.u[o]; dst.v[i] = src.v[o];
    dst.interval[i] = src.interval[o]; dst.quality[i] = src.quality[o]; dst.energyTH[i] = src.energyTH[o];
    for (int k = 0; k < 4; k++) dst.gradH[(size_t) i * 4 + k] = src.gradH[(size_t) o * 4 + k];
    for (int k = 0; k < 8; k++) dst.weights[(size_t) i * 8 + k] = src.weights[(size_t) o * 8 + k];
}
// read-back of an arbitrary list of slots (-1 = removed point) as packed cmltrc_point records + pixel: one copy instead of ten
__global__ void gather_points_kernel(const PointsDev pts, const int n, const int *__restrict__ slot, cmltrc_point *__restrict__ out, float2 *__restrict__ xy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slot[i];
    cmltrc_point o;
    memset(&o, 0, sizeof o);
    o.host_frame_slot = -1;
    float2 q = make_float2(0.f, 0.f);
    if (s >= 0) {
        o.status = pts.status[s]; o.host_frame_slot = pts.host[s]; o.idepth_min = pts.idmin[s]; o.idepth_max = pts.idmax[s];
        o.last_trace_uv[0] = pts.u[s]; o.last_trace_uv[1] = pts.v[s]; o.last_trace_pixel_interval = pts.interval[s]; o.quality = pts.quality[s];
        for (int k = 0; k < 4; k++) o.grad_h[k] = pts.gradH[(size_t) s * 4 + k];
        o.energy_th = pts.energyTH[s];
        q = pts.xy[s];
    }
    out[i] = o;
    if (xy) xy[i] = q;
}

struct Tracer {
    TrcParams P{};
    int device = 0;
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // frames
    struct Slot { bool used = false; int64_t id = 0; Pose cam; double exposure[3]; float *gray = nullptr; float4 *grad = nullptr; };
    Slot slots[TRC_MAXF];
    FrameImg *d_frames = nullptr;
    PairDev *d_pairs = nullptr;        // [TRC_MAXF * TRC_MAXF]
    int *d_slots = nullptr, *d_counts = nullptr;
    char *h_pin = nullptr; size_t pin_bytes = 0;
    // points
    PointsDev pts{};
    size_t cap = 0; int64_t num = 0;   // slots in use (live + dead until the next compaction)
    std::vector<int> host_slot;        // [num] per SLOT: host mirror of pts.host (-1 = dead)
    int64_t next_id = 0, n_dead = 0;   // ids handed out so far (never reused); dead slots
    std::vector<int> id2slot;          // [next_id] slot of a live point, -1 once removed
    std::vector<int64_t> slot2id;      // [num]
    int *d_gidx = nullptr; cmltrc_point *d_gpt = nullptr; float2 *d_gxy = nullptr; size_t g_cap = 0;   // gather scratch
    int *d_ids = nullptr; ActivateOut *d_out = nullptr; size_t act_cap = 0;

    ~Tracer() {
        for (auto &s : slots) { if (s.gray) cudaFree(s.gray); if (s.grad) cudaFree(s.grad); }
        free_points();
        if (d_frames) cudaFree(d_frames); if (d_pairs) cudaFree(d_pairs); if (d_slots) cudaFree(d_slots); if (d_counts) cudaFree(d_counts);
        if (d_ids) cudaFree(d_ids); if (d_out) cudaFree(d_out);
        if (d_gidx) cudaFree(d_gidx); if (d_gpt) cudaFree(d_gpt); if (d_gxy) cudaFree(d_gxy);
        if (h_pin) cudaFreeHost(h_pin);
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    void free_points() {
        void *v[] = {pts.host, pts.xy, pts.status, pts.idmin, pts.idmax, pts.u, pts.v, pts.interval, pts.quality, pts.gradH, pts.energyTH, pts.weights};
        for (void *p : v) if (p) cudaFree(p);
        pts = PointsDev{};
    }

    int create(const cmltrc_config &c, int dev, int W, int H, double fx, double fy, double cx, double cy) {
        if (W < 16 || H < 16 || !(fx > 0) || !(fy > 0)) { error = "bad image size or calibration"; return CMLTRC_ERR_ARG; }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) { error = "no CUDA device " + std::to_string(dev) + " (the tracer has no CPU path)"; return CMLTRC_ERR_CUDA; }
        device = dev;
        RCK(cudaSetDevice(dev));
        P.W = W; P.H = H; P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
        P.huber = c.huber_threshold; P.outlier_th = c.outlier_th; P.outlier_th_sum = c.outlier_th_sum_component; P.max_pix_search = c.max_pix_search;
        P.max_slack_interval = c.max_slack_interval; P.step_size = c.trace_step_size; P.min_improvement = c.min_improvement_factor; P.test_radius = c.min_trace_test_radius;
        P.extra_slack = c.extra_slack_on_th; P.min_idepth_h_act = c.min_idepth_h_act; P.gn_iterations = c.gn_iterations;
        RCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        RCK(cudaEventCreate(&ev0)); RCK(cudaEventCreate(&ev1));
        RCK(cudaMalloc(&d_frames, sizeof(FrameImg) * TRC_MAXF));
        RCK(cudaMalloc(&d_pairs, sizeof(PairDev) * TRC_MAXF * TRC_MAXF));
        RCK(cudaMalloc(&d_slots, sizeof(int) * TRC_MAXF));
        RCK(cudaMalloc(&d_counts, sizeof(int) * 8));
        pin_bytes = std::max((size_t) W * H * 4, sizeof(PairDev) * TRC_MAXF * TRC_MAXF + 4096) + 4096;
        RCK(cudaHostAlloc((void **) &h_pin, pin_bytes, cudaHostAllocDefault));
        return CMLTRC_OK;
    }

    int find(int64_t id) const { for (int s = 0; s < TRC_MAXF; s++) if (slots[s].used && slots[s].id == id) return s; return -1; }

    int upload_frame_table() {
        FrameImg f[TRC_MAXF];
        for (int s = 0; s < TRC_MAXF; s++) { f[s].gray = slots[s].gray; f[s].grad = slots[s].grad; }
        RCK(cudaMemcpyAsync(d_frames, f, sizeof f, cudaMemcpyHostToDevice, stream));
        RCK(cudaStreamSynchronize(stream));
        return CMLTRC_OK;
    }

    int add_frame(int64_t id, const float *gray, const double *cam, const double *expo, const void *d_texels = nullptr) {
        if (!gray || !cam || !expo) { error = "NULL argument"; return CMLTRC_ERR_ARG; }
        if (find(id) >= 0) { error = "frame id already present"; return CMLTRC_ERR_ARG; }
        int s = 0;
        while (s < TRC_MAXF && slots[s].used) s++;
        if (s == TRC_MAXF) { error = "frame group is full (16 frames)"; return CMLTRC_ERR_ARG; }
        RCK(cudaSetDevice(device));
        Slot &sl = slots[s];
        const size_t px = (size_t) P.W * P.H;
        if (!sl.gray) { RCK(cudaMalloc(&sl.gray, px * 4)); RCK(cudaMalloc(&sl.grad, px * 16)); }
        if (d_texels) {           // `gray` and `d_texels` are DEVICE pointers (level 0 of cmlimg): two device-to-device copies
            RCK(cudaMemcpyAsync(sl.gray, gray, px * 4, cudaMemcpyDeviceToDevice, stream));
            RCK(cudaMemcpyAsync(sl.grad, d_texels, px * 16, cudaMemcpyDeviceToDevice, stream));
        } else {
            RCK(cudaStreamSynchronize(stream));
            memcpy(h_pin, gray, px * 4);
            RCK(cudaMemcpyAsync(sl.gray, h_pin, px * 4, cudaMemcpyHostToDevice, stream));
            trc_grad_kernel<<<std::min<int>(1184, (int) ((px + 255) / 256)), 256, 0, stream>>>(sl.gray, sl.grad, P.W, P.H);
            RCK(cudaGetLastError());
        }
        sl.used = true; sl.id = id;
        memcpy(sl.cam.R, cam, 72); memcpy(sl.cam.t, cam + 9, 24); memcpy(sl.exposure, expo, 24);
        return upload_frame_table();
    }
    int set_pose(int64_t id, const double *cam, const double *expo) {
        const int s = find(id);
        if (s < 0 || !cam || !expo) { error = "unknown frame id or NULL argument"; return CMLTRC_ERR_ARG; }
        memcpy(slots[s].cam.R, cam, 72); memcpy(slots[s].cam.t, cam + 9, 24); memcpy(slots[s].exposure, expo, 24);
        return CMLTRC_OK;
    }
    int remove_frame(int64_t id) {
        const int s = find(id);
        if (s < 0) { error = "unknown frame id"; return CMLTRC_ERR_ARG; }
        slots[s].used = false;
        std::vector<int64_t> dead;
        for (int64_t p = 0; p < num; p++) if (host_slot[p] == s) dead.push_back(slot2id[p]);
        return dead.empty() ? CMLTRC_OK : remove_points((int) dead.size(), dead.data());
    }

    int alloc_points(PointsDev &n, size_t ncap) {
        RCK(cudaMalloc((void **) &n.host, ncap * 4)); RCK(cudaMalloc((void **) &n.xy, ncap * 8)); RCK(cudaMalloc((void **) &n.status, ncap * 4));
        RCK(cudaMalloc((void **) &n.idmin, ncap * 8)); RCK(cudaMalloc((void **) &n.idmax, ncap * 8)); RCK(cudaMalloc((void **) &n.u, ncap * 8)); RCK(cudaMalloc((void **) &n.v, ncap * 8));
        RCK(cudaMalloc((void **) &n.interval, ncap * 8)); RCK(cudaMalloc((void **) &n.quality, ncap * 8)); RCK(cudaMalloc((void **) &n.gradH, ncap * 32));
        RCK(cudaMalloc((void **) &n.energyTH, ncap * 8)); RCK(cudaMalloc((void **) &n.weights, ncap * 32));
        return CMLTRC_OK;
    }
    // squeeze the dead slots out once they are more than a quarter of the slots in use
    int maybe_compact() {
        if (n_dead < 1024 || n_dead * 4 <= num) return CMLTRC_OK;
        std::vector<int> perm;
        perm.reserve((size_t) (num - n_dead));
        for (int64_t p = 0; p < num; p++) if (host_slot[p] >= 0) perm.push_back((int) p);
        const int live = (int) perm.size();
        PointsDev n{};
        int rc = alloc_points(n, cap);
        if (rc) return rc;
        rc = gather_reserve((size_t) std::max(live, 1));
        if (rc) return rc;
        if (live) {
            RCK(cudaMemcpyAsync(d_gidx, perm.data(), (size_t) live * 4, cudaMemcpyHostToDevice, stream));
            compact_kernel<<<(live + 255) / 256, 256, 0, stream>>>(pts, n, live, d_gidx);
            RCK(cudaGetLastError());
        }
        RCK(cudaStreamSynchronize(stream));
        free_points();
        pts = n;
        std::vector<int> hs(live); std::vector<int64_t> s2i(live);
        for (int i = 0; i < live; i++) { hs[i] = host_slot[perm[i]]; s2i[i] = slot2id[perm[i]]; id2slot[(size_t) s2i[i]] = i; }
        host_slot.swap(hs); slot2id.swap(s2i);
        num = live; n_dead = 0;
        return CMLTRC_OK;
    }
    int gather_reserve(size_t n) {
        if (n <= g_cap) return CMLTRC_OK;
        if (d_gidx) cudaFree(d_gidx); if (d_gpt) cudaFree(d_gpt); if (d_gxy) cudaFree(d_gxy);
        d_gidx = nullptr; d_gpt = nullptr; d_gxy = nullptr; g_cap = 0;
        const size_t want = n + n / 2 + 256;
        RCK(cudaMalloc(&d_gidx, want * 4)); RCK(cudaMalloc(&d_gpt, want * sizeof(cmltrc_point))); RCK(cudaMalloc(&d_gxy, want * sizeof(float2)));
        g_cap = want;
        return CMLTRC_OK;
    }
    // packed records (+ pixels) of the given point ids; a removed id yields host_frame_slot = -1
    int gather(int count, const int64_t *ids, int64_t first, cmltrc_point *out, float2 *xy) {
        if (count == 0) return CMLTRC_OK;
        RCK(cudaSetDevice(device));
        int rc = gather_reserve((size_t) count);
        if (rc) return rc;
        std::vector<int> sl(count);
        for (int i = 0; i < count; i++) sl[i] = id2slot[(size_t) (ids ? ids[i] : first + i)];
        RCK(cudaMemcpyAsync(d_gidx, sl.data(), (size_t) count * 4, cudaMemcpyHostToDevice, stream));
        gather_points_kernel<<<(count + 255) / 256, 256, 0, stream>>>(pts, count, d_gidx, d_gpt, xy ? d_gxy : nullptr);
        RCK(cudaGetLastError());
        RCK(cudaMemcpyAsync(out, d_gpt, (size_t) count * sizeof(cmltrc_point), cudaMemcpyDeviceToHost, stream));
        if (xy) RCK(cudaMemcpyAsync(xy, d_gxy, (size_t) count * sizeof(float2), cudaMemcpyDeviceToHost, stream));
        RCK(cudaStreamSynchronize(stream));
        return CMLTRC_OK;
    }

    int grow(size_t want) {
        if (want <= cap) return CMLTRC_OK;
        const size_t ncap = std::max(want + want / 2, (size_t) 4096);
        PointsDev n{};
        auto mv = [&](auto *&dst, auto *src, size_t per) -> cudaError_t {
            cudaError_t e = cudaMalloc((void **) &dst, ncap * per);
            if (e != cudaSuccess) return e;
            if (src && num) e = cudaMemcpyAsync(dst, src, (size_t) num * per, cudaMemcpyDeviceToDevice, stream);
            return e;
        };
        RCK(mv(n.host, pts.host, 4)); RCK(mv(n.xy, pts.xy, 8)); RCK(mv(n.status, pts.status, 4));
        RCK(mv(n.idmin, pts.idmin, 8)); RCK(mv(n.idmax, pts.idmax, 8)); RCK(mv(n.u, pts.u, 8)); RCK(mv(n.v, pts.v, 8));
        RCK(mv(n.interval, pts.interval, 8)); RCK(mv(n.quality, pts.quality, 8)); RCK(mv(n.gradH, pts.gradH, 32)); RCK(mv(n.energyTH, pts.energyTH, 8));
        RCK(mv(n.weights, pts.weights, 32));
        RCK(cudaStreamSynchronize(stream));
        free_points();
        pts = n; cap = ncap;
        return CMLTRC_OK;
    }

    int make_new_traces(int64_t id, int count, const float *xy, int64_t *first) {
        const int s = find(id);
        if (s < 0 || count < 0 || (count > 0 && !xy)) { error = "unknown frame id or bad arguments"; return CMLTRC_ERR_ARG; }
        for (int i = 0; i < count; i++) {      // the pattern (|offset| <= 2) and its bilinear taps must stay inside the image
            const float x = xy[2 * i], y = xy[2 * i + 1];
            if (!(x >= 3 && y >= 3 && x < P.W - 4 && y < P.H - 4)) { error = "corner too close to the image border"; return CMLTRC_ERR_ARG; }
        }
        RCK(cudaSetDevice(device));
        int rc = maybe_compact();
        if (rc) return rc;
        rc = grow((size_t) num + count);
        if (rc) return rc;
        if (first) *first = next_id;
        if (count == 0) return CMLTRC_OK;
        RCK(cudaMemcpyAsync(pts.xy + num, xy, (size_t) count * 8, cudaMemcpyHostToDevice, stream));
        const FrameImg img{slots[s].gray, slots[s].grad};
        trc_init_kernel<<<(count + 127) / 128, 128, 0, stream>>>(P, pts, (int) num, count, s, img);
        RCK(cudaGetLastError());
        RCK(cudaStreamSynchronize(stream));
        host_slot.resize((size_t) num + count, s);
        for (int i = 0; i < count; i++) { id2slot.push_back((int) num + i); slot2id.push_back(next_id + i); }
        num += count; next_id += count;
        return CMLTRC_OK;
    }

    int remove_points(int count, const int64_t *ids) {
        if (count < 0 || (count > 0 && !ids)) { error = "bad arguments"; return CMLTRC_ERR_ARG; }
        for (int i = 0; i < count; i++) if (ids[i] < 0 || ids[i] >= next_id) { error = "unknown point id"; return CMLTRC_ERR_ARG; }     // validate everything before anything changes
        RCK(cudaSetDevice(device));
        int lo = INT32_MAX, hi = -1;
        for (int i = 0; i < count; i++) {
            const int sl = id2slot[(size_t) ids[i]];
            if (sl < 0) continue;                      // already removed
            host_slot[sl] = -1; id2slot[(size_t) ids[i]] = -1; n_dead++;
            lo = std::min(lo, sl); hi = std::max(hi, sl);
        }
        if (hi >= 0) RCK(cudaMemcpyAsync(pts.host + lo, host_slot.data() + lo, (size_t) (hi - lo + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));     // the changed range of the mirror
        RCK(cudaStreamSynchronize(stream));
        return CMLTRC_OK;
    }

    // pair constants in fp64; for trace only the column `target` is needed, for activation the whole table
    void fill_pair(PairDev &pd, const Slot &h, const Slot &t) {
        const Pose rel = cmlba::pose_mul(t.cam, cmlba::pose_inv(h.cam));      // Camera::to
        const double K[9] = {P.fx, 0, P.cx, 0, P.fy, P.cy, 0, 0, 1};
        const double Ki[9] = {1.0 / P.fx, 0, -P.cx / P.fx, 0, 1.0 / P.fy, -P.cy / P.fy, 0, 0, 1};
        double KR[9];
        cmlba::mat3_mul(K, rel.R, KR);
        cmlba::mat3_mul(KR, Ki, pd.KRKi);
        cmlba::mat3_vec(K, rel.t, pd.Kt);
        memcpy(pd.R, rel.R, 72); memcpy(pd.t, rel.t, 24);
        pd.a = std::exp(t.exposure[1] - h.exposure[1]) * t.exposure[0] / h.exposure[0];       // Exposure::to
        pd.b = t.exposure[2] - pd.a * h.exposure[2];
    }

    int trace(int64_t id, int32_t *hist, float *gpu_ms) {
        const int tgt = find(id);
        if (tgt < 0) { error = "unknown frame id"; return CMLTRC_ERR_ARG; }
        RCK(cudaSetDevice(device));
        RCK(cudaStreamSynchronize(stream));
        PairDev *hp = (PairDev *) h_pin;
        for (int s = 0; s < TRC_MAXF; s++) if (slots[s].used) fill_pair(hp[s], slots[s], slots[tgt]);
        RCK(cudaMemcpyAsync(d_pairs, hp, sizeof(PairDev) * TRC_MAXF, cudaMemcpyHostToDevice, stream));
        RCK(cudaMemsetAsync(d_counts, 0, 32, stream));
        RCK(cudaEventRecord(ev0, stream));
        if (num > 0) trc_trace_kernel<<<(int) ((num + 3) / 4), 128, 0, stream>>>(P, pts, (int) num, tgt, d_pairs, d_frames, d_counts);
        RCK(cudaEventRecord(ev1, stream));
        RCK(cudaGetLastError());
        int32_t c[8];
        RCK(cudaMemcpyAsync(c, d_counts, 32, cudaMemcpyDeviceToHost, stream));
        RCK(cudaStreamSynchronize(stream));
        if (hist) memcpy(hist, c, 24);
        if (gpu_ms) RCK(cudaEventElapsedTime(gpu_ms, ev0, ev1));
        return CMLTRC_OK;
    }

    int activate(int count, const int64_t *ids, int min_obs, cmltrc_activation *res, float *gpu_ms) {
        if (count < 0 || (count > 0 && (!ids || !res))) { error = "bad arguments"; return CMLTRC_ERR_ARG; }
        if (gpu_ms) *gpu_ms = 0.f;
        if (count == 0) return CMLTRC_OK;
        RCK(cudaSetDevice(device));
        std::vector<int> id32(count);
        for (int i = 0; i < count; i++) {
            if (ids[i] < 0 || ids[i] >= next_id || id2slot[(size_t) ids[i]] < 0) { error = "unknown or removed point id"; return CMLTRC_ERR_ARG; }
            id32[i] = id2slot[(size_t) ids[i]];
        }
        // window order: newest frame first (OrderedSet<PFrame, Comparator> orders by descending id)
        std::vector<int> order;
        for (int s = 0; s < TRC_MAXF; s++) if (slots[s].used) order.push_back(s);
        std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[a].id > slots[b].id; });
        RCK(cudaStreamSynchronize(stream));
        PairDev *hp = (PairDev *) h_pin;
        for (int a : order) for (int b : order) if (a != b) fill_pair(hp[a * TRC_MAXF + b], slots[a], slots[b]);
        RCK(cudaMemcpyAsync(d_pairs, hp, sizeof(PairDev) * TRC_MAXF * TRC_MAXF, cudaMemcpyHostToDevice, stream));
        RCK(cudaMemcpyAsync(d_slots, order.data(), order.size() * 4, cudaMemcpyHostToDevice, stream));
        if ((size_t) count > act_cap) {
            if (d_ids) cudaFree(d_ids); if (d_out) cudaFree(d_out);
            d_ids = nullptr; d_out = nullptr; act_cap = 0;
            RCK(cudaMalloc(&d_ids, (size_t) count * 2 * 4)); RCK(cudaMalloc(&d_out, (size_t) count * 2 * sizeof(ActivateOut)));
            act_cap = (size_t) count * 2;
        }
        RCK(cudaMemcpyAsync(d_ids, id32.data(), (size_t) count * 4, cudaMemcpyHostToDevice, stream));
        RCK(cudaEventRecord(ev0, stream));
        trc_activate_kernel<<<(count + 3) / 4, 128, 0, stream>>>(P, pts, count, d_ids, min_obs, (int) order.size(), d_slots, d_pairs, d_frames, d_out);
        RCK(cudaEventRecord(ev1, stream));
        RCK(cudaGetLastError());
        std::vector<ActivateOut> ho(count);
        RCK(cudaMemcpyAsync(ho.data(), d_out, (size_t) count * sizeof(ActivateOut), cudaMemcpyDeviceToHost, stream));
        RCK(cudaStreamSynchronize(stream));
        for (int i = 0; i < count; i++) { res[i].rc = ho[i].rc; res[i].idepth = ho[i].idepth; res[i].in_mask = ho[i].in_mask; }
        if (gpu_ms) RCK(cudaEventElapsedTime(gpu_ms, ev0, ev1));
        return CMLTRC_OK;
    }

    // ---- activatePoints: host control flow of DSOTracer.cpp:62-278 around the activation kernel
    double min_distance = 2.0;         // mCurrentMinimumDistance

    // utils/DistanceMap.h is an 8-neighbour BFS capped at maxDist = the Chebyshev distance to the nearest added integer pixel; a uniform grid of
    // cells >= maxDist answers it from the 3x3 neighbourhood
    struct ChebGrid {
        int W, H, maxd, cell, gw, gh;
        std::vector<std::vector<int>> cells;      // packed (y << 16 | x)
        ChebGrid(int w, int h, int md) : W(w), H(h), maxd(md), cell(std::max(md, 8)), gw((w + cell - 1) / cell), gh((h + cell - 1) / cell), cells((size_t) gw * gh) {}
        void add(double fx, double fy) {
            const int x = (int) fx, y = (int) fy;                       // _add(point.x(), point.y()): truncation, out-of-image points are ignored by queue()
            if (x < 0 || y < 0 || x >= W || y >= H) return;
            cells[(size_t) (y / cell) * gw + x / cell].push_back((y << 16) | x);
        }
        int get(double fx, double fy) const {
            const int x = (int) fx, y = (int) fy, cx = x / cell, cy = y / cell;
            int best = maxd;
            for (int gy = std::max(cy - 1, 0); gy <= std::min(cy + 1, gh - 1); gy++)
                for (int gx = std::max(cx - 1, 0); gx <= std::min(cx + 1, gw - 1); gx++)
                    for (int q : cells[(size_t) gy * gw + gx]) best = std::min(best, std::max(std::abs((q & 0xffff) - x), std::abs((q >> 16) - y)));
            return best;
        }
    };

    int activate_points(int64_t last_id, int nact, const double *axy, int desired, float min_quality, int nim, const int64_t *ids, const float *types, int capacity,
                        int64_t *act_ids, cmltrc_activation *act, int32_t *n_act, int64_t *rem_ids, int32_t *n_rem, cmltrc_activate_stats *st) {
        const int last = find(last_id);
        if (last < 0 || nact < 0 || nim < 0 || (nact > 0 && !axy) || (nim > 0 && !ids) || !n_act || !n_rem || capacity < 0 || (capacity > 0 && (!act_ids || !act || !rem_ids))) {
            error = "unknown frame id or bad arguments"; return CMLTRC_ERR_ARG;
        }
        for (int i = 0; i < nim; i++) if (ids[i] < 0 || ids[i] >= next_id || id2slot[(size_t) ids[i]] < 0) { error = "unknown or removed point id"; return CMLTRC_ERR_ARG; }
        cmltrc_activate_stats s{};
        // minimum-distance adaptation (DSOTracer.cpp:62-85)
        double md = min_distance;
        const int n = nact;
        if (n < desired * 0.66) md -= 0.8;
        if (n < desired * 0.8) md -= 0.5; else if (n < desired * 0.9) md -= 0.2; else if (n < desired) md -= 0.1;
        if (n > desired * 1.5) md += 0.8;
        if (n > desired * 1.3) md += 0.5;
        if (n > desired * 1.15) md += 0.2;
        if (n > desired) md += 0.1;
        s.urgently_need_new_points = md < 1 ? 1 : 0;
        if (md < 0) md = 0;
        if (md > 4) md = 4;
        min_distance = md; s.current_minimum_distance = md;
        const float maxType = 10;
        ChebGrid dmap(P.W, P.H, (int) (md * maxType));
        for (int i = 0; i < nact; i++) dmap.add(axy[2 * i], axy[2 * i + 1]);
        // states of the candidates only (one gathered read-back), then the sequential gating in the caller's order
        std::vector<cmltrc_point> pt((size_t) nim);
        std::vector<float2> pxy((size_t) nim);
        if (nim) { int rc = gather(nim, ids, 0, pt.data(), pxy.data()); if (rc) return rc; }
        std::vector<int64_t> to_opt, removed;
        std::vector<int> to_opt_k;
        for (int i = 0; i < nim; i++) {
            const int64_t id = ids[i];
            const cmltrc_point &p = pt[i];
            if (p.host_frame_slot == last) continue;
            if (!std::isfinite(p.idepth_max) || p.status == IPS_OUTLIER) { removed.push_back(id); s.num_deleted_outlier++; continue; }
            const bool okStatus = p.status == IPS_GOOD || p.status == IPS_SKIPPED || p.status == IPS_BADCONDITION || p.status == IPS_OOB;
            const bool okInterval = p.last_trace_pixel_interval < 8, okQuality = p.quality > (double) min_quality, okDepth = (p.idepth_max + p.idepth_min) > 0;
            if (!okStatus) s.num_skipped_status++;
            if (!okInterval) s.num_skipped_pixel_interval++;
            if (!okQuality) s.num_skipped_quality++;
            if (!okDepth) s.num_skipped_depth++;
            if (!(okStatus && okInterval && okQuality && okDepth)) {
                if (p.status == IPS_OOB) { removed.push_back(id); s.num_deleted_oob++; }
                continue;
            }
            // getWorldCoordinateIf(idepth).project(lastFrame): host pixel -> last frame
            const double idepth = (p.idepth_min + p.idepth_max) / 2.0;
            const Slot &hs = slots[p.host_frame_slot];
            const Pose rel = cmlba::pose_mul(slots[last].cam, cmlba::pose_inv(hs.cam));
            const float2 xy = pxy[i];
            const double ray[3] = {((double) xy.x - P.cx) / P.fx / idepth, ((double) xy.y - P.cy) / P.fy / idepth, 1.0 / idepth};
            double X[3];
            cmlba::mat3_vec(rel.R, ray, X);
            for (int k = 0; k < 3; k++) X[k] += rel.t[k];
            const double px = P.fx * (X[0] / X[2]) + P.cx, py = P.fy * (X[1] / X[2]) + P.cy;
            if (px >= 0 && py >= 0 && px < P.W && py < P.H) {
                const double dist = dmap.get(px, py) + (px - std::floor(px));
                if (dist >= md * (double) (types ? types[i] : 1.f)) { dmap.add(px, py); to_opt.push_back(id); to_opt_k.push_back(i); }
            } else removed.push_back(id);
        }
        s.num_to_optimize = (int) to_opt.size();
        std::vector<cmltrc_activation> res(to_opt.size());
        if (!to_opt.empty()) { int rc = activate((int) to_opt.size(), to_opt.data(), 1, res.data(), nullptr); if (rc) return rc; }
        int na = 0;
        std::vector<int64_t> gone;
        for (size_t k = 0; k < to_opt.size(); k++) {
            if (res[k].rc == 1) {
                if (na < capacity) { act_ids[na] = to_opt[k]; act[na] = res[k]; }
                na++; s.num_mapped++; gone.push_back(to_opt[k]);
            } else if (res[k].rc == -1 || pt[to_opt_k[k]].status == IPS_OOB) { removed.push_back(to_opt[k]); s.num_dropped++; }
            else s.num_non_mapped++;
        }
        for (size_t k = 0; k < removed.size() && (int) k < capacity; k++) rem_ids[k] = removed[k];
        *n_act = na; *n_rem = (int32_t) removed.size();
        gone.insert(gone.end(), removed.begin(), removed.end());
        if (!gone.empty()) { int rc = remove_points((int) gone.size(), gone.data()); if (rc) return rc; }
        if (st) *st = s;
        return CMLTRC_OK;
    }

    int get_points(int64_t first, int count, cmltrc_point *out) {
        if (first < 0 || count < 0 || first + count > next_id || (count > 0 && !out)) { error = "bad range"; return CMLTRC_ERR_ARG; }
        return gather(count, nullptr, first, out, nullptr);
    }
};

}  // namespace cmltrc

using cmltrc::Tracer;
#define TH(h) reinterpret_cast<Tracer *>(h)

extern "C" {

void cmltrc_default_config(cmltrc_config *c) {
    if (!c) return;
    c->min_idepth_h_act = 100.0f; c->gn_iterations = 3; c->huber_threshold = 9.0f; c->outlier_th = 12.0f * 12.0f; c->outlier_th_sum_component = 50.0f * 50.0f;
    c->max_pix_search = 0.027f; c->max_slack_interval = 1.5f; c->trace_step_size = 1.0f; c->min_improvement_factor = 2.0f; c->min_trace_test_radius = 2.0f;
    c->extra_slack_on_th = 1.2f;
}

int cmltrc_create(const cmltrc_config *cfg, int device, int width, int height, double fx, double fy, double cx, double cy, cmltrc_handle *out) {
    if (!out) { cmltrc::g_create_error = "out is NULL"; return CMLTRC_ERR_ARG; }
    *out = nullptr;
    cmltrc_config c;
    if (cfg) c = *cfg; else cmltrc_default_config(&c);
    Tracer *t = new Tracer();
    const int rc = t->create(c, device, width, height, fx, fy, cx, cy);
    if (rc) { cmltrc::g_create_error = t->error; delete t; return rc; }
    *out = reinterpret_cast<cmltrc_handle>(t);
    return CMLTRC_OK;
}
void cmltrc_destroy(cmltrc_handle h) { delete TH(h); }
const char *cmltrc_last_error(cmltrc_handle h) { return h ? TH(h)->error.c_str() : cmltrc::g_create_error.c_str(); }
int cmltrc_add_frame(cmltrc_handle h, int64_t id, const float *gray, const double cam[12], const double exposure[3]) { return h ? TH(h)->add_frame(id, gray, cam, exposure) : CMLTRC_ERR_ARG; }
int cmltrc_add_frame_device(cmltrc_handle h, int64_t id, const float *d_gray, const void *d_texels, const double cam[12], const double exposure[3]) {
    if (!h || !d_texels) return CMLTRC_ERR_ARG;
    return TH(h)->add_frame(id, d_gray, cam, exposure, d_texels);
}
int cmltrc_set_frame_pose(cmltrc_handle h, int64_t id, const double cam[12], const double exposure[3]) { return h ? TH(h)->set_pose(id, cam, exposure) : CMLTRC_ERR_ARG; }
int cmltrc_remove_frame(cmltrc_handle h, int64_t id) { return h ? TH(h)->remove_frame(id) : CMLTRC_ERR_ARG; }
int cmltrc_make_new_traces(cmltrc_handle h, int64_t id, int count, const float *xy, int64_t *first_id) { return h ? TH(h)->make_new_traces(id, count, xy, first_id) : CMLTRC_ERR_ARG; }
int cmltrc_remove_points(cmltrc_handle h, int count, const int64_t *ids) { return h ? TH(h)->remove_points(count, ids) : CMLTRC_ERR_ARG; }
int64_t cmltrc_num_points(cmltrc_handle h) { return h ? TH(h)->next_id : CMLTRC_ERR_ARG; }
int cmltrc_trace_new_coarse(cmltrc_handle h, int64_t id, int32_t *hist, float *gpu_ms) { return h ? TH(h)->trace(id, hist, gpu_ms) : CMLTRC_ERR_ARG; }
int cmltrc_optimize_immature(cmltrc_handle h, int count, const int64_t *ids, int min_obs, cmltrc_activation *results, float *gpu_ms) {
    return h ? TH(h)->activate(count, ids, min_obs, results, gpu_ms) : CMLTRC_ERR_ARG;
}
int cmltrc_get_points(cmltrc_handle h, int64_t first_id, int count, cmltrc_point *out) { return h ? TH(h)->get_points(first_id, count, out) : CMLTRC_ERR_ARG; }
int cmltrc_activate_points(cmltrc_handle h, int64_t last_frame_id, int num_active, const double *active_xy, int desired_point_density, float min_trace_quality, int num_immature,
                           const int64_t *immature_ids, const float *immature_types, int capacity, int64_t *activated_ids, cmltrc_activation *activated, int32_t *num_activated,
                           int64_t *removed_ids, int32_t *num_removed, cmltrc_activate_stats *stats) {
    return h ? TH(h)->activate_points(last_frame_id, num_active, active_xy, desired_point_density, min_trace_quality, num_immature, immature_ids, immature_types, capacity,
                                      activated_ids, activated, num_activated, removed_ids, num_removed, stats)
             : CMLTRC_ERR_ARG;
}
int cmltrc_set_minimum_distance(cmltrc_handle h, double v) { if (!h) return CMLTRC_ERR_ARG; TH(h)->min_distance = v; return CMLTRC_OK; }

}  // extern "C"
