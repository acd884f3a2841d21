// tracker.cu -- host side of the coarse tracker and the C ABI of include/cmltrk.h (SURVEY.md 8f NEXT #1).
//
// Host-side reference anchors (under /root/reference/src/cml):
//   Tracker::create            capture/CaptureImage.cpp:39-72 (pyramid size rule), PinholeUndistorter level calibration
//   Tracker::make_coarse_depth optimization/dso/DSOTracker.cpp:494-725
//   Tracker::set_frame         capture/CaptureImage.cpp:209-262
//   Tracker::optimize          optimization/dso/DSOTracker.cpp:15-59 (refToNew), :240 (camera = reference o refToNew)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cmltrk.h"
#include "tracker.cuh"

namespace cmltrk {

static thread_local std::string g_create_error;

#define TCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return CMLTRK_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

struct Tracker {
    double last_coarse_rmse = INFINITY, first_rmse = -1.0;     // mLastCoarseRMSE / mFirstRMSE of the class (trackWithMotionModel)
    cmltrk_config cfg{};
    int device = 0, W = 0, H = 0, L = 0;
    int w[MAXL]{}, h[MAXL]{};
    double K[MAXL][4]{};
    std::string error;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    char *d_block = nullptr;
    char *h_block = nullptr;     // pinned staging
    size_t h_bytes = 0, stage_gray = 0;
    // device views
    float *gray_new[MAXL]{}, *gray_ref[MAXL]{};
    float4 *grad_new[MAXL]{}, *pc[MAXL]{};
    long long *mapI[MAXL]{}, *mapW[MAXL]{};
    int *row_count[MAXL]{}, *row_offset[MAXL]{}, *pc_n = nullptr;
    Candidate *d_cand = nullptr;
    TrackOut *d_out = nullptr;
    char *d_pts = nullptr; size_t pts_cap = 0;
    bool have_ref = false, have_frame = false, frame_external = false, ref_levels_on_device = false;
    const float4 *ext_grad[MAXL]{};
    Pose ref_pose{};
    double ref_exposure[3]{1, 0, 0}, new_tau = 1.0;
    TrackParams params{};
    int last_K = 0;
    int cluster = 8, threads = 384;
    typedef void (*TrackFn)(const TrackParams, const Candidate *, TrackOut *);
    TrackFn kernel() const { return threads == 256 ? track_kernel<256> : threads == 512 ? track_kernel<512> : track_kernel<384>; }
    long launches = 0;

    ~Tracker() {
        if (d_block) cudaFree(d_block);
        if (d_pts) cudaFree(d_pts);
        if (h_block) cudaFreeHost(h_block);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }

    int create(const cmltrk_config &c, int dev, int width, int height, double fx, double fy, double cx, double cy) {
        cfg = c; device = dev; W = width; H = height;
        if (width < 16 || height < 16 || !(fx > 0) || !(fy > 0)) { error = "bad image size or calibration"; return CMLTRK_ERR_ARG; }
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || dev < 0 || dev >= count) { error = "no CUDA device " + std::to_string(dev) + " (the tracker has no CPU path)"; return CMLTRK_ERR_CUDA; }
        TCK(cudaSetDevice(dev));
        // pyramid size (CaptureImage.cpp:39-72): halve until the area is <= 25 x 25 and at least 5 levels exist
        int levels = cfg.levels;
        if (levels <= 0) {
            levels = 0;
            double sx = width, sy = height;
            for (;;) {
                if (sx * sy <= 25.0 * 25.0 && levels >= 5) break;
                levels++; sx /= 2; sy /= 2;
                if (levels >= 32) break;
            }
        }
        L = std::min(levels, (int) MAXL);
        for (int l = 0; l < L; l++) {
            w[l] = l ? w[l - 1] / 2 : width; h[l] = l ? h[l - 1] / 2 : height;
            if (w[l] < 1 || h[l] < 1) { L = l; break; }
            const double s = std::ldexp(1.0, l);
            K[l][0] = fx / s; K[l][1] = fy / s; K[l][2] = (cx + 0.5) / s - 0.5; K[l][3] = (cy + 0.5) / s - 0.5;
        }
        threads = (cfg.cta_threads == 256 || cfg.cta_threads == 512) ? cfg.cta_threads : 384;
        cluster = std::max(1, std::min(cfg.cluster_ctas > 0 ? cfg.cluster_ctas : 16, (int) TRK_MAX_CLUSTER));
        TCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        TCK(cudaEventCreate(&ev0)); TCK(cudaEventCreate(&ev1));
        // one device block for everything that scales with the image
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = (off + 255) & ~(size_t) 255; off = o + bytes; return o; };
        size_t o_gn[MAXL], o_gr[MAXL], o_gd[MAXL], o_pc[MAXL], o_mi[MAXL], o_mw[MAXL], o_rc[MAXL], o_ro[MAXL];
        for (int l = 0; l < L; l++) {
            const size_t px = (size_t) w[l] * h[l];
            o_gn[l] = take(px * 4); o_gr[l] = take(px * 4); o_gd[l] = take(px * 16); o_pc[l] = take(px * 16);
            o_mi[l] = take(px * 8); o_mw[l] = take(px * 8); o_rc[l] = take((size_t) h[l] * 4); o_ro[l] = take((size_t) h[l] * 4);
        }
        const size_t o_n = take(MAXL * 4), o_cand = take(sizeof(Candidate) * CMLTRK_MAX_CANDIDATES), o_out = take(sizeof(TrackOut) * CMLTRK_MAX_CANDIDATES);
        TCK(cudaMalloc(&d_block, off));
        TCK(cudaMemsetAsync(d_block, 0, off, stream));
        for (int l = 0; l < L; l++) {
            gray_new[l] = (float *) (d_block + o_gn[l]); gray_ref[l] = (float *) (d_block + o_gr[l]); grad_new[l] = (float4 *) (d_block + o_gd[l]);
            pc[l] = (float4 *) (d_block + o_pc[l]); mapI[l] = (long long *) (d_block + o_mi[l]); mapW[l] = (long long *) (d_block + o_mw[l]);
            row_count[l] = (int *) (d_block + o_rc[l]); row_offset[l] = (int *) (d_block + o_ro[l]);
        }
        pc_n = (int *) (d_block + o_n); d_cand = (Candidate *) (d_block + o_cand); d_out = (TrackOut *) (d_block + o_out);
        // pinned staging: [start poses | results | gray image]; the image part is also handed out by cmltrk_frame_buffer (zero-copy producers)
        stage_gray = ((sizeof(TrackOut) + sizeof(Candidate)) * CMLTRK_MAX_CANDIDATES + 4095) & ~(size_t) 4095;
        h_bytes = stage_gray + (size_t) W * H * 4;
        TCK(cudaHostAlloc((void **) &h_block, h_bytes, cudaHostAllocDefault));
        if (cluster > 8) {       // 16 CTAs per cluster is the non-portable size: take it when this device can co-schedule one, else fall back to 8
            bool ok = cudaFuncSetAttribute(kernel(), cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
            if (ok) {
                cudaLaunchConfig_t lc{};
                lc.gridDim = dim3(cluster); lc.blockDim = dim3(threads);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                int n = 0;
                ok = cudaOccupancyMaxActiveClusters(&n, kernel(), &lc) == cudaSuccess && n >= 1;
            }
            if (!ok) { cudaGetLastError(); cluster = 8; }
        }
        TCK(cudaStreamSynchronize(stream));
        return CMLTRK_OK;
    }

    PyrDev pyr(float **gray, bool with_grad) {
        PyrDev p{};
        p.levels = L;
        for (int l = 0; l < L; l++) { p.w[l] = w[l]; p.h[l] = h[l]; p.gray[l] = gray[l]; p.grad[l] = with_grad ? grad_new[l] : nullptr; }
        return p;
    }

    // host gray -> pinned -> device level 0, then the pyramid (and the derivative texels for the frame to track)
    int upload_pyramid(const float *gray, float **dst, bool with_grad) {
        const size_t bytes = (size_t) W * H * 4;
        float *stage = (float *) (h_block + stage_gray);
        if (gray != stage) {
            TCK(cudaStreamSynchronize(stream));       // a previous upload may still be reading the staging image
            memcpy(stage, gray, bytes);
        }
        TCK(cudaMemcpyAsync(dst[0], stage, bytes, cudaMemcpyHostToDevice, stream));
        const PyrDev p = pyr(dst, with_grad);
        const dim3 tiles((W + PYR_TILE - 1) / PYR_TILE, (H + PYR_TILE - 1) / PYR_TILE);
        pyr_gray_kernel<<<tiles, 256, 0, stream>>>(p); launches++;
        if (with_grad) {
            const dim3 g(std::min(592, (W * H + 255) / 256), L);
            pyr_grad_kernel<<<g, 256, 0, stream>>>(p); launches++;
        }
        TCK(cudaGetLastError());
        return CMLTRK_OK;
    }

    // frame to track from device-resident texel levels (float4 (I, dx, dy, *), e.g. cmlimg's): zero-copy, the pointers are sampled in place
    int set_frame_device(int levels, const void *const *d_texels, double tau) {
        if (levels < std::min(L, (int) OPTL) || !d_texels) { error = "need the texels of every level the optimisation uses"; return CMLTRK_ERR_ARG; }
        for (int l = 0; l < std::min(L, (int) OPTL); l++) if (!d_texels[l]) { error = "NULL level"; return CMLTRK_ERR_ARG; }
        for (int l = 0; l < MAXL; l++) ext_grad[l] = (l < levels && l < L) ? static_cast<const float4 *>(d_texels[l]) : nullptr;
        new_tau = tau; have_frame = true; frame_external = true;
        return CMLTRK_OK;
    }

    int set_frame(const float *gray, double tau) {
        if (!gray) { error = "gray is NULL"; return CMLTRK_ERR_ARG; }
        TCK(cudaSetDevice(device));
        int rc = upload_pyramid(gray, gray_new, true);
        if (rc) return rc;
        new_tau = tau; have_frame = true; frame_external = false;
        return CMLTRK_OK;
    }

    int make_coarse_depth(const float *ref_gray, const double *ref_cam, const double *ref_exp, int F, const double *frame_cams, int P, const int32_t *pt_frame,
                          const float *pt_xy, const double *pt_idepth, const double *pt_unc) {
        if (!ref_gray || !ref_cam || !ref_exp || F < 0 || P < 0 || (P > 0 && (!frame_cams || !pt_frame || !pt_xy || !pt_idepth || !pt_unc))) { error = "NULL argument"; return CMLTRK_ERR_ARG; }
        for (int i = 0; i < P; i++) if (pt_frame[i] < 0 || pt_frame[i] >= F) { error = "pt_frame out of range"; return CMLTRK_ERR_ARG; }
        TCK(cudaSetDevice(device));
        if (ref_levels_on_device) {       // ref_gray = array of L device pointers to the gray levels (cmlimg): device-to-device copies, no pyramid kernel
            const float *const *lv = reinterpret_cast<const float *const *>(ref_gray);
            for (int l = 0; l < L; l++) {
                if (!lv[l]) { error = "NULL gray level"; return CMLTRK_ERR_ARG; }
                TCK(cudaMemcpyAsync(gray_ref[l], lv[l], (size_t) w[l] * h[l] * 4, cudaMemcpyDeviceToDevice, stream));
            }
        } else {
            int rc = upload_pyramid(ref_gray, gray_ref, false);
            if (rc) return rc;
        }
        memcpy(ref_pose.R, ref_cam, 72); memcpy(ref_pose.t, ref_cam + 9, 24);
        memcpy(ref_exposure, ref_exp, 24);
        // host-to-reference transforms in fp64 (Camera::to), one row of 12 per frame
        std::vector<double> rel((size_t) std::max(F, 1) * 12);
        for (int f = 0; f < F; f++) {
            Pose hp; memcpy(hp.R, frame_cams + (size_t) f * 12, 72); memcpy(hp.t, frame_cams + (size_t) f * 12 + 9, 24);
            const Pose r = cmlba::pose_mul(ref_pose, cmlba::pose_inv(hp));
            memcpy(&rel[(size_t) f * 12], r.R, 72); memcpy(&rel[(size_t) f * 12 + 9], r.t, 24);
        }
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = (off + 255) & ~(size_t) 255; off = o + bytes; return o; };
        const size_t o_rel = take(rel.size() * 8), o_fr = take((size_t) P * 4), o_xy = take((size_t) P * 8), o_id = take((size_t) P * 8), o_un = take((size_t) P * 8);
        if (off > pts_cap) {
            if (d_pts) cudaFree(d_pts);
            d_pts = nullptr; pts_cap = 0;
            TCK(cudaMalloc(&d_pts, off + off / 4));
            pts_cap = off + off / 4;
        }
        TCK(cudaMemcpyAsync(d_pts + o_rel, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, stream));
        if (P) {
            TCK(cudaMemcpyAsync(d_pts + o_fr, pt_frame, (size_t) P * 4, cudaMemcpyHostToDevice, stream));
            TCK(cudaMemcpyAsync(d_pts + o_xy, pt_xy, (size_t) P * 8, cudaMemcpyHostToDevice, stream));
            TCK(cudaMemcpyAsync(d_pts + o_id, pt_idepth, (size_t) P * 8, cudaMemcpyHostToDevice, stream));
            TCK(cudaMemcpyAsync(d_pts + o_un, pt_unc, (size_t) P * 8, cudaMemcpyHostToDevice, stream));
        }
        CoarseDev c{};
        c.levels = L; c.pc_n = pc_n;
        c.fx = K[0][0]; c.fy = K[0][1]; c.cx = K[0][2]; c.cy = K[0][3];
        int rows = 0;
        for (int l = 0; l < L; l++) {
            c.w[l] = w[l]; c.h[l] = h[l]; c.mapI[l] = mapI[l]; c.mapW[l] = mapW[l]; c.gray[l] = gray_ref[l]; c.pc[l] = pc[l];
            c.row_count[l] = row_count[l]; c.row_offset[l] = row_offset[l];
            c.row_base[l] = rows; rows += std::max(h[l] - 4, 0);
        }
        for (int l = L; l <= MAXL; l++) c.row_base[l] = rows;
        TCK(cudaMemsetAsync(mapI[0], 0, (size_t) W * H * 8, stream));
        TCK(cudaMemsetAsync(mapW[0], 0, (size_t) W * H * 8, stream));
        if (P) { cd_project_kernel<<<(P + 255) / 256, 256, 0, stream>>>(c, P, (const double *) (d_pts + o_rel), (const int *) (d_pts + o_fr), (const float2 *) (d_pts + o_xy),
                                                                         (const double *) (d_pts + o_id), (const double *) (d_pts + o_un)); launches++; }
        for (int l = 1; l < L; l++) { cd_downsum_kernel<<<std::min(592, (w[l] * h[l] + 255) / 256), 256, 0, stream>>>(c, l); launches++; }
        if (rows > 0) { cd_rows_kernel<0><<<rows, 128, 0, stream>>>(c); launches++; }
        cd_rowscan_kernel<<<L, 32, 0, stream>>>(c); launches++;
        if (rows > 0) { cd_rows_kernel<1><<<rows, 128, 0, stream>>>(c); launches++; }
        TCK(cudaGetLastError());
        TCK(cudaStreamSynchronize(stream));       // pageable sources above must stay valid until the copies are done
        have_ref = true;
        return CMLTRK_OK;
    }

    void fill_params(const double *last_rmse) {
        TrackParams &p = params;
        p.max_level = std::min(L - 1, 4);
        for (int l = 0; l < OPTL; l++) {
            const int s = std::min(l, L - 1);
            p.lv[l] = LevelDev{w[s], h[s], (float) K[s][0], (float) K[s][1], (float) K[s][2], (float) K[s][3], pc[s], pc_n + s, (frame_external && ext_grad[s]) ? ext_grad[s] : grad_new[s]};
        }
        p.huber = (float) cfg.huber_threshold; p.cutoff = (float) cfg.cutoff_threshold;
        const float sc[8] = {(float) cfg.scale_rotation, (float) cfg.scale_rotation, (float) cfg.scale_rotation, (float) cfg.scale_translation, (float) cfg.scale_translation,
                             (float) cfg.scale_translation, (float) cfg.scale_light_a, (float) cfg.scale_light_b};
        for (int i = 0; i < 8; i++) p.scale[i] = (double) sc[i];         // the reference's parameters are floats (Parameter::f())
        p.optimize_a = cfg.optimize_a; p.optimize_b = cfg.optimize_b; p.sat_th = cfg.saturated_ratio_threshold;
        p.ref_tau = ref_exposure[0]; p.ref_a = ref_exposure[1]; p.ref_b = ref_exposure[2]; p.new_tau = new_tau;
        p.has_last = last_rmse ? 1 : 0;
        for (int l = 0; l < OPTL; l++) p.last_rmse[l] = last_rmse ? last_rmse[l] : 0.0;
    }

    int launch_track(int Kc) {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(Kc * cluster); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = 0; lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        TCK(cudaLaunchKernelEx(&lc, kernel(), params, (const Candidate *) d_cand, d_out));
        launches++;
        return CMLTRK_OK;
    }

    int optimize(int Kc, const double *cams, const double *aff, const double *last_rmse, cmltrk_result *res, long launches_before) {
        if (!have_ref || !have_frame) { error = "optimize() needs make_coarse_depth() and set_frame() first"; return CMLTRK_ERR_STATE; }
        if (Kc < 1 || Kc > CMLTRK_MAX_CANDIDATES || !cams || !aff || !res) { error = "bad candidate count or NULL argument"; return CMLTRK_ERR_ARG; }
        TCK(cudaSetDevice(device));
        fill_params(last_rmse);
        Candidate *hc = (Candidate *) h_block;
        TrackOut *ho = (TrackOut *) (h_block + sizeof(Candidate) * CMLTRK_MAX_CANDIDATES);
        const Pose ref_inv = cmlba::pose_inv(ref_pose);
        for (int k = 0; k < Kc; k++) {
            Pose c; memcpy(c.R, cams + (size_t) k * 12, 72); memcpy(c.t, cams + (size_t) k * 12 + 9, 24);
            const Pose r2n = cmlba::pose_mul(c, ref_inv);          // reference->getCamera().to(camera)
            memcpy(hc[k].R, r2n.R, 72); memcpy(hc[k].t, r2n.t, 24);
            hc[k].a = aff[2 * k]; hc[k].b = aff[2 * k + 1];
        }
        TCK(cudaMemcpyAsync(d_cand, hc, sizeof(Candidate) * Kc, cudaMemcpyHostToDevice, stream));
        TCK(cudaEventRecord(ev0, stream));
        int rc = launch_track(Kc);
        if (rc) return rc;
        TCK(cudaEventRecord(ev1, stream));
        TCK(cudaMemcpyAsync(ho, d_out, sizeof(TrackOut) * Kc, cudaMemcpyDeviceToHost, stream));
        TCK(cudaStreamSynchronize(stream));
        float ms = 0.f;
        TCK(cudaEventElapsedTime(&ms, ev0, ev1));
        last_K = Kc;
        for (int k = 0; k < Kc; k++) {
            const TrackOut &o = ho[k];
            cmltrk_result &r = res[k];
            memset(&r, 0, sizeof r);
            if (o.is_correct) {
                Pose rel; memcpy(rel.R, o.R, 72); memcpy(rel.t, o.t, 24);
                const Pose cam = cmlba::pose_mul(rel, ref_pose);  // reference->getCamera().compose(refToNew)
                memcpy(r.cam, cam.R, 72); memcpy(r.cam + 9, cam.t, 24);
            } else {
                memcpy(r.cam, cams + (size_t) k * 12, 96);
            }
            r.affine[0] = o.a; r.affine[1] = o.b;
            for (int l = 0; l < OPTL; l++) { r.E[l] = o.E[l]; r.num_terms_in_E[l] = o.nT[l]; r.num_saturated[l] = o.nS[l]; r.num_robust[l] = o.nR[l]; r.level_cutoff_repeat[l] = o.rep[l]; }
            for (int i = 0; i < 3; i++) r.flow_vector[i] = o.flow[i];
            r.rel_aff[0] = o.rel_aff[0]; r.rel_aff[1] = o.rel_aff[1];
            for (int i = 0; i < 6; i++) r.covariance[i] = o.cov[i];
            r.is_correct = o.is_correct; r.too_many_saturated = o.sat_ok; r.iterations = o.iterations;
            r.levels_used = params.max_level + 1; r.gpu_ms = ms; r.kernel_launches = (int) (launches - launches_before);
        }
        return CMLTRK_OK;
    }

    int bench(int repeats, float *ms_out) {
        if (last_K < 1) { error = "bench_optimize() needs a previous optimize()"; return CMLTRK_ERR_STATE; }
        TCK(cudaSetDevice(device));
        TCK(cudaEventRecord(ev0, stream));
        for (int i = 0; i < repeats; i++) { int rc = launch_track(last_K); if (rc) return rc; }
        TCK(cudaEventRecord(ev1, stream));
        TCK(cudaStreamSynchronize(stream));
        float ms = 0.f;
        TCK(cudaEventElapsedTime(&ms, ev0, ev1));
        *ms_out = ms / std::max(repeats, 1);
        return CMLTRK_OK;
    }

    int64_t read(const char *name, void *dst, int64_t cap) {
        std::string n(name);
        auto out = [&](const void *src, size_t bytes, bool dev) -> int64_t {
            if ((int64_t) bytes > cap) { error = "buffer too small for " + n; return CMLTRK_ERR_ARG; }
            if (dev) { if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { error = "cudaMemcpy failed"; return CMLTRK_ERR_CUDA; } }
            else memcpy(dst, src, bytes);
            return (int64_t) bytes;
        };
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { error = "device error"; return CMLTRK_ERR_CUDA; }
        if (n == "pc_n") return out(pc_n, (size_t) L * 4, true);
        if (n == "cycles") {      // of the last optimize(), start pose 0: evaluations, then SM cycles in advance / evaluate / reduce+exchange
            const TrackOut *ho = (const TrackOut *) (h_block + sizeof(Candidate) * CMLTRK_MAX_CANDIDATES);
            const long long v[4] = {ho[0].evals, ho[0].cyc_advance, ho[0].cyc_eval, ho[0].cyc_reduce};
            return out(v, sizeof v, false);
        }
        if (n == "levels_wh") { int v[2 * MAXL]; for (int l = 0; l < L; l++) { v[2 * l] = w[l]; v[2 * l + 1] = h[l]; } return out(v, (size_t) L * 8, false); }
        if (n == "K") return out(K, (size_t) L * 32, false);
        if (n.size() >= 3 && n.compare(0, 2, "pc") == 0 && isdigit(n[2])) {
            const int l = n[2] - '0';
            if (l >= L) { error = "no such level"; return CMLTRK_ERR_ARG; }
            int cnt = 0;
            if (cudaMemcpy(&cnt, pc_n + l, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { error = "cudaMemcpy failed"; return CMLTRK_ERR_CUDA; }
            return out(pc[l], (size_t) cnt * 16, true);
        }
        if (n.size() >= 5 && n.compare(0, 4, "grad") == 0 && isdigit(n[4])) {
            const int l = n[4] - '0';
            if (l >= L) { error = "no such level"; return CMLTRK_ERR_ARG; }
            return out(grad_new[l], (size_t) w[l] * h[l] * 16, true);
        }
        error = "unknown buffer " + n;
        return CMLTRK_ERR_ARG;
    }
};

}  // namespace cmltrk

using cmltrk::Tracker;

extern "C" {

void cmltrk_default_config(cmltrk_config *c) {
    if (!c) return;
    c->huber_threshold = 9.0; c->cutoff_threshold = 20.0;
    c->scale_rotation = 1.0; c->scale_translation = 0.5; c->scale_light_a = 10.0; c->scale_light_b = 1000.0;
    c->optimize_a = 1; c->optimize_b = 1; c->saturated_ratio_threshold = 0.33;
    c->levels = 0; c->cluster_ctas = 16; c->cta_threads = 384;
}

int cmltrk_create(const cmltrk_config *cfg, int device, int width, int height, double fx, double fy, double cx, double cy, cmltrk_handle *out) {
    if (!out) { cmltrk::g_create_error = "out is NULL"; return CMLTRK_ERR_ARG; }
    *out = nullptr;
    cmltrk_config c;
    if (cfg) c = *cfg; else cmltrk_default_config(&c);
    Tracker *t = new Tracker();
    const int rc = t->create(c, device, width, height, fx, fy, cx, cy);
    if (rc) { cmltrk::g_create_error = t->error; delete t; return rc; }
    *out = reinterpret_cast<cmltrk_handle>(t);
    return CMLTRK_OK;
}

void cmltrk_destroy(cmltrk_handle h) { delete reinterpret_cast<Tracker *>(h); }

const char *cmltrk_last_error(cmltrk_handle h) { return h ? reinterpret_cast<Tracker *>(h)->error.c_str() : cmltrk::g_create_error.c_str(); }

int cmltrk_make_coarse_depth(cmltrk_handle h, const float *ref_gray, const double ref_cam[12], const double ref_exposure[3], int num_frames, const double *frame_cams,
                             int num_points, const int32_t *pt_frame, const float *pt_xy, const double *pt_idepth, const double *pt_uncertainty) {
    if (!h) return CMLTRK_ERR_ARG;
    return reinterpret_cast<Tracker *>(h)->make_coarse_depth(ref_gray, ref_cam, ref_exposure, num_frames, frame_cams, num_points, pt_frame, pt_xy, pt_idepth, pt_uncertainty);
}

int cmltrk_make_coarse_depth_device(cmltrk_handle h, int levels, const float *const *d_gray_levels, const double ref_cam[12], const double ref_exposure[3], int num_frames,
                                    const double *frame_cams, int num_points, const int32_t *pt_frame, const float *pt_xy, const double *pt_idepth, const double *pt_uncertainty) {
    if (!h || !d_gray_levels) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    if (levels < t->L) { t->error = "need every pyramid level of the reference keyframe"; return CMLTRK_ERR_ARG; }
    t->ref_levels_on_device = true;
    const int rc = t->make_coarse_depth(reinterpret_cast<const float *>(d_gray_levels), ref_cam, ref_exposure, num_frames, frame_cams, num_points, pt_frame, pt_xy, pt_idepth,
                                        pt_uncertainty);
    t->ref_levels_on_device = false;
    return rc;
}

int cmltrk_set_frame_device(cmltrk_handle h, int levels, const void *const *d_texel_levels, double exposure_time) {
    if (!h) return CMLTRK_ERR_ARG;
    return reinterpret_cast<Tracker *>(h)->set_frame_device(levels, d_texel_levels, exposure_time);
}

int cmltrk_set_frame(cmltrk_handle h, const float *gray, double exposure_time) {
    if (!h) return CMLTRK_ERR_ARG;
    return reinterpret_cast<Tracker *>(h)->set_frame(gray, exposure_time);
}

int cmltrk_optimize(cmltrk_handle h, int num_candidates, const double *start_cams, const double *start_affine, const double *last_rmse, cmltrk_result *results) {
    if (!h) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    return t->optimize(num_candidates, start_cams, start_affine, last_rmse, results, t->launches);
}

int cmltrk_track(cmltrk_handle h, const float *gray, double exposure_time, int num_candidates, const double *start_cams, const double *start_affine, const double *last_rmse,
                 cmltrk_result *results) {
    if (!h) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    const long before = t->launches;
    const int rc = t->set_frame(gray, exposure_time);
    if (rc) return rc;
    return t->optimize(num_candidates, start_cams, start_affine, last_rmse, results, before);
}

// DSOTracker::trackWithMotionModel (DSOTracker.h:240-360); mirrors libcml_b200/tracker.py::trackWithMotionModelPy step by step
static double trk_rmse(const cmltrk_result &r, int l) { return r.num_terms_in_E[l] > 0 ? r.E[l] / r.num_terms_in_E[l] : 0.0; }
int cmltrk_track_with_motion_model(cmltrk_handle h, int num_cameras, const double *cameras, const double initial_affine[2], int failure_mode, int *ok, cmltrk_result *best_out,
                                   int *tried) {
    if (!h || !cameras || !initial_affine || !ok || !best_out || num_cameras < 1) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    bool have = false, have_best = false;
    cmltrk_result best; memset(&best, 0, sizeof(best));
    double achieved = INFINITY;
    int n_tried = 0;
    auto last_of = [&](double *last) -> const double * {                    // mLastResidual = trackingResult: the rmse gate of the next optimize
        if (!have_best || !best.is_correct) return nullptr;
        for (int l = 0; l < CMLTRK_OPT_LEVELS; l++) last[l] = l < best.levels_used ? trk_rmse(best, l) : 0.0;
        return last;
    };
    for (int i = 0; i < num_cameras; i++) {
        n_tried = i + 1;
        double last[CMLTRK_OPT_LEVELS];
        cmltrk_result test;
        const int rc = t->optimize(1, cameras + 12 * i, initial_affine, last_of(last), &test, t->launches);
        if (rc) return rc;
        const double test_rmse = test.num_terms_in_E[0] > 0 ? test.E[0] / test.num_terms_in_E[0] : INFINITY;
        const bool test_ok = test.is_correct && test.num_terms_in_E[0] > 0 && std::isfinite(test_rmse);
        const bool best_sat = have_best ? best.too_many_saturated != 0 : true;                     // Residual(): tooManySaturated = true
        if (best_sat && !test.too_many_saturated && test_ok) { have = true; best = test; have_best = true; }
        if (test_ok && !(test_rmse >= achieved)) {
            if ((have_best ? best.too_many_saturated != 0 : true) || !test.too_many_saturated) { have = true; best = test; have_best = true; }
        }
        if (have && test.num_terms_in_E[0] > 0 && test_rmse < achieved) achieved = test_rmse;
        if (have && achieved < t->last_coarse_rmse * 1.5) break;
        if (have && i >= 50) break;
    }
    if (tried) *tried = n_tried;
    if (!have) {
        if (failure_mode != 1) { *ok = 0; *best_out = best; return CMLTRK_OK; }
        double last[CMLTRK_OPT_LEVELS];
        const int rc = t->optimize(1, cameras, initial_affine, last_of(last), &best, t->launches);
        if (rc) return rc;
        *ok = 1; *best_out = best;
        return CMLTRK_OK;
    }
    t->last_coarse_rmse = achieved;
    if (t->first_rmse < 0) t->first_rmse = achieved;
    *ok = 1; *best_out = best;
    return CMLTRK_OK;
}
int cmltrk_reset_motion_model(cmltrk_handle h) {
    if (!h) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    t->last_coarse_rmse = INFINITY; t->first_rmse = -1.0;
    return CMLTRK_OK;
}
int cmltrk_motion_model_state(cmltrk_handle h, double *last_coarse_rmse, double *first_rmse) {
    if (!h) return CMLTRK_ERR_ARG;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    if (last_coarse_rmse) *last_coarse_rmse = t->last_coarse_rmse;
    if (first_rmse) *first_rmse = t->first_rmse;
    return CMLTRK_OK;
}

float *cmltrk_frame_buffer(cmltrk_handle h) {
    if (!h) return nullptr;
    Tracker *t = reinterpret_cast<Tracker *>(h);
    return (float *) (t->h_block + t->stage_gray);
}

int64_t cmltrk_read(cmltrk_handle h, const char *name, void *dst, int64_t capacity) {
    if (!h || !name || !dst) return CMLTRK_ERR_ARG;
    return reinterpret_cast<Tracker *>(h)->read(name, dst, capacity);
}

int cmltrk_bench_optimize(cmltrk_handle h, int repeats, float *ms_per_launch) {
    if (!h || !ms_per_launch || repeats < 1) return CMLTRK_ERR_ARG;
    return reinterpret_cast<Tracker *>(h)->bench(repeats, ms_per_launch);
}

}  // extern "C"
