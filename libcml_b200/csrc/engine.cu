// engine.cu -- host side of libcmlba.so: window bookkeeping (the DSOContext role), device buffers,
// the run() kernel sequence and the C ABI of include/cmlba.h.
//
// Host-side reference anchors (under /root/reference/src/cml/optimization/dso):
//   Engine::add_frame   DSOBundleAdjustment.cpp:417-462 (addNewFrame), DSOFrame.h:98-106 (setEvalPT_scaled)
//   Engine::add_points  DSOBundleAdjustment.cpp:382-415 (addPoints), :336-380 (createResidual), DSOContext.h:76-91
//   Engine::prepare     DSOBundleAdjustment.cpp:753-782 (run prologue), :1030-1101 (computeAdjoints), :1103-1194
//                       (computeDelta priors), :2365-2417 (computeNullspaces), :1196-1250 (nullspace projector)
//   Engine::run         DSOBundleAdjustment.cpp:744-910
//   Engine::finish_run  DSOBundleAdjustment.cpp:1613-1642 (residual dropping, outliers), DSOPoint.h:107-118
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cmlba.h"
#include "kernels.cuh"

namespace cmlba {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(_e));                         \
            return CMLBA_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

// host-side stage timers (cmlba_read "host_timing"): where an end-to-end call spends its wall time
struct HostTimers {
    std::map<std::string, std::pair<double, long>> acc;   // name -> (ms, calls)
    struct Scope {
        HostTimers &t; const char *name; std::chrono::steady_clock::time_point t0;
        Scope(HostTimers &t_, const char *n) : t(t_), name(n), t0(std::chrono::steady_clock::now()) {}
        ~Scope() { auto &e = t.acc[name]; e.first += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); e.second++; }
    };
    std::string text() const { std::string s; char b[160]; for (auto &kv : acc) { snprintf(b, sizeof b, "%-28s %10.3f ms %6ld calls\n", kv.first.c_str(), kv.second.first, kv.second.second); s += b; } return s; }
};
struct Lap {   // sequential section timer: lap("name") charges the time since the previous lap
    HostTimers &t; std::chrono::steady_clock::time_point last;
    explicit Lap(HostTimers &t_) : t(t_), last(std::chrono::steady_clock::now()) {}
    void operator()(const char *name) { auto now = std::chrono::steady_clock::now(); auto &e = t.acc[name]; e.first += std::chrono::duration<double, std::milli>(now - last).count(); e.second++; last = now; }
};
#define TSCOPE(name) HostTimers::Scope _ts_##__LINE__(timers, name)

static long g_dev_mallocs = 0, g_pin_mallocs = 0;   // allocation counters (cmlba_read "host_timing")
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    bool view = false;            // p points into an UploadArena: nothing to free
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p && !view) cudaFree(p);
        p = nullptr; cap = 0; view = false;
        size_t want = n + n / 4 + 16;
        g_dev_mallocs++;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; view = false; }
};

// page-locked host staging (async H2D/D2H without the driver's bounce buffer)
template <typename T> struct PinnedBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 16;
        g_pin_mallocs++;
        cudaError_t e = cudaHostAlloc((void **) &p, want * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One pinned host block mirrored by one device block: the window arrays are laid out once (256-byte aligned),
// filled in place on the host and uploaded with a single copy.
struct UploadArena {
    PinnedBuf<char> h; DevBuf<char> d;
    size_t used = 0;
    void begin() { used = 0; }
    template <typename T> size_t take(size_t count) { const size_t off = (used + 255) & ~(size_t) 255; used = off + std::max<size_t>(count, 1) * sizeof(T); return off; }
    cudaError_t commit() { cudaError_t e = h.reserve(used); if (e != cudaSuccess) return e; return d.reserve(used); }
    template <typename T> T *host(size_t off) { return reinterpret_cast<T *>(h.p + off); }
    template <typename T> T *dev(size_t off) { return reinterpret_cast<T *>(d.p + off); }
};

// open-addressing map point id -> index (rebuilt wholesale on compaction; std::unordered_map cost ~1 ms per 16k points)
struct IdMap {
    std::vector<int64_t> keys; std::vector<int> vals; size_t n = 0, mask = 0;
    static constexpr int64_t EMPTY = INT64_MIN;
    static size_t hash(int64_t k) { uint64_t x = (uint64_t) k * 0x9E3779B97F4A7C15ull; return (size_t) (x ^ (x >> 29)); }
    void clear() { keys.clear(); vals.clear(); n = 0; mask = 0; }
    void rehash(size_t cap) {
        std::vector<int64_t> ok; std::vector<int> ov; ok.swap(keys); ov.swap(vals);
        keys.assign(cap, EMPTY); vals.assign(cap, -1); mask = cap - 1; n = 0;
        for (size_t i = 0; i < ok.size(); i++) if (ok[i] != EMPTY) set(ok[i], ov[i]);
    }
    void reserve(size_t want) { size_t cap = 64; while (cap < want * 2) cap <<= 1; if (cap > keys.size()) rehash(cap); }
    int find(int64_t k) const {
        if (keys.empty()) return -1;
        for (size_t i = hash(k) & mask;; i = (i + 1) & mask) { if (keys[i] == k) return vals[i]; if (keys[i] == EMPTY) return -1; }
    }
    // returns the stored value if k is present, else inserts (k, v) and returns -1
    int find_or_insert(int64_t k, int v) {
        if ((n + 1) * 2 > keys.size()) rehash(std::max<size_t>(64, keys.size() * 2));
        for (size_t i = hash(k) & mask;; i = (i + 1) & mask) { if (keys[i] == k) return vals[i]; if (keys[i] == EMPTY) { keys[i] = k; vals[i] = v; n++; return -1; } }
    }
    void set(int64_t k, int v) {
        if ((n + 1) * 2 > keys.size()) rehash(std::max<size_t>(64, keys.size() * 2));
        for (size_t i = hash(k) & mask;; i = (i + 1) & mask) { if (keys[i] == k) { vals[i] = v; return; } if (keys[i] == EMPTY) { keys[i] = k; vals[i] = v; n++; return; } }
    }
};

struct FrameHost {
    int64_t id = 0;
    Pose evalpt;                 // worldToCam_evalPT
    double state[10] = {0}, state_zero[10] = {0};
    double prior[8] = {0};
    double exposure = 1.0;
    float energy_th = 8 * 8 * 8; // DSOFrame.h:35
    int keyid = 0;
    bool is_init = false;
    float4 *d_img = nullptr;
    bool flagged = false;        // flaggedForMarginalization
    int num_marginalized = 0, num_residuals_out = 0;   // DSOFrame counters fed by DSOContext::removePoint / removeResiduals
    Pose pre;                    // PRE_worldToCam (last known)
    double aff_a = 0, aff_b = 0; // aff_g2l (scaled a,b)
};

struct PointHost {
    int64_t id = 0;
    int64_t host_id = 0;
    int host = 0;                // index of the host frame in frames_
    uint16_t res_mask = 0;       // bit t: residual to frames_[t]
    float x = 0, y = 0;
    double idepth = 0;
    float idepth_zero = 0;
    bool has_prior = false;
    int num_good = 0;
    float max_rel_bs = 0, idepth_hessian = 0;
    int64_t last_frame[2] = {-1, -1};   // frames of lastResiduals[0/1] (DSOPoint.h:119-156); -1 = nullptr
    int last_state[2] = {CMLBA_RES_OOB, CMLBA_RES_OOB};
    bool alive = true;
    bool to_marginalize = false; // DSOTOMARGINALIZE
};

// Residual bookkeeping: a residual (point, target frame) exists iff bit `target slot` of PointHost::res_mask is set
// (DSOContext keeps a pointer set per point, DSOContext.h:58-75).  State/energy of the last run() live in a
// device-order snapshot (ResSnapshot) that cmlba_get_residuals decodes on demand.
struct ResSnapshot {
    bool valid = false;
    bool own_map = false;                          // r_point / r_target copied out of the upload arena (done lazily, before the arena is rebuilt)
    int R = 0;
    std::vector<int64_t> frame_id, point_id;       // frame ids by slot, point ids by device position, at run() time
    std::vector<int> r_point;                      // device point position per residual
    std::vector<uint8_t> r_target;
    const uint8_t *state = nullptr, *alive = nullptr;   // into the pinned read-back block of the last finish_run
    const float *energy = nullptr;
};

// --- NCCL through dlopen (multi-GPU only) -----------------------------------------------------
struct NcclUid { char b[128]; };
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = "cannot dlopen libnccl.so.2 (add torch's nvidia/nccl/lib to LD_LIBRARY_PATH or import torch first)"; return false; }
        GetUniqueId = (decltype(GetUniqueId)) dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank)) dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce)) dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather)) dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy)) dlsym(lib, "ncclCommDestroy");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
static NcclApi g_nccl;

class Engine {
public:
    cmlba_config cfg;
    int device = 0;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_calib = false;
    double fx = 0, fy = 0, cx = 0, cy = 0;
    int W = 0, H = 0;
    std::vector<FrameHost> frames_;
    std::vector<PointHost> points_;   // dead points stay in place (alive = false) until a quarter of the slots is dead
    size_t n_dead = 0;
    ResSnapshot snap;
    IdMap point_index_;
    // point ids usually arrive in increasing order (Map::createMapPoint counts up): then points_ is sorted by id, look-ups are a binary search and
    // add_points skips the hash inserts (17 ns per point); the first id that breaks the order switches to the hash index for good (until reset)
    bool ids_sorted = true;
    std::vector<int64_t> outliers_;
    int key_counter = 0;
    bool dirty = true;            // device window must be rebuilt
    bool prepared = false;
    std::vector<double> HM, bM;   // marginalisation prior mMarginalizedHessian / mMarginalizedB, (8N+4)^2 and 8N+4 (only kept when !disable_marginalization)
    bool hm_frame_done = false;           // marginalize_frames already folded the frame into H_M (remove_frame must not drop the block again)
    bool subset_to_marginalize = false;   // build_device_window keeps only the points marked DSOTOMARGINALIZE (marginalize_points)
    // device-window layout (host mirrors)
    std::vector<int> pt_order;    // device point i -> points_ index
    DevWin dw = {};
    int launches = 0;
    // multi-GPU
    void *comm = nullptr; int rank = 0, world = 1;
    DevBuf<double> d_post_send, d_post_recv; DevBuf<int> d_cap;   // multi-GPU post-linearize exchange
    // reduced-system exchange over NVLink peer memory (cudaIpc); falls back to ncclAllReduce when not opened
    char *p2p_local = nullptr; char *p2p_peer[MAXF] = {nullptr}; bool p2p_ready = false; unsigned long long p2p_epoch = 0;
    size_t p2p_bytes() const { return 256 + (size_t) 2 * world * (P2P_SLOT_DOUBLES + P2P_POST_DOUBLES) * sizeof(double); }   // flags | sys slots[2][world] | post records[2][world]
    unsigned long long p2p_post_epoch = 0;

    // device buffers
    DevBuf<FrameDev> d_frames; DevBuf<PairPre> d_pairs; DevBuf<Ctrl> d_ctrl;
    DevBuf<double> d_AH, d_AT, d_HM, d_bM, d_Pns, d_pt_idepth, d_pt_step, d_energy_part, d_st_out, d_sys, d_x, d_xAd, d_pt_part;
    DevBuf<int> d_pt_host, d_pt_num_good, d_pt_ngood_cur, d_r_point, d_res_bin_begin, d_sc_chunk_host,
        d_sc_chunk_begin, d_sc_chunk_count, d_host_chunk_begin;
    // tile binning (linearize.cuh): sorted residual order, tile jobs, partial-block bookkeeping, final states in host order
    DevBuf<int> d_bin_key, d_bin_hist, d_bin_offs, d_job_of_tile, d_r_job, d_r_src, d_bin_ticket, d_job_begin, d_cta_info;
    DevBuf<uint32_t> d_job_desc, d_r_pht;
    DevBuf<uint8_t> d_fin_state, d_fin_alive;
    DevBuf<float> d_fin_energy;
    DevBuf<float4> d_r_pt4;
    DevBuf<uint8_t> d_bin_g;
    DevBuf<long long> d_lt_trace;
    DevBuf<unsigned long long> d_ktrace;   // development timeline (CMLBA_KTRACE=1)
    double stats_[19] = {0};      // latest value of every Statistic series of BA.h:215-233 (cmlba_get_statistics)
    bool want_ktrace = false;
    int kt_seq = 0; std::vector<int> kt_sites;   // launch slots of the timeline (slot 0: the reset)
    TileMaps tile_maps;           // one tensor map per window frame (re-encoded by build_device_window)
    int n_sm = 148;
    int lt_variant = 0, lt_mode = 0, lt_exact = 0;
    bool use_pdl = true;
    int launch_rc = CMLBA_OK;     // sticky status of the enqueue helpers (kernel launch / NCCL call failed); run() and the stage calls report it
    int pdl_mask = 316;            // development (CMLBA_PDL_MASK): 1 linearize 2 post 4 accumulate 8 schur 16 stitch 32 assemble 64 solve 128 point step 256 peer all-reduce
    int tail_cluster_max = 0;     // largest cluster tail_kernel can be scheduled with (0: fused tail unavailable -> schur / stitch_pair / assemble)   // development switches (CMLBA_LT_VARIANT, CMLBA_LT_MODE): kernel shape, streaming-only mode
    DevBuf<float> d_pt_x, d_pt_y, d_pt_idz, d_pt_idb, d_pt_colors, d_pt_weights, d_pt_priorF, d_pt_Hdd, d_pt_bd, d_pt_Hcd, d_pt_HdiF, d_pt_bdSumF, d_pt_idh, d_pt_mrb,
        d_r_energy0, d_r_energy1, d_r_new_energy, d_r_new_energy_wo, d_r_center, d_rj0, d_rj1, d_T0, d_T1, d_dbg, d_acc_bin, d_sc_part, d_stage[MAXF];
    int stage_flip = 0;
    cudaEvent_t ev_copy = nullptr;
    // accumulate_kernel (Jacobian records -> 13x13 blocks) and schur_kernel (Schur rows -> per-chunk blocks) are independent consumers of one
    // linearization: the first runs on a side stream forked off / joined back into the main one (CMLBA_NO_FORK=1 serialises them)
    cudaStream_t side_stream = nullptr, launch_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool use_fork = true;
    int acc_jobs_launched = 0;     // accumulate jobs launched since prepare() (= the value Ctrl.acc_done_count reaches), see launch_schur
    bool side_pending = false;     // work on the side stream that the main stream has not joined yet
    UploadArena up;                               // every array build_device_window uploads
    UploadArena prep;                             // per-run constants uploaded by prepare()
    std::vector<int> res_bin_begin;               // [N*N+1] first device residual of every bin (t*N+h)
    size_t up_o_rp = 0, up_o_rt = 0;              // arena offsets of r_point / r_target (host mirrors stay valid until the next build)
    PinnedBuf<char> fin_h;                        // finish_run read-back
    PinnedBuf<int> peek_h;                        // Ctrl.done peeked between iteration batches
    DevBuf<uint8_t> d_r_host, d_r_target, d_r_state0, d_r_state1, d_r_good0, d_r_good1, d_r_new_state, d_r_alive;
    bool want_dbg = false;
    DevBuf<float4> d_flush;
    std::vector<float4 *> img_pool;   // recycled image allocations (reset())
    int n_chunks = 0, n_sc_chunks = 0;
    std::vector<int> h_host_chunk_begin;

    HostTimers timers;
    void set_error(const std::string &s) { err = s; }

    int init() {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libcmlba has no CPU fallback)"); return CMLBA_ERR_CUDA; }
        if (device < 0 || device >= ndev) { set_error("bad device ordinal"); return CMLBA_ERR_ARG; }
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1)); CK(cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        launch_stream = stream;
        if (getenv("CMLBA_NO_FORK")) use_fork = false;
        if (getenv("CMLBA_KTRACE")) want_ktrace = true;
        CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
#define LT_ATTR(CW, ST)                                                                                                                                  \
    CK(cudaFuncSetAttribute(linearize_tile_kernel<false, CW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::min<size_t>(lt_smem_bytes(MAXF, CW, ST), 227 * 1024))); \
    CK(cudaFuncSetAttribute(linearize_tile_kernel<true, CW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::min<size_t>(lt_smem_bytes(MAXF, CW, ST), 227 * 1024)));
        LT_ATTR(12, 4) LT_ATTR(8, 4) LT_ATTR(16, 3) LT_ATTR(12, 3)   // warps are allocated in groups of four: 8 / 12 / 16 warps per CTA incl. the producer
#undef LT_ATTR
        // fused tail (tail.cuh): one cluster per host frame, one CTA per frame -> clusters of 8 (portable) or 16 (opt-in) CTAs
        if (getenv("CMLBA_TAIL_FUSION") && cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) tail_smem_bytes(MAXF)) == cudaSuccess) {
            tail_cluster_max = 8;
            if (cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
                cudaLaunchConfig_t lc = {}; cudaLaunchAttribute at[1];
                lc.gridDim = dim3(16); lc.blockDim = dim3(TAIL_THREADS); lc.dynamicSmemBytes = tail_smem_bytes(MAXF);
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1;
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, tail_kernel, &lc) == cudaSuccess && nc >= 1) tail_cluster_max = 16;
            }
        }
        cudaGetLastError();
        if (const char *v = getenv("CMLBA_LT_VARIANT")) lt_variant = atoi(v);
        if (const char *v = getenv("CMLBA_LT_MODE")) lt_mode = atoi(v);
        if (getenv("CMLBA_NO_PDL")) use_pdl = false;
        if (const char *v = getenv("CMLBA_PDL_MASK")) pdl_mask = atoi(v);
        if (const char *v = getenv("CMLBA_LT_EXACT")) lt_exact = atoi(v);
        CK(cudaFuncSetAttribute(bin_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        CK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        return CMLBA_OK;
    }

    ~Engine() {
        cudaSetDevice(device);
        for (auto &f : frames_) if (f.d_img) cudaFree(f.d_img);
        for (auto *p : img_pool) cudaFree(p);
        d_flush.release();
        if (stream) cudaStreamDestroy(stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_copy) cudaEventDestroy(ev_copy);
        if (side_stream) cudaStreamDestroy(side_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        up.h.release(); up.d.release(); prep.h.release(); prep.d.release(); fin_h.release(); peek_h.release();
        for (int q = 0; q < MAXF; q++) if (p2p_peer[q] && p2p_peer[q] != p2p_local) cudaIpcCloseMemHandle(p2p_peer[q]);
        if (p2p_local) cudaFree(p2p_local);
        if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
        // DevBuf members leak-free:
        DevBuf<double> *dd[] = {&d_AH, &d_AT, &d_HM, &d_bM, &d_Pns, &d_pt_idepth, &d_pt_step, &d_energy_part, &d_st_out, &d_sys, &d_x, &d_xAd, &d_pt_part};
        for (auto *b : dd) b->release();
        d_post_send.release(); d_post_recv.release(); d_cap.release();
        DevBuf<int> *di[] = {&d_pt_host, &d_pt_num_good, &d_pt_ngood_cur, &d_r_point, &d_res_bin_begin, &d_sc_chunk_host, &d_sc_chunk_begin, &d_sc_chunk_count, &d_host_chunk_begin,
                             &d_bin_key, &d_bin_hist, &d_bin_offs, &d_job_of_tile, &d_r_job, &d_r_src, &d_bin_ticket, &d_job_begin, &d_cta_info};
        for (auto *b : di) b->release();
        d_job_desc.release(); d_r_pht.release(); d_r_pt4.release(); d_bin_g.release(); d_ktrace.release(); d_fin_state.release(); d_fin_alive.release(); d_fin_energy.release();
        DevBuf<float> *df[] = {&d_pt_x, &d_pt_y, &d_pt_idz, &d_pt_idb, &d_pt_colors, &d_pt_weights, &d_pt_priorF, &d_pt_Hdd, &d_pt_bd, &d_pt_Hcd, &d_pt_HdiF, &d_pt_bdSumF, &d_pt_idh, &d_pt_mrb,
                               &d_r_energy0, &d_r_energy1, &d_r_new_energy, &d_r_new_energy_wo, &d_r_center, &d_rj0, &d_rj1, &d_T0, &d_T1, &d_dbg, &d_acc_bin, &d_sc_part};
        for (auto *b : df) b->release();
        for (auto &b : d_stage) b.release();
        DevBuf<uint8_t> *du[] = {&d_r_host, &d_r_target, &d_r_state0, &d_r_state1, &d_r_good0, &d_r_good1, &d_r_new_state, &d_r_alive};
        for (auto *b : du) b->release();
        d_frames.release(); d_pairs.release(); d_ctrl.release();
    }

    int frame_index(int64_t id) const {
        for (size_t i = 0; i < frames_.size(); i++) if (frames_[i].id == id) return (int) i;
        return -1;
    }

    void scales(double *sc) const {
        const double v[10] = {cfg.scale_translation, cfg.scale_translation, cfg.scale_translation, cfg.scale_rotation, cfg.scale_rotation, cfg.scale_rotation,
                              cfg.scale_light_a, cfg.scale_light_b, cfg.scale_light_a, cfg.scale_light_b};
        for (int k = 0; k < 10; k++) sc[k] = v[k];
    }

    // ------------------------------------------------------------------ window maintenance
    int set_calib(double fx_, double fy_, double cx_, double cy_, int w, int h) {
        if (w < 8 || h < 8 || fx_ <= 0 || fy_ <= 0) { set_error("bad calibration"); return CMLBA_ERR_ARG; }
        if (have_calib && (w != W || h != H) && !frames_.empty()) { set_error("image size change with frames in the window"); return CMLBA_ERR_STATE; }
        if (w != W || h != H) { cudaSetDevice(device); for (auto *p : img_pool) cudaFree(p); img_pool.clear(); }
        fx = fx_; fy = fy_; cx = cx_; cy = cy_; W = w; H = h; have_calib = true; dirty = true;
        return CMLBA_OK;
    }

    // gray != 0: `grad` is the rectified level-0 gray image (W*H floats); the derivative image is built on the device
    int add_frame(int64_t id, const double *w2c, double a, double b, double exposure, const float *grad, int is_init, int gray = 0) {
        TSCOPE("add_frame");
        if (!have_calib) { set_error("cmlba_set_calib must be called before cmlba_add_frame"); return CMLBA_ERR_STATE; }
        if (!w2c || !grad) { set_error("null pointer"); return CMLBA_ERR_ARG; }
        if ((int) frames_.size() >= MAXF) { set_error("window full (CMLBA_MAX_FRAMES)"); return CMLBA_ERR_ARG; }
        for (auto &f : frames_) if (id <= f.id) { set_error("frame ids must increase (DSOContext.h:139-143 aborts here)"); return CMLBA_ERR_ARG; }
        if (exposure <= 0) { set_error("exposure must be > 0"); return CMLBA_ERR_ARG; }
        CK(cudaSetDevice(device));
        FrameHost f;
        f.id = id;
        for (int k = 0; k < 9; k++) f.evalpt.R[k] = w2c[k];
        for (int k = 0; k < 3; k++) f.evalpt.t[k] = w2c[9 + k];
        f.pre = f.evalpt;
        double sc[10]; scales(sc);
        // setEvalPT_scaled: state_scaled = (0,..,a,b), state = scaled / scale, state_zero = state
        f.state[6] = a / sc[6]; f.state[7] = b / sc[7];
        for (int k = 0; k < 10; k++) f.state_zero[k] = f.state[k];
        f.aff_a = a; f.aff_b = b;
        f.exposure = exposure;
        f.keyid = key_counter++;
        f.is_init = is_init != 0;
        // image: AoS (I,dx,dy) -> float4 texels on the device.  The caller's buffer is only ours for the duration
        // of the call: wait for the H2D copy (event), the repack kernel stays asynchronous on the stream.
        const size_t npix = (size_t) W * H;
        // staging buffers: with synchronous uploads two alternate (the previous repack may still run); asynchronous uploads
        // keep one per window slot, because several copies are in flight
        const int sb = cfg.async_image_upload ? (int) frames_.size() : (stage_flip ^= 1);
        const size_t nfl = gray ? npix : npix * 3;
        if (!img_pool.empty()) { f.d_img = img_pool.back(); img_pool.pop_back(); } else CK(cudaMalloc(&f.d_img, npix * sizeof(float4)));
        if (gray == 2) {          // `grad` is a DEVICE pointer to float4 (I, dx, dy, *) texels of this device (cmlimg's level 0): one device-to-device copy
            CK(cudaMemcpyAsync(f.d_img, grad, npix * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
            CK(cudaEventRecord(ev_copy, stream));
        } else {
            CK(d_stage[sb].reserve(nfl));
            CK(cudaMemcpyAsync(d_stage[sb].p, grad, nfl * sizeof(float), cudaMemcpyHostToDevice, stream));
            CK(cudaEventRecord(ev_copy, stream));
            if (gray) gradient_texel_kernel<<<dim3((unsigned) ((W + 127) / 128), (unsigned) H), 128, 0, stream>>>(d_stage[sb].p, f.d_img, W, H);
            else repack_image_kernel<<<(unsigned) ((npix + 255) / 256), 256, 0, stream>>>(d_stage[sb].p, f.d_img, (int) npix);
            CK(cudaGetLastError());
        }
        const int slot = (int) frames_.size();
        frames_.push_back(f);
        if (keep_prior()) hm_grow();
        // residuals from all existing points to the new frame (BA:455-460); lastResiduals slot 0 (BA:374-375)
        for (size_t p = 0; p < points_.size(); p++) {
            if (!points_[p].alive) continue;
            points_[p].res_mask |= (uint16_t) (1u << slot);
            points_[p].last_frame[1] = points_[p].last_frame[0]; points_[p].last_state[1] = points_[p].last_state[0];
            points_[p].last_frame[0] = id; points_[p].last_state[0] = CMLBA_RES_IN;
        }
        if (!cfg.async_image_upload || gray == 2) CK(cudaEventSynchronize(ev_copy));     // a device source may be overwritten by its owner right after the call
        dirty = true; prepared = false;
        return CMLBA_OK;
    }

    int add_points(int n, const int64_t *pid, const int64_t *host_id, const float *xy, const double *idepth) {
        TSCOPE("add_points");
        Lap lap(timers);
        if (n < 0 || (n > 0 && (!pid || !host_id || !xy || !idepth))) { set_error("null pointer"); return CMLBA_ERR_ARG; }
        if (frames_.empty()) { set_error("no frames in the window"); return CMLBA_ERR_STATE; }
        CK(cudaSetDevice(device));
        const size_t first = points_.size();
        const int NF = (int) frames_.size();
        points_.reserve(first + n);
        if (ids_sorted) {          // still sorted after this batch?
            int64_t prev = first ? points_[first - 1].id : INT64_MIN;
            bool inc = true;
            for (int i = 0; i < n && inc; i++) { inc = pid[i] > prev; prev = pid[i]; }
            if (!inc) { ids_sorted = false; reindex(); }
        }
        if (!ids_sorted) point_index_.reserve(first + n);
        int64_t last_hid = INT64_MIN; int last_h = -1;
        for (int i = 0; i < n; i++) {
            if (pid[i] == IdMap::EMPTY) { set_error("point id INT64_MIN is reserved"); points_.resize(first); reindex(); return CMLBA_ERR_ARG; }
            if (!ids_sorted) {   // BA:386-388: a point that is already in the window is skipped (a removed one may come back); increasing ids are new by construction
                const int prev = point_index_.find_or_insert(pid[i], (int) points_.size());
                if (prev >= 0) { if (points_[prev].alive) continue; point_index_.set(pid[i], (int) points_.size()); }
            }
            if (host_id[i] != last_hid) { last_hid = host_id[i]; last_h = frame_index(host_id[i]); }
            const int h = last_h;
            if (h < 0) { set_error("point's host frame is not in the window"); points_.resize(first); reindex(); return CMLBA_ERR_ARG; }
            const float x = xy[2 * i], y = xy[2 * i + 1];
            if (!(x >= 3 && y >= 3 && x < W - 4 && y < H - 4)) { set_error("point closer than 3 px to the image border"); points_.resize(first); reindex(); return CMLBA_ERR_ARG; }
            if (!(idepth[i] > 0) || !std::isfinite(idepth[i])) { set_error("inverse depth must be finite and > 0"); points_.resize(first); reindex(); return CMLBA_ERR_ARG; }
            PointHost p;
            p.id = pid[i]; p.host_id = host_id[i]; p.host = h; p.x = x; p.y = y; p.idepth = idepth[i];
            p.idepth_zero = (float) idepth[i];
            p.has_prior = frames_[h].is_init;
            points_.push_back(p);
        }
        const size_t nn = points_.size() - first;
        lap("addp.validate");
        if (nn == 0) return CMLBA_OK;
        // residual bookkeeping (reference colours and gradient weights of addPoint, DSOContext.h:86-91 / BA:405-411, are computed on
        // the device when the window is built: they only depend on the host image and the pixel)
        const int64_t newest = frames_.back().id, second = NF >= 2 ? frames_[NF - 2].id : -1;
        const uint16_t all = (uint16_t) ((1u << NF) - 1u);
        for (size_t i = 0; i < nn; i++) {
            PointHost &p = points_[first + i];
            p.res_mask = (uint16_t) (all & ~(1u << p.host));      // createResidual towards every other frame (BA:399-403)
            if (p.host != NF - 1) { p.last_frame[0] = newest; p.last_state[0] = CMLBA_RES_IN; }
            if (NF >= 2 && p.host != NF - 2) { p.last_frame[1] = second; p.last_state[1] = CMLBA_RES_IN; }
        }
        lap("addp.residuals");
        dirty = true; prepared = false;
        return CMLBA_OK;
    }

    void materialize_snapshot() {
        if (!snap.valid || snap.own_map) return;
        const int *rp = up.host<int>(up_o_rp); const uint8_t *rt = up.host<uint8_t>(up_o_rt);
        snap.r_point.assign(rp, rp + snap.R); snap.r_target.assign(rt, rt + snap.R);
        snap.own_map = true;
    }
    int find_point(int64_t id) const {
        if (!ids_sorted) return point_index_.find(id);
        size_t lo = 0, hi = points_.size();
        while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (points_[mid].id < id) lo = mid + 1; else hi = mid; }
        return (lo < points_.size() && points_[lo].id == id && points_[lo].alive) ? (int) lo : -1;
    }
    void reindex() {
        if (ids_sorted) return;
        point_index_.clear(); point_index_.reserve(points_.size());
        for (size_t i = 0; i < points_.size(); i++) if (points_[i].alive) point_index_.set(points_[i].id, (int) i);
    }
    // squeeze dead points out and rebuild the id index
    void compact() {
        size_t np = 0;
        for (size_t i = 0; i < points_.size(); i++) if (points_[i].alive) { if (np != i) points_[np] = points_[i]; np++; }
        points_.resize(np);
        n_dead = 0;
        reindex();
        dirty = true; prepared = false;
    }
    void kill_point(size_t idx) { if (points_[idx].alive) { points_[idx].alive = false; n_dead++; } }
    // DSOContext::removePoint(point, marginalize) (DSOContext.h:94-110) -> removeResiduals (:207-218): counters of the target frames
    void drop_point(size_t idx, bool marginalize) {
        PointHost &p = points_[idx];
        if (!p.alive) return;
        for (unsigned m = p.res_mask; m; m &= m - 1) { FrameHost &f = frames_[__builtin_ctz(m)]; f.num_residuals_out++; if (marginalize) f.num_marginalized++; }
        p.res_mask = 0;
        kill_point(idx);
    }
    void maybe_compact() { if (n_dead * 4 > points_.size()) compact(); dirty = true; prepared = false; }

    int remove_point(int64_t id) {
        const int idx = find_point(id);
        if (idx < 0 || !points_[idx].alive) return CMLBA_OK;   // DSOContext.h:95-97
        drop_point(idx, false);
        maybe_compact();
        return CMLBA_OK;
    }

    int remove_frame(int64_t id) {
        const int fi = frame_index(id);
        if (fi < 0) { set_error("unknown frame id"); return CMLBA_ERR_ARG; }
        const uint16_t low = (uint16_t) ((1u << fi) - 1u);
        for (size_t i = 0; i < points_.size(); i++) {
            PointHost &p = points_[i];
            if (!p.alive) continue;
            if (p.host == fi) { drop_point(i, false); continue; }          // removePoints(frame's points) (DSOContext.h:155-156)
            if (p.host > fi) p.host--;
            p.res_mask = (uint16_t) ((p.res_mask & low) | ((p.res_mask >> (fi + 1)) << fi));   // squeeze slot fi out
            if (p.res_mask == 0) kill_point(i);       // points left without residuals disappear as well (DSOContext.h:205-216)
        }
        cudaSetDevice(device);
        cudaStreamSynchronize(stream);
        if (keep_prior() && !hm_frame_done) hm_drop(fi, 8 * (int) frames_.size() + 4);
        if (frames_[fi].d_img) img_pool.push_back(frames_[fi].d_img);
        frames_.erase(frames_.begin() + fi);
        maybe_compact();
        return CMLBA_OK;
    }

    // ------------------------------------------------------------------ marginalisation prior H_M, b_M (host, fp64)
    bool keep_prior() const { return !cfg.disable_marginalization; }
    void hm_fit(int n) { if ((int) HM.size() != n * n) { HM.assign((size_t) n * n, 0.0); bM.assign(n, 0.0); } }
    // addNewFrame: conservativeResize + zero tail (BA:438-443)
    void hm_grow() {
        const int n = 8 * (int) frames_.size() + 4, o = n - 8;
        std::vector<double> H((size_t) n * n, 0.0), b(n, 0.0);
        if ((int) HM.size() == o * o) { for (int r = 0; r < o; r++) { memcpy(&H[(size_t) r * n], &HM[(size_t) r * o], o * sizeof(double)); b[r] = bM[r]; } }
        HM.swap(H); bM.swap(b);
    }
    // plain removal of a frame's 8 rows / columns (cmlba_remove_frame)
    void hm_drop(int fi, int n) {
        if ((int) HM.size() != n * n) return;
        const int io = 4 + 8 * fi, m = n - 8;
        std::vector<double> H((size_t) m * m), b(m);
        auto src = [&](int i) { return i < io ? i : i + 8; };
        for (int r = 0; r < m; r++) { b[r] = bM[src(r)]; for (int c = 0; c < m; c++) H[(size_t) r * m + c] = HM[(size_t) src(r) * n + src(c)]; }
        HM.swap(H); bM.swap(b);
    }
    // marginalizeFrame (BA:464-548): move the frame's block to the end, add its prior, Jacobi-scale, Schur-complement the
    // 8x8 block out, unscale, symmetrise.
    void hm_marginalize_frame(int fi) {
        const int N = (int) frames_.size(), n = 8 * N + 4, m = n - 8, io = 4 + 8 * fi;
        hm_fit(n);
        std::vector<int> perm(n);
        for (int i = 0; i < n; i++) perm[i] = i < io ? i : (i < m ? i + 8 : io + (i - m));
        std::vector<double> H((size_t) n * n), b(n);
        for (int r = 0; r < n; r++) { b[r] = bM[perm[r]]; for (int c = 0; c < n; c++) H[(size_t) r * n + c] = HM[(size_t) perm[r] * n + perm[c]]; }
        const FrameHost &f = frames_[fi];
        for (int k = 0; k < 8; k++) { H[(size_t) (m + k) * n + m + k] += f.prior[k]; b[m + k] += f.prior[k] * f.state[k]; }   // delta_prior = state - prior_zero, prior_zero = 0
        std::vector<double> sv(n), si(n);
        for (int i = 0; i < n; i++) { sv[i] = sqrt(fabs(H[(size_t) i * n + i]) + 10.0); si[i] = 1.0 / sv[i]; }
        for (int r = 0; r < n; r++) { b[r] *= si[r]; for (int c = 0; c < n; c++) H[(size_t) r * n + c] *= si[r] * si[c]; }
        // hpi = inverse of the bottom-right 8x8 block (Gauss-Jordan, partial pivoting)
        double A[8][16];
        for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { A[r][c] = H[(size_t) (m + r) * n + m + c]; A[r][8 + c] = r == c ? 1.0 : 0.0; }
        for (int k = 0; k < 8; k++) {
            int piv = k; for (int r = k + 1; r < 8; r++) if (fabs(A[r][k]) > fabs(A[piv][k])) piv = r;
            if (piv != k) for (int c = 0; c < 16; c++) std::swap(A[k][c], A[piv][c]);
            const double d = 1.0 / A[k][k];
            for (int c = 0; c < 16; c++) A[k][c] *= d;
            for (int r = 0; r < 8; r++) if (r != k) { const double fct = A[r][k]; if (fct != 0.0) for (int c = 0; c < 16; c++) A[r][c] -= fct * A[k][c]; }
        }
        // bli = bottomLeft^T * hpi (m x 8); top-left -= bli * bottomLeft; b.head -= bli * b.tail
        std::vector<double> bli((size_t) m * 8);
        for (int r = 0; r < m; r++) for (int c = 0; c < 8; c++) { double s = 0; for (int k = 0; k < 8; k++) s += H[(size_t) (m + k) * n + r] * A[k][8 + c]; bli[(size_t) r * 8 + c] = s; }
        std::vector<double> Hn((size_t) m * m), bn(m);
        for (int r = 0; r < m; r++) {
            double s = b[r]; for (int k = 0; k < 8; k++) s -= bli[(size_t) r * 8 + k] * b[m + k];
            bn[r] = s * sv[r];
            for (int c = 0; c < m; c++) { double v = H[(size_t) r * n + c]; for (int k = 0; k < 8; k++) v -= bli[(size_t) r * 8 + k] * H[(size_t) (m + k) * n + c]; Hn[(size_t) r * m + c] = v * sv[r] * sv[c]; }
        }
        HM.assign((size_t) m * m, 0.0); bM = bn;
        for (int r = 0; r < m; r++) for (int c = 0; c < m; c++) HM[(size_t) r * m + c] = 0.5 * (Hn[(size_t) r * m + c] + Hn[(size_t) c * m + r]);
    }

    // ------------------------------------------------------------------ window maintenance decisions (host, exact)
    int frame_residual_count(int t) const { int c = 0; for (auto &p : points_) if (p.alive && ((p.res_mask >> t) & 1)) c++; return c; }

    // flagFramesForMarginalization (BA:603-708)
    int flag_frames(const double *cams, const int32_t *num_immature, int64_t *ids, int *n_ids) {
        const int N = (int) frames_.size();
        if (N == 0) { if (n_ids) *n_ids = 0; return CMLBA_OK; }
        auto cam = [&](int i) { Pose c; if (cams) { for (int k = 0; k < 9; k++) c.R[k] = cams[12 * i + k]; for (int k = 0; k < 3; k++) c.t[k] = cams[12 * i + 9 + k]; } else c = frames_[i].pre; return c; };
        int flagged = 0;
        const FrameHost &back = frames_[N - 1];
        for (int i = 0; i < N; i++) {
            FrameHost &f = frames_[i];
            const double in = (double) frame_residual_count(i) + (num_immature ? (double) num_immature[i] : 0.0);
            const double out = (double) f.num_marginalized + (double) f.num_residuals_out;
            // frameBack->getExposure().to(frame->getExposure()) (map/Exposure.h:119-123): a = exp(a_f - a_back) * tau_f / tau_back
            const double ref_to_fh = exp(f.aff_a - back.aff_a) * f.exposure / back.exposure;
            const bool not_enough = in < 0.05 * (in + out);
            const bool too_big = fabs(log(ref_to_fh)) > 0.7 && N - flagged > cfg.max_frames - 2;
            if (not_enough || too_big) { f.flagged = true; flagged++; }
        }
        if (N - flagged >= cfg.max_frames) {
            double smallest = 1; int to_marg = -1;
            const int latest_key = back.keyid;
            const Pose cb = cam(N - 1);
            for (int r = 0; r < N; r++) {
                const FrameHost &ref = frames_[r];
                if (ref.keyid > latest_key - cfg.frame_min_age || ref.keyid == 0) continue;
                const Pose cr_inv = pose_inv(cam(r));
                double score = 0;
                for (int t = 0; t < N; t++) {
                    if (t == r) continue;
                    if (frames_[t].keyid > latest_key - cfg.frame_min_age + 1) continue;
                    const Pose rel = pose_mul(cam(t), cr_inv);                 // reference->getCamera().to(target->getCamera())
                    score += 1.0 / (1e-5 + sqrt(rel.t[0] * rel.t[0] + rel.t[1] * rel.t[1] + rel.t[2] * rel.t[2]));
                }
                const Pose relb = pose_mul(cb, cr_inv);
                score *= -sqrt(sqrt(relb.t[0] * relb.t[0] + relb.t[1] * relb.t[1] + relb.t[2] * relb.t[2]));
                if (score < smallest) { smallest = score; to_marg = r; }
            }
            if (to_marg >= 0) { frames_[to_marg].flagged = true; flagged++; }
        }
        if (n_ids) {
            const int cap = *n_ids; int k = 0;
            for (auto &f : frames_) if (f.flagged) { if (ids && k < cap) ids[k] = f.id; k++; }
            *n_ids = k;
        }
        return CMLBA_OK;
    }

    // tryMarginalize (BA:2240-2363) with isOOB (BA:2515-2554); residual states are those of the last run()
    int try_marginalize(int *n_dropped, int *n_to_marg) {
        const int N = (int) frames_.size();
        const size_t PA = points_.size();
        std::vector<int> num_in(PA, 0), vis(PA, 0);
        std::vector<uint16_t> seen(PA, 0);
        if (snap.valid) {
            const int *rp = snap.own_map ? snap.r_point.data() : up.host<int>(up_o_rp);
            const uint8_t *rt = snap.own_map ? snap.r_target.data() : up.host<uint8_t>(up_o_rt);
            std::vector<int> pidx(snap.point_id.size()), fslot(snap.frame_id.size());
            for (size_t i = 0; i < pidx.size(); i++) pidx[i] = find_point(snap.point_id[i]);
            for (size_t i = 0; i < fslot.size(); i++) fslot[i] = frame_index(snap.frame_id[i]);
            for (int i = 0; i < snap.R; i++) {
                if (!snap.alive[i]) continue;
                const int q = pidx[rp[i]], t = fslot[rt[i]];
                if (q < 0 || t < 0 || !points_[q].alive || !((points_[q].res_mask >> t) & 1)) continue;
                seen[q] |= (uint16_t) (1u << t);
                if (snap.state[i] == RES_IN) { num_in[q]++; if (frames_[t].flagged) vis[q]++; }
            }
        }
        int dropped = 0, tomarg = 0, st_oob = 0, st_in = 0, st_inin = 0;     // flag_oob / flag_in / flag_inin of BA:2268 (flag_nores is never incremented there)
        std::vector<size_t> to_drop;
        for (size_t q = 0; q < PA; q++) {
            PointHost &p = points_[q];
            if (!p.alive) continue;
            // residuals created since the last run() are in state IN (DSOResidual.h:81-86)
            for (unsigned m = p.res_mask & ~seen[q]; m; m &= m - 1) { num_in[q]++; if (frames_[__builtin_ctz(m)].flagged) vis[q]++; }
            const int nres = __builtin_popcount(p.res_mask);
            if (p.idepth < 0 || nres == 0) { to_drop.push_back(q); st_oob++; continue; }
            bool oob;
            if (num_in[q] >= 3 && p.num_good > 4 + 10 && num_in[q] - vis[q] < 3) oob = true;
            else if (p.last_state[0] == CMLBA_RES_OOB) oob = true;
            else if (num_in[q] < 2) oob = false;
            else oob = p.last_state[0] == CMLBA_RES_OUTLIER && p.last_state[1] == CMLBA_RES_OUTLIER;
            if (oob || frames_[p.host].flagged) {
                if (nres >= 3 && p.num_good >= 4) st_in++;
                if (nres >= 3 && p.num_good >= 4 && p.idepth_hessian > cfg.min_idepth_h_marg) { p.to_marginalize = true; tomarg++; st_inin++; }
                else { to_drop.push_back(q); st_oob++; }
            }
        }
        for (size_t q : to_drop) { outliers_.push_back(points_[q].id); drop_point(q, false); dropped++; }
        (void) N;
        if (n_dropped) *n_dropped = dropped;
        if (n_to_marg) *n_to_marg = tomarg;
        stats_[14] = st_oob; stats_[15] = st_in; stats_[16] = st_inin; stats_[17] = 0.0;      // BA:2356-2359
        if (dropped) maybe_compact();
        return CMLBA_OK;
    }

    // marginalizePointsF (BA:2466-2513).  With a live prior (disableMarginalization = false) the marked points are first
    // accumulated in MARGINALIZED mode on a device window that holds only them, and H_M += 1/4 (M - M_sc), b_M += 1/4 (b - b_sc).
    int marginalize_points(int64_t *ids, int *n) {
        const int cap = n ? *n : 0; int k = 0;
        bool any = false;
        for (auto &p : points_) if (p.alive && p.to_marginalize) { any = true; break; }
        if (any && keep_prior() && world > 1) { set_error("the marginalisation prior is not supported together with point sharding"); return CMLBA_ERR_UNSUPPORTED; }
        if (any && keep_prior()) {
            subset_to_marginalize = true; dirty = true;
            int rc = prepare(nullptr);                                   // setZero, computeAdjoints, computeDelta on the subset window (BA:2474-2486)
            if (rc == CMLBA_OK) {
                dw.marg_mode = 1;
                launch_linearize(0, 0);                                  // tryMarginalize's resetOOB + linearize + applyRes + fixLinearization (BA:2289-2300)
                commit_candidate_kernel<<<1, 32, 0, stream>>>(dw); launches++;
                launch_tail(0);
                const int nn_ = dw.n * dw.n, n_ = dw.n;
                std::vector<double> sys((size_t) 2 * nn_ + 2 * n_);
                if (cudaMemcpyAsync(sys.data(), d_sys.p, sys.size() * 8, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { set_error("marginalize_points: device error"); rc = CMLBA_ERR_CUDA; }
                else {
                    hm_fit(n_);
                    const double *HA = sys.data(), *bA = HA + nn_, *HS = bA + n_, *bS = HS + nn_;
                    for (int i = 0; i < nn_; i++) HM[i] += 0.25 * (HA[i] - HS[i]);          // setting_margWeightFac = 0.5 * 0.5 (BA:2505-2508)
                    for (int i = 0; i < n_; i++) bM[i] += 0.25 * (bA[i] - bS[i]);
                }
                dw.marg_mode = 0;
            }
            subset_to_marginalize = false; dirty = true; prepared = false;
            if (rc) return rc;
        }
        for (size_t q = 0; q < points_.size(); q++) {
            PointHost &p = points_[q];
            if (!p.alive || !p.to_marginalize) continue;
            if (ids && k < cap) ids[k] = p.id;
            k++;
            p.to_marginalize = false;
            drop_point(q, true);
        }
        if (n) *n = k;
        if (k) maybe_compact();
        return CMLBA_OK;
    }

    // marginalizeFrames (BA:710-742)
    int marginalize_frames(int64_t *ids, int *n) {
        const int cap = n ? *n : 0; int k = 0;
        std::vector<int64_t> gone;
        for (auto &f : frames_) if (f.flagged) gone.push_back(f.id);
        for (int64_t id : gone) {
            if (ids && k < cap) ids[k] = id;
            k++;
            if (keep_prior()) { hm_marginalize_frame(frame_index(id)); hm_frame_done = true; }
            int rc = remove_frame(id);
            hm_frame_done = false;
            if (rc) return rc;
        }
        if (n) *n = k;
        return CMLBA_OK;
    }

    // ------------------------------------------------------------------ device window
    // Layout: points sorted by host frame, residuals sorted by bin = t*N+h then by device point
    // position.  Both orders come from stable counting sorts (O(P+R)); every uploaded array lives in one pinned
    // arena mirrored on the device and travels in a single copy.
    int build_device_window() {
        TSCOPE("build_device_window");
        Lap lap(timers);
        const int N = (int) frames_.size(), PA = (int) points_.size();
        const int n = 8 * N + 4;
        CK(cudaSetDevice(device));
        // alive points (or, for marginalize_points, only the marked ones): stable counting sort by host
        auto in_window = [&](const PointHost &p) { return p.alive && (!subset_to_marginalize || p.to_marginalize); };
        std::vector<int> hcnt(N + 1, 0);
        for (int i = 0; i < PA; i++) if (in_window(points_[i])) hcnt[points_[i].host + 1]++;
        for (int h = 0; h < N; h++) hcnt[h + 1] += hcnt[h];
        const int P = hcnt[N];
        pt_order.resize(P);
        { std::vector<int> o(hcnt.begin(), hcnt.end() - 1); for (int i = 0; i < PA; i++) if (in_window(points_[i])) pt_order[o[points_[i].host]++] = i; }
        // residuals per bin = t*N+h from the per-point masks (device order inside a bin = device point order: no sort)
        std::vector<int> bcnt(N * N + 1, 0);
        std::vector<uint16_t> dmask(P);
        for (int i = 0; i < P; i++) {
            const PointHost &p = points_[pt_order[i]];
            dmask[i] = p.res_mask;
            for (unsigned m = p.res_mask; m; m &= m - 1) bcnt[__builtin_ctz(m) * N + p.host + 1]++;
        }
        for (int b = 0; b < N * N; b++) bcnt[b + 1] += bcnt[b];
        const int R = bcnt[N * N];
        res_bin_begin = bcnt;
        lap("bdw.sort");
        if (P >= (1 << 24)) { set_error("more than 2^24 points in the window"); return CMLBA_ERR_ARG; }
        // warp passes of the sampling kernel (32 consecutive residuals of the tile-sorted order), Schur chunks per host
        n_chunks = (R + ACC_CHUNK - 1) / ACC_CHUNK;
        const int tiles_x = (W + LT_TILE_W - 1) / LT_TILE_W, tiles_y = (H + LT_TILE_H - 1) / LT_TILE_H, n_tiles = tiles_x * tiles_y;
        const size_t n_keys = (size_t) N * n_tiles * N;
        h_host_chunk_begin.assign(N + 1, 0);
        for (int h = 0; h < N; h++) h_host_chunk_begin[h + 1] = h_host_chunk_begin[h] + (hcnt[h + 1] - hcnt[h] + SC_CHUNK - 1) / SC_CHUNK;
        n_sc_chunks = h_host_chunk_begin[N];
        const int NB = 8 * N;
        const int sc_stride = ((NB * NB + NB * 4 + NB + 20) + 3) & ~3;
        const int n_pt_blocks = (P + 255) / 256;
        // arena layout
        up.begin();
        const size_t o_pt_host = up.take<int>(P), o_x = up.take<float>(P), o_y = up.take<float>(P), o_idz = up.take<float>(P), o_prior = up.take<float>(P),
                     o_mrb = up.take<float>(P), o_idh = up.take<float>(P), o_ng = up.take<int>(P), o_id = up.take<double>(P),
                     o_rp = up.take<int>(R), o_rh = up.take<uint8_t>(R), o_rt = up.take<uint8_t>(R),
                     o_rbb = up.take<int>(N * N + 1),
                     o_sh = up.take<int>(n_sc_chunks), o_sbeg = up.take<int>(n_sc_chunks), o_scnt = up.take<int>(n_sc_chunks), o_hcb = up.take<int>(N + 1);
        CK(cudaStreamSynchronize(stream));          // the previous upload from this pinned block must have landed
        materialize_snapshot();                     // the residual snapshot of the last run() still points into this block
        CK(up.commit());
        up_o_rp = o_rp; up_o_rt = o_rt;
        {
            int *h_pt_host = up.host<int>(o_pt_host), *h_ng = up.host<int>(o_ng);
            float *h_x = up.host<float>(o_x), *h_y = up.host<float>(o_y), *h_idz = up.host<float>(o_idz), *h_prior = up.host<float>(o_prior), *h_mrb = up.host<float>(o_mrb),
                  *h_idh = up.host<float>(o_idh);
            double *h_id = up.host<double>(o_id);
            for (int i = 0; i < P; i++) {
                const PointHost &p = points_[pt_order[i]];
                h_pt_host[i] = p.host; h_x[i] = p.x; h_y[i] = p.y; h_id[i] = p.idepth; h_idz[i] = p.idepth_zero;
                h_prior[i] = p.has_prior ? (float) cfg.idepth_fix_prior : 0.f;
                h_ng[i] = p.num_good; h_mrb[i] = p.max_rel_bs; h_idh[i] = p.idepth_hessian;
            }
            int *h_rp = up.host<int>(o_rp); uint8_t *h_rh = up.host<uint8_t>(o_rh), *h_rt = up.host<uint8_t>(o_rt);
            for (int t = 0; t < N; t++) for (int h = 0; h < N; h++) {
                int k = bcnt[t * N + h];
                const int ke = bcnt[t * N + h + 1];
                if (k == ke) continue;
                memset(h_rh + k, h, ke - k); memset(h_rt + k, t, ke - k);
                for (int i = hcnt[h]; i < hcnt[h + 1]; i++) if ((dmask[i] >> t) & 1) h_rp[k++] = i;
            }
            int *sh = up.host<int>(o_sh), *sbeg = up.host<int>(o_sbeg), *scnt = up.host<int>(o_scnt);
            for (int h = 0, c = 0; h < N; h++)
                for (int s0 = hcnt[h]; s0 < hcnt[h + 1]; s0 += SC_CHUNK, c++) { sh[c] = h; sbeg[c] = s0; scnt[c] = std::min(SC_CHUNK, hcnt[h + 1] - s0); }
            memcpy(up.host<int>(o_rbb), bcnt.data(), (N * N + 1) * sizeof(int));
            memcpy(up.host<int>(o_hcb), h_host_chunk_begin.data(), (N + 1) * sizeof(int));
        }
        const int newest_begin = N > 0 ? bcnt[(N - 1) * N] : R;   // bins are t-major: residuals targeting the newest frame are the tail
        lap("bdw.pack");
        // device-only buffers
        const size_t Rz = std::max(R, 1), Pz = std::max(P, 1);
        CK(d_pairs.reserve(MAXF * MAXF));
        CK(d_pt_idb.reserve(Pz)); CK(d_pt_colors.reserve(Pz * 8)); CK(d_pt_weights.reserve(Pz * 8));
        CK(d_pt_Hdd.reserve(Pz)); CK(d_pt_bd.reserve(Pz)); CK(d_pt_Hcd.reserve(Pz * 4)); CK(d_pt_HdiF.reserve(Pz)); CK(d_pt_bdSumF.reserve(Pz));
        CK(d_pt_ngood_cur.reserve(Pz)); CK(d_pt_step.reserve(Pz));
        CK(d_r_state0.reserve(Rz)); CK(d_r_state1.reserve(Rz)); CK(d_r_energy0.reserve(Rz)); CK(d_r_energy1.reserve(Rz)); CK(d_r_good0.reserve(Rz)); CK(d_r_good1.reserve(Rz));
        CK(d_r_new_state.reserve(Rz)); CK(d_r_new_energy.reserve(Rz)); CK(d_r_new_energy_wo.reserve(Rz)); CK(d_r_alive.reserve(Rz)); CK(d_r_center.reserve(Rz * 3));
        CK(d_rj0.reserve(Rz * RJ_STRIDE)); CK(d_rj1.reserve(Rz * RJ_STRIDE)); CK(d_T0.reserve(Pz * N * T_STRIDE)); CK(d_T1.reserve(Pz * N * T_STRIDE));
        if (want_dbg) CK(d_dbg.reserve(Rz * DBG_STRIDE));
        CK(d_energy_part.reserve(std::max(n_chunks, 1)));
        CK(d_acc_bin.reserve((size_t) N * N * ACC_SLICES * ACC_N));
        CK(d_bin_key.reserve(Rz)); CK(d_bin_hist.reserve(n_keys)); CK(d_bin_offs.reserve(n_keys)); CK(d_job_of_tile.reserve((size_t) N * n_tiles)); CK(d_job_desc.reserve((size_t) N * n_tiles)); CK(d_job_begin.reserve((size_t) N * n_tiles + 1)); CK(d_cta_info.reserve((size_t) LT_INFO_INTS * 1024));
        CK(d_r_pht.reserve(Rz)); CK(d_r_job.reserve(Rz)); CK(d_r_src.reserve(Rz)); CK(d_r_pt4.reserve(Rz * 5)); CK(d_bin_g.reserve(Rz));
        CK(d_fin_state.reserve(Rz)); CK(d_fin_alive.reserve(Rz)); CK(d_fin_energy.reserve(Rz));
        if (!d_bin_ticket.p) { CK(d_bin_ticket.reserve(2)); CK(cudaMemsetAsync(d_bin_ticket.p, 0, 2 * sizeof(int), stream)); }
        CK(d_sc_part.reserve((size_t) std::max(n_sc_chunks, 1) * sc_stride));
        CK(d_st_out.reserve((size_t) N * N * st_stride(N)));
        CK(d_sys.reserve((size_t) 2 * n * n + 2 * n)); CK(d_x.reserve(n)); CK(d_xAd.reserve((size_t) N * N * 8)); CK(d_pt_part.reserve((size_t) std::max(n_pt_blocks, 1) * 4));
        lap("bdw.alloc");
        if (up.used) CK(cudaMemcpyAsync(up.d.p, up.h.p, up.used, cudaMemcpyHostToDevice, stream));
        // views into the arena
#define VIEW(buf, T, off) do { (buf).p = up.dev<T>(off); (buf).view = true; (buf).cap = 0; } while (0)
        VIEW(d_pt_host, int, o_pt_host); VIEW(d_pt_x, float, o_x); VIEW(d_pt_y, float, o_y); VIEW(d_pt_idz, float, o_idz); VIEW(d_pt_priorF, float, o_prior);
        VIEW(d_pt_mrb, float, o_mrb); VIEW(d_pt_idh, float, o_idh); VIEW(d_pt_num_good, int, o_ng); VIEW(d_pt_idepth, double, o_id);
        VIEW(d_r_point, int, o_rp); VIEW(d_r_host, uint8_t, o_rh); VIEW(d_r_target, uint8_t, o_rt);
        VIEW(d_res_bin_begin, int, o_rbb);
        VIEW(d_sc_chunk_host, int, o_sh); VIEW(d_sc_chunk_begin, int, o_sbeg); VIEW(d_sc_chunk_count, int, o_scnt); VIEW(d_host_chunk_begin, int, o_hcb);
#undef VIEW
        lap("bdw.upload");
        // DevWin
        DevWin &w = dw;
        memset(&w, 0, sizeof(w));
        w.N = N; w.P = P; w.R = R; w.W = W; w.H = H; w.n = n; w.newest_begin = newest_begin;
        w.n_chunks = n_chunks; w.n_sc_chunks = n_sc_chunks;
        w.tiles_x = tiles_x; w.tiles_y = tiles_y; w.n_tiles = n_tiles;
        w.fx = fx; w.fy = fy; w.cx = cx; w.cy = cy; w.fxi = 1.0 / fx; w.fyi = 1.0 / fy;
        w.huber = cfg.huber_threshold; w.cth = cfg.outlier_th_sum; w.scaleF = cfg.scale_f; w.scaleC = cfg.scale_c;
        w.scaleA = cfg.scale_light_a; w.scaleB = cfg.scale_light_b; w.scaleT = cfg.scale_translation; w.scaleR = cfg.scale_rotation;
        w.th_opt = cfg.th_opt_iterations; w.optA = cfg.optimize_light_a; w.optB = cfg.optimize_light_b;
        w.force_accept = cfg.force_accept; w.fix_lambda = cfg.fix_lambda; w.idepth_fix_prior = cfg.idepth_fix_prior; w.fixed_lambda = (double) cfg.fixed_lambda;
        for (int i = 0; i < N; i++) w.img[i] = frames_[i].d_img;
        w.frames = d_frames.p; w.pairs = d_pairs.p; w.ctrl = d_ctrl.p; w.AH = d_AH.p; w.AT = d_AT.p; w.HM = d_HM.p; w.bM = d_bM.p; w.Pns = d_Pns.p;
        w.pt_host = d_pt_host.p; w.pt_x = d_pt_x.p; w.pt_y = d_pt_y.p; w.pt_idepth = d_pt_idepth.p; w.pt_idepth_zero = d_pt_idz.p; w.pt_idepth_backup = d_pt_idb.p;
        w.pt_colors = d_pt_colors.p; w.pt_weights = d_pt_weights.p; w.pt_priorF = d_pt_priorF.p;
        w.pt_Hdd = d_pt_Hdd.p; w.pt_bd = d_pt_bd.p; w.pt_Hcd = d_pt_Hcd.p; w.pt_HdiF = d_pt_HdiF.p; w.pt_bdSumF = d_pt_bdSumF.p; w.pt_idepth_hessian = d_pt_idh.p; w.pt_max_rel_bs = d_pt_mrb.p;
        w.pt_num_good = d_pt_num_good.p; w.pt_ngood_cur = d_pt_ngood_cur.p; w.pt_step = d_pt_step.p;
        w.r_point = d_r_point.p; w.r_host = d_r_host.p; w.r_target = d_r_target.p; w.res_bin_begin = d_res_bin_begin.p;
        w.bin_key = d_bin_key.p; w.bin_hist = d_bin_hist.p; w.bin_offs = d_bin_offs.p; w.job_of_tile = d_job_of_tile.p; w.job_desc = d_job_desc.p; w.job_begin = d_job_begin.p; w.cta_info = d_cta_info.p; w.r_pt4 = d_r_pt4.p; w.bin_g = d_bin_g.p; w.lt_grid = std::max(1, std::min(std::min(n_sm, 1024), n_chunks));
        w.r_pht = d_r_pht.p; w.r_job = d_r_job.p; w.r_src = d_r_src.p;
        w.bin_ticket = d_bin_ticket.p; w.fin_state = d_fin_state.p; w.fin_alive = d_fin_alive.p; w.fin_energy = d_fin_energy.p;
        w.tma_on = encode_tile_maps() ? 1 : 0;
        // sparse windows: a staged 42 KB box only pays when enough residuals sample it.  Below ~24 residuals per (frame, tile) on average the taps
        // are read from the image directly (measured at the C5 shape, 1920x1080 with 14 residuals per tile: 115 us staged, 91 us direct; C2 has 93)
        if (!getenv("CMLBA_FORCE_TMA") && (double) R < 24.0 * (double) N * (double) n_tiles) w.tma_on = 0;
        w.r_state[0] = d_r_state0.p; w.r_state[1] = d_r_state1.p; w.r_energy[0] = d_r_energy0.p; w.r_energy[1] = d_r_energy1.p; w.r_good[0] = d_r_good0.p; w.r_good[1] = d_r_good1.p;
        w.r_new_state = d_r_new_state.p; w.r_new_energy = d_r_new_energy.p; w.r_new_energy_wo = d_r_new_energy_wo.p; w.r_alive = d_r_alive.p; w.r_center = d_r_center.p;
        w.rj[0] = d_rj0.p; w.rj[1] = d_rj1.p; w.T[0] = d_T0.p; w.T[1] = d_T1.p; w.dbg = want_dbg ? d_dbg.p : nullptr;
        w.energy_part = d_energy_part.p; w.acc_bin = d_acc_bin.p;
        w.sc_part = d_sc_part.p; w.sc_stride = sc_stride; w.sc_chunk_host = d_sc_chunk_host.p; w.sc_chunk_begin = d_sc_chunk_begin.p; w.sc_chunk_count = d_sc_chunk_count.p;
        w.host_chunk_begin = d_host_chunk_begin.p;
        w.st_out = d_st_out.p; w.sys = d_sys.p; w.x = d_x.p; w.xAd = d_xAd.p;
        w.pt_part = d_pt_part.p; w.n_pt_blocks = n_pt_blocks;
        // addPoint's reference colours (integer-pixel read) and gradient weights (DSOContext.h:86-91, BA:405-411) from the host images
        if (P > 0) { point_init_kernel<<<(unsigned) (((size_t) P * 8 + 255) / 256), 256, 0, stream>>>(w, 0, P, d_pt_colors.p, d_pt_weights.p); launches++; CK(cudaGetLastError()); }
        w.world = world; w.rank = rank; w.cand_cap = 0;
        w.p2p_on = (world > 1 && p2p_ready) ? 1 : 0;
        for (int q = 0; q < MAXF; q++) w.p2p_base[q] = p2p_peer[q];
        if (world > 1) {
            // record capacity = the largest per-rank count of residuals towards the newest frame (one small all-reduce per window build)
            int mine = R - newest_begin, cap = 0;
            CK(d_cap.reserve(1));
            CK(cudaMemcpyAsync(d_cap.p, &mine, sizeof(int), cudaMemcpyHostToDevice, stream));
            if (g_nccl.AllReduce(d_cap.p, d_cap.p, 1, /*ncclInt32*/ 2, /*ncclMax*/ 2, comm, stream) != 0) { set_error("ncclAllReduce(max) failed"); return CMLBA_ERR_CUDA; }
            CK(cudaMemcpyAsync(&cap, d_cap.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            w.cand_cap = std::max(cap, 1);
            const size_t rec_d = 8 + (size_t) (w.cand_cap + 1) / 2;
            CK(d_post_send.reserve(rec_d)); CK(d_post_recv.reserve(rec_d * world));
            w.post_send = d_post_send.p; w.post_recv = d_post_recv.p; w.post_stride = (int) rec_d;
            w.p2p_post_on = (w.p2p_on && w.cand_cap <= P2P_POST_CAND_MAX) ? 1 : 0;
            if (w.p2p_post_on) w.post_stride = (int) P2P_POST_DOUBLES;
        }
        dirty = false;
        return CMLBA_OK;
    }

    // One CUtensorMap per window frame for the TMA box loads of linearize_tile_kernel: the image (H rows of W float4 texels) seen as a
    // 2-D tensor of 8-byte elements (2W x H; a 16-byte element type does not exist), box = (2 * LT_BOX_W) x LT_BOX_H = 46 080 B,
    // out-of-bounds elements zero-filled.  cuTensorMapEncodeTiled comes through cudaGetDriverEntryPoint (libcuda is not linked).
    // false (-> every tap is read from global memory) if the driver entry point is missing, an encode fails, or CMLBA_NO_TMA is set.
    bool encode_tile_maps() {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        static bool tried = false;
        if (!tried) {
            tried = true;
            void *fp = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) encode = (EncodeFn) fp;
            cudaGetLastError();
        }
        if (!encode || getenv("CMLBA_NO_TMA")) return false;
        for (size_t i = 0; i < frames_.size(); i++) {
            const cuuint64_t gdim[2] = {(cuuint64_t) 2 * W, (cuuint64_t) H};
            const cuuint64_t gstr[1] = {(cuuint64_t) W * 16};
            const cuuint32_t box[2] = {2 * LT_BOX_W, LT_BOX_H};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult rc = encode(&tile_maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, frames_[i].d_img, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) return false;
        }
        return true;
    }

    // 7x7 symmetric eigen-decomposition (cyclic Jacobi), eigenvectors in columns of V
    static void jacobi_eig(int k, double *A, double *V) {
        for (int i = 0; i < k * k; i++) V[i] = (i % (k + 1) == 0) ? 1.0 : 0.0;
        for (int sweep = 0; sweep < 60; sweep++) {
            double off = 0;
            for (int p = 0; p < k; p++) for (int q = p + 1; q < k; q++) off += A[p * k + q] * A[p * k + q];
            if (off < 1e-30) break;
            for (int p = 0; p < k; p++) for (int q = p + 1; q < k; q++) {
                if (fabs(A[p * k + q]) < 1e-300) continue;
                const double th = (A[q * k + q] - A[p * k + p]) / (2 * A[p * k + q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int i = 0; i < k; i++) { const double a = A[i * k + p], b = A[i * k + q]; A[i * k + p] = c * a - s * b; A[i * k + q] = s * a + c * b; }
                for (int i = 0; i < k; i++) { const double a = A[p * k + i], b = A[q * k + i]; A[p * k + i] = c * a - s * b; A[q * k + i] = s * a + c * b; }
                for (int i = 0; i < k; i++) { const double a = V[i * k + p], b = V[i * k + q]; V[i * k + p] = c * a - s * b; V[i * k + q] = s * a + c * b; }
            }
        }
    }

    // run() prologue: updateCamera, computeAdjoints, computeDelta (priors), nullspace projector; reset residuals
    int prepare(const double *cams) {
        TSCOPE("prepare(total)");
        Lap lap(timers);
        if (!have_calib) { set_error("calibration not set"); return CMLBA_ERR_STATE; }
        const int N = (int) frames_.size();
        if (N < 1) { set_error("no frames"); return CMLBA_ERR_STATE; }
        if (points_.size() == n_dead) { set_error("No points..."); return CMLBA_ERR_STATE; }   // BA:759-762
        CK(cudaSetDevice(device));
        join_side();                                  // nothing of the previous run may still count into the control block that is rebuilt below
        acc_jobs_launched = 0;
        if (dirty) { int rc = build_device_window(); if (rc) return rc; }
        const int n = 8 * N + 4;
        double sc[10]; scales(sc);
        lap("prep.build");
        // updateCamera -> setStateFromCamera (DSOFrame.h:143-151)
        for (int i = 0; i < N; i++) {
            FrameHost &f = frames_[i];
            if (cams) {
                Pose c;
                for (int k = 0; k < 9; k++) c.R[k] = cams[12 * i + k];
                for (int k = 0; k < 3; k++) c.t[k] = cams[12 * i + 9 + k];
                double xi[6];
                se3_log(pose_mul(c, pose_inv(f.evalpt)), xi);
                for (int k = 0; k < 6; k++) f.state[k] = xi[k] / sc[k];
            }
        }
        // computeAdjoints (BA:1062-1097); stored [h*N+t]
        std::vector<double> AH((size_t) N * N * 64, 0.0), AT((size_t) N * N * 64, 0.0);
        for (int h = 0; h < N; h++) for (int t = 0; t < N; t++) {
            Pose T0 = pose_mul(frames_[t].evalpt, pose_inv(frames_[h].evalpt));
            double adj[36]; se3_adj(T0, adj);
            const double a_h = frames_[h].state_zero[6] * sc[6], a_t = frames_[t].state_zero[6] * sc[6];
            const double a0 = exp(a_t - a_h) * frames_[t].exposure / frames_[h].exposure;
            double *ah = &AH[(size_t) (h * N + t) * 64], *at = &AT[(size_t) (h * N + t) * 64];
            for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) ah[r * 8 + c] = -adj[c * 6 + r];
            for (int r = 0; r < 6; r++) at[r * 8 + r] = 1.0;
            at[6 * 8 + 6] = -a0; ah[6 * 8 + 6] = a0; at[7 * 8 + 7] = -1.0; ah[7 * 8 + 7] = a0;
            for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { ah[r * 8 + c] *= sc[r]; at[r * 8 + c] *= sc[r]; }
        }
        // priors (BA:1127-1170); float settings
        const float pa = cfg.optimize_light_a ? 1e12f : 1e14f, pb = cfg.optimize_light_b ? 1e8f : 1e14f;
        for (auto &f : frames_) {
            for (int k = 0; k < 8; k++) f.prior[k] = 0;
            if (f.keyid == 0) { for (int k = 0; k < 3; k++) { f.prior[k] = 1e10f; f.prior[3 + k] = 1e11f; } f.prior[6] = 1e14f; f.prior[7] = 1e14f; }
            else { f.prior[6] = pa; f.prior[7] = pb; }
        }
        // nullspaces (DSOFrame.h:160-179) stacked as computeNullspaces does (BA:2372-2414) and the projector of orthogonalize (BA:1217-1249)
        std::vector<double> Nm((size_t) n * 7, 0.0);
        for (int i = 0; i < N; i++) {
            const Pose &E = frames_[i].evalpt; const Pose Ei = pose_inv(E);
            const int o = 4 + 8 * i;
            for (int k = 0; k < 6; k++) {
                double eps[6] = {0, 0, 0, 0, 0, 0}, lp[6], lm[6];
                eps[k] = 1e-3; se3_log(pose_mul(pose_mul(E, se3_exp(eps)), Ei), lp);
                eps[k] = -1e-3; se3_log(pose_mul(pose_mul(E, se3_exp(eps)), Ei), lm);
                for (int r = 0; r < 6; r++) Nm[(size_t) (o + r) * 7 + k] = (lp[r] - lm[r]) / 2e-3 / sc[r];
            }
            Pose Pp = E, Pm = E; double lp[6], lm[6];
            for (int k = 0; k < 3; k++) { Pp.t[k] *= 1.00001; Pm.t[k] /= 1.00001; }
            se3_log(pose_mul(Pp, Ei), lp); se3_log(pose_mul(Pm, Ei), lm);
            for (int r = 0; r < 6; r++) Nm[(size_t) (o + r) * 7 + 6] = (lp[r] - lm[r]) / 2e-3 / sc[r];
        }
        for (int k = 0; k < 7; k++) {   // normalised columns
            double s = 0; for (int r = 0; r < n; r++) s += Nm[(size_t) r * 7 + k] * Nm[(size_t) r * 7 + k];
            s = sqrt(s); if (s > 0) for (int r = 0; r < n; r++) Nm[(size_t) r * 7 + k] /= s;
        }
        // thin SVD through eig(N^T N): N = U S V^T ; N N+^T = N V S^-2 V^T N^T with the cut S_i > delta * S_max
        double G[49], V[49];
        for (int a = 0; a < 7; a++) for (int b = 0; b < 7; b++) { double s = 0; for (int r = 0; r < n; r++) s += Nm[(size_t) r * 7 + a] * Nm[(size_t) r * 7 + b]; G[a * 7 + b] = s; }
        jacobi_eig(7, G, V);
        double smax = 0; for (int k = 0; k < 7; k++) smax = std::max(smax, sqrt(std::max(G[k * 8], 0.0)));
        std::vector<double> NV((size_t) n * 7, 0.0), Pns((size_t) n * n, 0.0);
        for (int r = 0; r < n; r++) for (int k = 0; k < 7; k++) { double s = 0; for (int a = 0; a < 7; a++) s += Nm[(size_t) r * 7 + a] * V[a * 7 + k]; NV[(size_t) r * 7 + k] = s; }
        for (int k = 0; k < 7; k++) {
            const double sv = sqrt(std::max(G[k * 8], 0.0));
            if (!(sv > (double) cfg.solver_mode_delta * smax)) continue;
            const double inv = 1.0 / (sv * sv);
            for (int r = 0; r < n; r++) { const double a = NV[(size_t) r * 7 + k] * inv; if (a == 0) continue; for (int c = 0; c < n; c++) Pns[(size_t) r * n + c] += a * NV[(size_t) c * 7 + k]; }
        }
        // frames -> device
        std::vector<FrameDev> fd(N);
        for (int i = 0; i < N; i++) {
            FrameHost &f = frames_[i]; FrameDev &d = fd[i];
            memset(&d, 0, sizeof(d));
            for (int k = 0; k < 9; k++) d.evalR[k] = f.evalpt.R[k];
            for (int k = 0; k < 3; k++) d.evalt[k] = f.evalpt.t[k];
            double ss[10];
            for (int k = 0; k < 10; k++) { d.state[k] = f.state[k]; d.state_zero[k] = f.state_zero[k]; d.state_backup[k] = f.state[k]; ss[k] = sc[k] * f.state[k]; d.state_scaled[k] = ss[k]; }
            Pose P = pose_mul(se3_exp(ss), f.evalpt);
            f.pre = P;
            for (int k = 0; k < 9; k++) d.preR[k] = P.R[k];
            for (int k = 0; k < 3; k++) d.pret[k] = P.t[k];
            for (int k = 0; k < 8; k++) d.prior[k] = f.prior[k];
            d.exposure = f.exposure; d.energy_th = f.energy_th; d.keyid = f.keyid;
        }
        hm_fit(n);
        if (cfg.disable_marginalization) { std::fill(HM.begin(), HM.end(), 0.0); std::fill(bM.begin(), bM.end(), 0.0); }   // BA:1395-1398
        dw.has_HM = 0;
        for (double v : HM) if (v != 0.0) { dw.has_HM = 1; break; }
        for (double v : bM) if (v != 0.0) { dw.has_HM = 1; break; }
        // adHTdeltaF (computeDelta, BA:1105-1117): delta_h^T AH + delta_t^T AT per pair, stored as float like the reference's casts
        std::vector<float> pdel((size_t) N * N * 8, 0.f);
        for (int h = 0; h < N; h++) for (int t = 0; t < N; t++) {
            const double *ah = &AH[(size_t) (h * N + t) * 64], *at = &AT[(size_t) (h * N + t) * 64];
            for (int c2 = 0; c2 < 8; c2++) {
                double v = 0;
                for (int r2 = 0; r2 < 8; r2++) v += (frames_[h].state[r2] - frames_[h].state_zero[r2]) * ah[r2 * 8 + c2] + (frames_[t].state[r2] - frames_[t].state_zero[r2]) * at[r2 * 8 + c2];
                pdel[(size_t) (h * N + t) * 8 + c2] = (float) v;
            }
        }
        Ctrl c; memset(&c, 0, sizeof(c));
        c.lambda = (double) cfg.fixed_lambda;
        lap("prep.host_math");
        // one pinned block mirrored on the device: frames | AH | AT | Pns | ctrl | (HM | bM when there is a prior)
        prep.begin();
        const size_t o_fr = prep.take<FrameDev>(MAXF), o_ah = prep.take<double>(AH.size()), o_at = prep.take<double>(AT.size()), o_pns = prep.take<double>(Pns.size()),
                     o_ctrl = prep.take<Ctrl>(1), o_hm = prep.take<double>(dw.has_HM ? HM.size() : 1), o_bm = prep.take<double>(dw.has_HM ? bM.size() : 1),
                     o_pdel = prep.take<float>(pdel.size());
        CK(cudaStreamSynchronize(stream));          // the previous upload from this pinned block must have landed
        CK(prep.commit());
        memcpy(prep.host<FrameDev>(o_fr), fd.data(), N * sizeof(FrameDev));
        memcpy(prep.host<double>(o_ah), AH.data(), AH.size() * 8); memcpy(prep.host<double>(o_at), AT.data(), AT.size() * 8);
        memcpy(prep.host<double>(o_pns), Pns.data(), Pns.size() * 8); memcpy(prep.host<Ctrl>(o_ctrl), &c, sizeof(c));
        if (dw.has_HM) { memcpy(prep.host<double>(o_hm), HM.data(), HM.size() * 8); memcpy(prep.host<double>(o_bm), bM.data(), bM.size() * 8); }
        memcpy(prep.host<float>(o_pdel), pdel.data(), pdel.size() * 4);
        CK(cudaMemcpyAsync(prep.d.p, prep.h.p, prep.used, cudaMemcpyHostToDevice, stream));
#define PVIEW(buf, T, off) do { (buf).release(); (buf).p = prep.dev<T>(off); (buf).view = true; } while (0)
        PVIEW(d_frames, FrameDev, o_fr); PVIEW(d_AH, double, o_ah); PVIEW(d_AT, double, o_at); PVIEW(d_Pns, double, o_pns); PVIEW(d_ctrl, Ctrl, o_ctrl);
        PVIEW(d_HM, double, o_hm); PVIEW(d_bM, double, o_bm);
#undef PVIEW
        dw.frames = d_frames.p; dw.AH = d_AH.p; dw.AT = d_AT.p; dw.Pns = d_Pns.p; dw.ctrl = d_ctrl.p; dw.HM = d_HM.p; dw.bM = d_bM.p;
        dw.pair_delta = prep.dev<float>(o_pdel); dw.marg_mode = 0;
        // resetOOB on every active residual (BA:766-779, DSOResidual.h:81-86), empty Schur tables
        reset_window_kernel<<<148 * 4, 256, 0, stream>>>(dw); launches++;
        pairs_kernel<<<(N * N + 63) / 64, 64, 0, stream>>>(dw, 0); launches++;
        if (dw.R > 0) {   // tile binning at the poses this run() starts from: residuals sorted by (target, tile, host)
            CK(cudaMemsetAsync(d_bin_hist.p, 0, (size_t) N * dw.n_tiles * N * sizeof(int), stream));
            bin_count_kernel<<<(dw.R + 255) / 256, 256, 0, stream>>>(dw); launches++;
            bin_scatter_kernel<<<N * N, 256, (size_t) 8 * dw.n_tiles * sizeof(int), stream>>>(dw); launches++;
            if (!getenv("CMLBA_NO_INTERLEAVE")) { bin_interleave_kernel<<<N * dw.n_tiles, 256, 0, stream>>>(dw); launches++; }
            bin_pack_kernel<<<(dw.R + 255) / 256, 256, 0, stream>>>(dw); launches++;
            bin_finish_kernel<<<1, 256, 0, stream>>>(dw); launches++;
        }
        CK(cudaGetLastError());
        lap("prep.upload_reset");
        prepared = true;
        return CMLBA_OK;
    }

    // ------------------------------------------------------------------ kernel sequences
    size_t stitch_smem() const { const int N = dw.N, NB = 8 * N; return sizeof(double) * ((size_t) 8 * NB + N * 64 + NB + 40 + ACC_N + 128 + N * 64); }
    size_t solve_smem() const { return sizeof(double) * solve_smem_doubles(dw.n); }
    size_t schur_smem() const { return schur_smem_bytes(dw.N); }

    // Kernels of the pass / Gauss-Newton chain are launched with programmatic stream serialization: kernel k+1 may be scheduled while
    // kernel k still runs; its CTAs block in pdl_enter() (griddepcontrol.wait) until kernel k has completed.  CMLBA_NO_PDL=1 turns it off.
    void kt_patch(DevWin &w) { w.ktrace = d_ktrace.p + 3 * kt_seq; }
    template <typename T> void kt_patch(T &) {}
    template <typename... KArgs, typename... Args>
    void launch_pdl(int site, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
        if (want_ktrace && d_ktrace.p && dw.ktrace && kt_seq < 127) { kt_seq++; kt_sites.push_back(site); (kt_patch(args), ...); }
        cudaLaunchConfig_t lc = {};
        lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = launch_stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = (use_pdl && (pdl_mask & site)) ? 1 : 0;
        lc.attrs = at; lc.numAttrs = 1;
        if (cudaLaunchKernelEx(&lc, kernel, KArgs(args)...) != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(cudaGetLastError())); launch_rc = CMLBA_ERR_CUDA; }
        launches++;
    }

    void launch_linearize(int fix, int respect_done) {
        if (dw.R == 0) return;
        const int grid = dw.lt_grid;                           // persistent CTAs: one per SM, a contiguous range of warp passes each
        dw.lt_mode = lt_mode; dw.lt_exact = lt_exact;
        if (lt_mode & 2) { if (d_lt_trace.reserve((size_t) 1024 * 16 * 32) == cudaSuccess) { cudaMemsetAsync(d_lt_trace.p, 0, (size_t) grid * 16 * 32 * 8, stream); dw.lt_trace = d_lt_trace.p; } }
#define LT_LAUNCH(CW, ST)                                                                                                                                      \
    do {                                                                                                                                                       \
        if (want_dbg) launch_pdl(1, linearize_tile_kernel<true, CW, ST>, dim3(grid), dim3(CW * 32), lt_smem_bytes(dw.N, CW, ST), dw, tile_maps, fix, respect_done);    \
        else launch_pdl(1, linearize_tile_kernel<false, CW, ST>, dim3(grid), dim3(CW * 32), lt_smem_bytes(dw.N, CW, ST), dw, tile_maps, fix, respect_done);            \
    } while (0)
        switch (lt_variant) {
            case 1: LT_LAUNCH(8, 4); break;
            case 2: LT_LAUNCH(16, 3); break;
            case 3: LT_LAUNCH(12, 3); break;
            default: if (lt_smem_bytes(dw.N, 12, 4) <= (size_t) 227 * 1024) LT_LAUNCH(12, 4); else LT_LAUNCH(12, 3); break;
        }
#undef LT_LAUNCH
    }
    void launch_post(int mode, int respect_done) {
        if (world > 1) {   // energy, convergence sums and the 0.7-quantile inputs of ALL ranks: one all-gather per linearization
            if (dw.p2p_post_on) {   // pushed through peer memory by pack_post_kernel itself
                dw.p2p_post_epoch = ++p2p_post_epoch;
                dw.post_recv = reinterpret_cast<const double *>(p2p_local + 256) + (size_t) 2 * world * P2P_SLOT_DOUBLES + (size_t) (dw.p2p_post_epoch & 1ull) * world * P2P_POST_DOUBLES;
                launch_pdl(2, pack_post_kernel, dim3(1), dim3(1024), 0, dw, respect_done);
            } else {
                launch_pdl(2, pack_post_kernel, dim3(1), dim3(1024), 0, dw, respect_done);
                const size_t rec_d = 8 + (size_t) (dw.cand_cap + 1) / 2;
                if (g_nccl.AllGather(d_post_send.p, d_post_recv.p, rec_d, /*ncclDouble*/ 8, comm, stream) != 0) { set_error("ncclAllGather failed"); launch_rc = CMLBA_ERR_CUDA; }
            }
        }
        launch_pdl(2, post_linearize_kernel, dim3(1), dim3(1024), 0, dw, mode, respect_done);
    }
    void launch_accumulate(int respect_done) {
        if (dw.R > 0) { launch_pdl(4, accumulate_kernel, dim3(dw.N * dw.N * ACC_SLICES), dim3(32), 0, dw, respect_done); acc_jobs_launched += dw.N * dw.N * ACC_SLICES; }
    }
    void launch_schur_only(int respect_done) {
        if (dw.n_sc_chunks > 0) launch_pdl(8, schur_kernel, dim3(dw.n_sc_chunks), dim3(256), schur_smem(), dw, respect_done);
    }
    // addToHessianTop from the Jacobian records ((bin, slice) jobs) on the side stream, the Schur chunks on the main one.  The side stream is
    // forked off with an event; it is NOT joined back per pass: stitch_pair_kernel waits on the device for Ctrl.acc_done_count to reach
    // dw.acc_target (a stream-level join costs ~3 us of idle time before the stitch), and the host joins once before it reads results
    // (join_side()).  A job that is skipped because the run is done counts as well, so the target stays in step.
    void launch_schur(int respect_done) {
        const bool fork = use_fork && dw.R > 0 && dw.n_sc_chunks > 0;
        if (fork) {
            if (cudaEventRecord(ev_fork, stream) != cudaSuccess || cudaStreamWaitEvent(side_stream, ev_fork, 0) != cudaSuccess) { set_error("fork failed"); launch_rc = CMLBA_ERR_CUDA; }
            launch_stream = side_stream; launch_accumulate(respect_done); launch_stream = stream;
            side_pending = true;
            dw.acc_target = acc_jobs_launched;
            launch_schur_only(respect_done);
        } else { dw.acc_target = 0; launch_accumulate(respect_done); launch_schur_only(respect_done); }
    }
    void join_side() {
        if (!side_pending) return;
        if (cudaEventRecord(ev_join, side_stream) != cudaSuccess || cudaStreamWaitEvent(stream, ev_join, 0) != cudaSuccess) { set_error("join failed"); launch_rc = CMLBA_ERR_CUDA; }
        side_pending = false;
    }
    // Schur complement + stitching + assembly of sys = [HA | bA | H_sc | b_sc]: one cluster kernel when a cluster of >= N CTAs is available
    bool tail_fused() const { return dw.N <= tail_cluster_max && dw.n_sc_chunks > 0; }
    void launch_tail(int respect_done) {
        if (!tail_fused()) { launch_schur(respect_done); launch_stitch(respect_done); return; }
        launch_pdl(4, accumulate_kernel, dim3(dw.N * dw.N * ACC_SLICES), dim3(32), 0, dw, respect_done);
        const int cs = tail_cluster_max;                 // 16 CTAs per host frame when the device can co-schedule them (twice the Schur parallelism), else 8
        if (dw.p2p_on) dw.p2p_epoch = ++p2p_epoch;       // the same count on every rank: one exchange per stitched system
        cudaLaunchConfig_t lc = {}; cudaLaunchAttribute at[1];
        lc.gridDim = dim3(dw.N * cs); lc.blockDim = dim3(TAIL_THREADS); lc.dynamicSmemBytes = tail_smem_bytes(dw.N); lc.stream = stream;
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        if (cudaLaunchKernelEx(&lc, tail_kernel, dw, respect_done) != cudaSuccess) { set_error(std::string("tail_kernel launch: ") + cudaGetErrorString(cudaGetLastError())); }
        launches++;
    }
    // one CTA per ordered frame pair, then the gather into sys = [HA | bA | H_sc | b_sc]
    void launch_stitch_only(int respect_done) { launch_pdl(16, stitch_pair_kernel, dim3(dw.N * dw.N), dim3(ST_THREADS), stitch_smem(), dw, respect_done); }
    void launch_assemble(int respect_done) {
        const int n = dw.n;
        if (dw.p2p_on) dw.p2p_epoch = ++p2p_epoch;       // the same count on every rank: one exchange per stitched system
        launch_pdl(32, assemble_kernel, dim3((2 * n * n + 2 * n + 255) / 256 + 3), dim3(256), 0, dw, respect_done);   // +3 CTAs: 20 warps for HA[C,C], bA[C]
    }
    void launch_stitch(int respect_done) { launch_stitch_only(respect_done); launch_assemble(respect_done); }
    int launch_solve_sequence(int respect_done) {
        launch_tail(respect_done);
        if (world > 1 && !dw.p2p_on) { int rc = allreduce_system(); if (rc) return rc; }      // with peer memory solve_kernel sums the ranks itself
        launch_pdl(64, solve_kernel, dim3(1), dim3(256), solve_smem(), dw, respect_done);
        if (dw.P > 0) launch_pdl(128, point_step_kernel, dim3(dw.n_pt_blocks), dim3(256), 0, dw, respect_done);
        return CMLBA_OK;
    }

    // status of a finished run / stage from the device control block: a peer exchange that timed out (a rank never reached it) is a
    // protocol / state error, not the reference's "run() returned false"
    int exchange_status(const Ctrl &c) {
        if (c.pad0) { set_error("peer-memory exchange timed out: a rank did not reach the exchange (every rank must call the same sequence)"); return CMLBA_ERR_STATE; }
        if (c.pad1) { set_error("internal: the side-stream accumulation did not finish before the stitch gave up waiting for it"); return CMLBA_ERR_STATE; }
        if (c.failed) { set_error("non-finite energy or step (reference run() returns false)"); return CMLBA_ERR_NUMERIC; }
        return CMLBA_OK;
    }

    int allreduce_system() {
        if (dw.p2p_on) { launch_pdl(256, p2p_allreduce_kernel, dim3((2 * dw.n * dw.n + 2 * dw.n + 255) / 256), dim3(256), 0, dw, 0); return CMLBA_OK; }
        const size_t cnt = (size_t) 2 * dw.n * dw.n + 2 * dw.n;
        const int rc = g_nccl.AllReduce(d_sys.p, d_sys.p, cnt, /*ncclDouble*/ 8, /*ncclSum*/ 0, comm, stream);
        if (rc != 0) { set_error("ncclAllReduce failed"); return CMLBA_ERR_CUDA; }
        return CMLBA_OK;
    }

    int run(const double *cams, int iterations, int update_points_only, cmlba_run_result *out) {
        TSCOPE("run(total)");
        Lap lap(timers);
        int rc = prepare(cams);
        if (rc) return rc;
        if (iterations <= 0) iterations = cfg.iterations;
        dw.update_points_only = update_points_only ? 1 : 0;
        const int l0 = launches;
        launch_rc = CMLBA_OK;
        if (world > 1 && comm) {   // align the ranks on the device before the first peer exchange of this run (NCCL has no timeout; the exchanges then only bridge in-run skew)
            CK(d_cap.reserve(1));
            if (g_nccl.AllReduce(d_cap.p, d_cap.p, 1, /*ncclInt32*/ 2, /*ncclMax*/ 2, comm, stream) != 0) { set_error("ncclAllReduce (rank alignment) failed"); return CMLBA_ERR_CUDA; }
        }
        if (want_ktrace) { CK(d_ktrace.reserve(3 * 128 + 32)); dw.ktrace = d_ktrace.p; dw.ktrace_base = d_ktrace.p; kt_seq = 0; kt_sites.clear(); ktrace_reset_kernel<<<1, 96, 0, stream>>>(d_ktrace.p); }
        CK(cudaEventRecord(ev0, stream));
        if (!cfg.force_accept) { point_prior_energy_kernel<<<1, 256, 0, stream>>>(dw); launches++; }
        launch_linearize(0, 0); launch_post(0, 0);
        // GN iterations are enqueued two at a time (run() cannot stop before it = 1, BA:879); between batches the host peeks at
        // Ctrl.done instead of enqueueing up to 7 no-op launches per iteration that the window no longer needs.  On one GPU the closing
        // sequence (new FEJ point of the newest frame, pair constants, linearizeAll(true), BA:885-905) is launched AHEAD of every peek with
        // guard 2 (it runs only if the loop is done): when the host then sees Ctrl.done the device has already finished the run, instead of
        // idling through the host round trip and four launches (~40 us of a 0.37 ms run at C2).  With several ranks the closing sequence
        // contains a peer exchange whose epochs must not be skipped, so it is enqueued after the loop as before.
        auto launch_closing = [&](int guard) {
            set_evalpt_newest_kernel<<<1, 32, 0, stream>>>(dw, guard); launches++;
            pairs_kernel<<<(dw.N * dw.N + 63) / 64, 64, 0, stream>>>(dw, guard); launches++;
            launch_linearize(1, guard); launch_post(2, guard);
            cudaEventRecord(ev1, stream);             // end of the device work of this run (the last record wins)
        };
        CK(peek_h.reserve(1));
        bool closed = false;
        for (int it = 0; it < iterations;) {
            const int batch_end = std::min(iterations, it + 2);
            for (; it < batch_end; it++) {
                rc = launch_solve_sequence(1); if (rc) return rc;
                launch_linearize(0, 1); launch_post(1, 1);
                if (!cfg.force_accept) launch_pdl(128, restore_state_kernel, dim3(std::max(dw.n_pt_blocks, 1)), dim3(256), 0, dw, it + 1);   // no-op unless the step was rejected
            }
            if (it < iterations) {
                if (world == 1) launch_closing(2);
                CK(cudaMemcpyAsync(peek_h.p, reinterpret_cast<const char *>(d_ctrl.p) + offsetof(Ctrl, done), sizeof(int), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                if (*peek_h.p) { closed = world == 1; break; }
            }
        }
        if (!closed) launch_closing(0);
        CK(cudaGetLastError());
        if (launch_rc) return launch_rc;
        lap("run.launch");
        rc = finish_run(out);
        if (out) {
            float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1);
            out->gpu_ms = ms; out->kernel_launches = launches - l0;
        }
        return rc;
    }

    // read results back, update the host bookkeeping (the "scatter" edge of the boundary)
    int finish_run(cmlba_run_result *out) {
        TSCOPE("finish_run");
        Lap flap(timers);
        const int N = dw.N, P = dw.P, R = dw.R;
        // one pinned block: frames | ctrl | idepth | idz idh mrb | ng | alive st0 st1 | en0 en1
        size_t off = 0;
        auto take = [&](size_t nb) { const size_t o = (off + 15) & ~(size_t) 15; off = o + nb; return o; };
        const size_t o_fd = take(N * sizeof(FrameDev)), o_c = take(sizeof(Ctrl)), o_id = take((size_t) P * 8), o_idz = take((size_t) P * 4), o_idh = take((size_t) P * 4),
                     o_mrb = take((size_t) P * 4), o_ng = take((size_t) P * 4), o_al = take(R), o_s0 = take(R), o_e0 = take((size_t) R * 4);
        CK(fin_h.reserve(off));
        join_side();
        char *hb = fin_h.p;
        CK(cudaMemcpyAsync(hb + o_fd, d_frames.p, N * sizeof(FrameDev), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(hb + o_c, d_ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
        if (P) {
            CK(cudaMemcpyAsync(hb + o_id, d_pt_idepth.p, (size_t) P * 8, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_idz, d_pt_idz.p, (size_t) P * 4, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_idh, d_pt_idh.p, (size_t) P * 4, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_mrb, d_pt_mrb.p, (size_t) P * 4, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_ng, d_pt_num_good.p, (size_t) P * 4, cudaMemcpyDeviceToHost, stream));
        }
        if (R) {
            // final residual states in the host's own order (written by the fixLinearization pass)
            CK(cudaMemcpyAsync(hb + o_al, d_fin_alive.p, R, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_s0, d_fin_state.p, R, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(hb + o_e0, d_fin_energy.p, (size_t) R * 4, cudaMemcpyDeviceToHost, stream));
        }
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        flap("finish.wait_d2h");
        const FrameDev *fd = reinterpret_cast<const FrameDev *>(hb + o_fd);
        const Ctrl c = *reinterpret_cast<const Ctrl *>(hb + o_c);
        const double *id = reinterpret_cast<const double *>(hb + o_id);
        const float *idz = reinterpret_cast<const float *>(hb + o_idz), *idh = reinterpret_cast<const float *>(hb + o_idh), *mrb = reinterpret_cast<const float *>(hb + o_mrb);
        const int *ng = reinterpret_cast<const int *>(hb + o_ng);
        const uint8_t *alive = reinterpret_cast<const uint8_t *>(hb + o_al);
        const uint8_t *st = reinterpret_cast<const uint8_t *>(hb + o_s0);
        const float *en = reinterpret_cast<const float *>(hb + o_e0);
        for (int i = 0; i < N; i++) {
            FrameHost &f = frames_[i]; const FrameDev &d = fd[i];
            for (int k = 0; k < 10; k++) { f.state[k] = d.state[k]; f.state_zero[k] = d.state_zero[k]; }
            for (int k = 0; k < 9; k++) { f.evalpt.R[k] = d.evalR[k]; f.pre.R[k] = d.preR[k]; }
            for (int k = 0; k < 3; k++) { f.evalpt.t[k] = d.evalt[k]; f.pre.t[k] = d.pret[k]; }
            f.energy_th = d.energy_th;
            f.aff_a = d.state_scaled[6]; f.aff_b = d.state_scaled[7];
        }
        // residuals (device order).  Dropped ones leave the masks; lastResiduals bookkeeping (BA:1616-1620, 1630-1633) only
        // concerns residuals towards the two newest frames (the only ids PointHost::last_frame can hold).
        const int *r_point = up.host<int>(up_o_rp); const uint8_t *r_target = up.host<uint8_t>(up_o_rt);
        int dropped = 0;
        {
            int i = 0;
            for (; i + 8 <= R; i += 8) {            // alive[] is almost all ones: test 8 flags at a time
                uint64_t wd; memcpy(&wd, alive + i, 8);
                if (wd == 0x0101010101010101ull) continue;
                for (int k = i; k < i + 8; k++) if (!alive[k]) { points_[pt_order[r_point[k]]].res_mask &= (uint16_t) ~(1u << r_target[k]); frames_[r_target[k]].num_residuals_out++; dropped++; }
            }
            for (; i < R; i++) if (!alive[i]) { points_[pt_order[r_point[i]]].res_mask &= (uint16_t) ~(1u << r_target[i]); frames_[r_target[i]].num_residuals_out++; dropped++; }
        }
        for (int t = std::max(0, N - 2); t < N; t++) {
            const int64_t tid = frames_[t].id;
            for (int i = res_bin_begin[t * N]; i < res_bin_begin[(t + 1) * N]; i++) {
                PointHost &p = points_[pt_order[r_point[i]]];
                for (int s2 = 0; s2 < 2; s2++) if (p.last_frame[s2] == tid) { p.last_state[s2] = st[i]; if (!alive[i]) p.last_frame[s2] = -1; break; }   // setResidualState for every active residual, then first = nullptr for the deleted ones
            }
        }
        // points: results, outliers (points left without residuals, BA:1636-1640) and the id list of the residual snapshot
        snap.valid = true; snap.own_map = false; snap.R = R;
        snap.state = st; snap.alive = alive; snap.energy = en;
        snap.frame_id.resize(N); for (int i = 0; i < N; i++) snap.frame_id[i] = frames_[i].id;
        snap.point_id.resize(P);
        outliers_.clear();
        int nout = 0;
        for (int i = 0; i < P; i++) {
            const int q = pt_order[i];
            PointHost &p = points_[q];
            p.idepth = id[i]; p.idepth_zero = idz[i]; p.idepth_hessian = idh[i]; p.max_rel_bs = mrb[i]; p.num_good = ng[i];
            snap.point_id[i] = p.id;
            if (p.res_mask == 0) { kill_point(q); outliers_.push_back(p.id); nout++; }
        }
        {   // Statistic series (BA:798-802, 847-851, 1415-1425, 2204): the values of the last iteration
            const double nres = std::max(R, 1);
            const double eM = 0.0;                                      // calcMEnergy is folded into the linearised energy here (post_linearize_kernel)
            stats_[0] = c.energy_last / nres; stats_[1] = 0.0; stats_[2] = c.energyL_last; stats_[3] = eM;
            stats_[4] = (c.energy_last + c.energyL_last + eM) / nres;
            stats_[5] = c.stats[4];                                     // X Norm
            stats_[6] = c.stats[5]; stats_[7] = c.stats[6]; stats_[8] = c.stats[7]; stats_[9] = c.stats[8];          // Hessian P / L / M / SC norms
            stats_[10] = c.stats[9]; stats_[11] = c.stats[10]; stats_[12] = c.stats[11]; stats_[13] = c.stats[12];   // B norms (declared, never fed by the reference)
            stats_[18] = 0.0;                                           // Num Linearized: run() never holds a linearised residual
        }
        if (out) {
            out->iterations_done = c.iteration; out->num_residuals = R; out->num_dropped = dropped; out->num_outliers = nout;
            out->energy_first = c.energy_first; out->energy_last = c.energy_last; out->num_rejected = c.rejected;
        }
        flap("finish.scatter");
        dirty = dropped > 0 || nout > 0;
        if (n_dead * 4 > points_.size()) compact();
        flap("finish.compact");
        prepared = false;
        return exchange_status(c);
    }

    // drop the whole window but keep the allocations (used by streaming callers and the end-to-end benchmark)
    int reset() {
        cudaSetDevice(device);
        for (auto &f : frames_) if (f.d_img) { img_pool.push_back(f.d_img); f.d_img = nullptr; }
        frames_.clear(); points_.clear(); n_dead = 0; HM.clear(); bM.clear(); snap.valid = false; point_index_.clear(); ids_sorted = true; outliers_.clear();
        key_counter = 0; dirty = true; prepared = false;
        return CMLBA_OK;
    }

    // Benchmark helper: `steps` passes of the Jacobian + Schur accumulation hot path
    // (linearize -> accumulate -> schur -> stitch) on the prepared window, CUDA events on the launching stream.
    // flush_l2: evict the window between passes by writing a buffer larger than L2 (outside the timed region).
    int bench_pass(int steps, int warmup, int flush_l2, cmlba_bench_result *out) {
        if (!prepared) { set_error("cmlba_prepare first"); return CMLBA_ERR_STATE; }
        if (!out || steps < 1) { set_error("bad arguments"); return CMLBA_ERR_ARG; }
        CK(cudaSetDevice(device));
        const size_t flush_bytes = (size_t) 384 << 20;
        if (flush_l2) CK(d_flush.reserve(flush_bytes / sizeof(float4)));
        cudaEvent_t ev[7];
        for (auto &e : ev) CK(cudaEventCreate(&e));
        // after the flush the ranks are re-aligned on the device (peer-memory barrier, outside the timed region): the flush lengths differ per rank
        auto flush = [&]() {
            if (flush_l2) l2_flush_kernel<<<148 * 8, 256, 0, stream>>>(d_flush.p, flush_bytes / sizeof(float4));
            if (dw.p2p_on) { dw.p2p_epoch = ++p2p_epoch; p2p_barrier_kernel<<<1, 32, 0, stream>>>(dw); }
        };
        // the committed buffers must hold a linearization for schur/stitch to chew on
        launch_linearize(0, 0); launch_post(0, 0);
        for (int i = 0; i < warmup; i++) { flush(); launch_linearize(0, 0); launch_tail(0); }
        CK(cudaStreamSynchronize(stream));
        double tot = 0, tk[6] = {0, 0, 0, 0, 0, 0};
        const int l0 = launches;
        for (int i = 0; i < steps; i++) {          // whole pass, two events only
            flush();
            if (want_ktrace) { CK(d_ktrace.reserve(3 * 128 + 32)); dw.ktrace = d_ktrace.p; dw.ktrace_base = d_ktrace.p; kt_seq = 0; kt_sites.clear(); ktrace_reset_kernel<<<1, 96, 0, stream>>>(d_ktrace.p); }
            CK(cudaEventRecord(ev[0], stream));
            launch_linearize(0, 0);
            launch_tail(0);
            if (world > 1) { int rc = allreduce_system(); if (rc) return rc; }
            CK(cudaEventRecord(ev[1], stream));
            CK(cudaStreamSynchronize(stream));
            CK(cudaStreamSynchronize(side_stream));
            float ms = 0; CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); tot += ms;
        }
        const int per_pass = (launches - l0) / steps;
        // per-kernel breakdown: the same kernels SERIALISED on the main stream (no fork), one event-to-event interval each; every interval
        // carries one event-record overhead, measured by the empty interval ev[1]..ev[2]
        const bool fork_saved = use_fork; use_fork = false;
        unsigned long long *kt_saved = dw.ktrace; dw.ktrace = nullptr;       // the timeline keeps the last whole pass
        for (int i = 0; i < steps; i++) {
            flush();
            CK(cudaEventRecord(ev[0], stream)); launch_linearize(0, 0);
            CK(cudaEventRecord(ev[1], stream));
            CK(cudaEventRecord(ev[2], stream));
            if (tail_fused()) { launch_tail(0); for (int k = 3; k < 7; k++) CK(cudaEventRecord(ev[k], stream)); }      // fused: the whole tail is reported as ms_accumulate
            else {
                launch_accumulate(0); CK(cudaEventRecord(ev[3], stream));
                launch_schur_only(0); CK(cudaEventRecord(ev[4], stream));
                launch_stitch_only(0); CK(cudaEventRecord(ev[5], stream));
                launch_assemble(0); CK(cudaEventRecord(ev[6], stream));
            }
            CK(cudaStreamSynchronize(stream));
            for (int k = 0; k < 6; k++) { float ms = 0; CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); tk[k] += ms; }
        }
        use_fork = fork_saved; dw.ktrace = kt_saved;
        CK(cudaGetLastError());
        for (auto &e : ev) cudaEventDestroy(e);
        out->steps = steps; out->residuals = dw.R; out->points = dw.P; out->frames = dw.N;
        out->ms_pass = tot / steps; out->ms_linearize = tk[0] / steps; out->ms_event_overhead = tk[1] / steps; out->ms_accumulate = tk[2] / steps; out->ms_schur = tk[3] / steps;
        out->ms_stitch = tk[4] / steps; out->ms_assemble = tk[5] / steps;
        out->launches_per_pass = per_pass;
        return CMLBA_OK;
    }

    // ------------------------------------------------------------------ peer-memory exchange set-up
    int comm_ipc_handle(void *out64) {
        CK(cudaSetDevice(device));
        if (world < 2) { set_error("cmlba_comm_init first"); return CMLBA_ERR_STATE; }
        if (!p2p_local) { CK(cudaMalloc(&p2p_local, p2p_bytes())); CK(cudaMemset(p2p_local, 0, p2p_bytes())); CK(cudaDeviceSynchronize()); }
        cudaIpcMemHandle_t hd;
        CK(cudaIpcGetMemHandle(&hd, p2p_local));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(out64, &hd, 64);
        return CMLBA_OK;
    }
    int comm_ipc_open(const void *all) {
        if (world < 2 || !p2p_local) { set_error("cmlba_comm_init and cmlba_comm_ipc_handle first"); return CMLBA_ERR_STATE; }
        CK(cudaSetDevice(device));
        for (int q = 0; q < world; q++) {
            if (q == rank) { p2p_peer[q] = p2p_local; continue; }
            cudaIpcMemHandle_t hd; memcpy(&hd, (const char *) all + 64 * q, 64);
            void *ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
            p2p_peer[q] = (char *) ptr;
        }
        p2p_ready = true; dirty = true; prepared = false;
        return CMLBA_OK;
    }

    // ------------------------------------------------------------------ named read-back (parity tests)
    template <typename T> int copy_out(const T *dev, size_t count, void *dst, size_t cap, size_t *bytes) {
        const size_t nb = count * sizeof(T);
        if (bytes) *bytes = nb;
        if (dst && cap) { CK(cudaMemcpy(dst, dev, std::min(cap, nb), cudaMemcpyDeviceToHost)); }
        return CMLBA_OK;
    }
    int host_out(const void *src, size_t nb, void *dst, size_t cap, size_t *bytes) {
        if (bytes) *bytes = nb;
        if (dst && cap) memcpy(dst, src, std::min(cap, nb));
        return CMLBA_OK;
    }
    int read(const std::string &name, void *dst, size_t cap, size_t *bytes) {
        if (name == "host_timing") { const std::string t = timers.text() + "device allocations " + std::to_string(g_dev_mallocs) + ", pinned allocations " + std::to_string(g_pin_mallocs) + "\n"; return host_out(t.data(), t.size(), dst, cap, bytes); }
        if (name == "host_timing_reset") { timers.acc.clear(); if (bytes) *bytes = 0; return CMLBA_OK; }
        if (name.rfind("image", 0) == 0) {   // "image<slot>": the float4 texels (I, dx, dy, 0) of a window frame
            const int slot = atoi(name.c_str() + 5);
            if (slot < 0 || slot >= (int) frames_.size() || !frames_[slot].d_img) { set_error("no such frame slot"); return CMLBA_ERR_ARG; }
            CK(cudaSetDevice(device)); CK(cudaStreamSynchronize(stream));
            return copy_out(frames_[slot].d_img, (size_t) W * H, dst, cap, bytes);
        }
        if (name == "HM") return host_out(HM.data(), HM.size() * 8, dst, cap, bytes);
        if (name == "bM") return host_out(bM.data(), bM.size() * 8, dst, cap, bytes);
        if (name == "frame_counters") {   // [N][4] int32: flagged, numMarginalized, numResidualsOut, residuals targeting the frame
            std::vector<int32_t> v;
            for (size_t i = 0; i < frames_.size(); i++) { v.push_back(frames_[i].flagged); v.push_back(frames_[i].num_marginalized); v.push_back(frames_[i].num_residuals_out); v.push_back(frame_residual_count((int) i)); }
            return host_out(v.data(), v.size() * 4, dst, cap, bytes);
        }
        if (name == "ktrace" || name == "ktrace_solve") { if (!d_ktrace.p) { set_error("CMLBA_KTRACE=1 first"); return CMLBA_ERR_STATE; } CK(cudaSetDevice(device)); CK(cudaStreamSynchronize(stream)); 
            std::vector<unsigned long long> v(4 * 128, 0);      // [slot][sched, past wait, done, site]
            std::vector<unsigned long long> raw(3 * 128 + 32);
            CK(cudaMemcpy(raw.data(), d_ktrace.p, raw.size() * 8, cudaMemcpyDeviceToHost));
            for (int i = 0; i <= kt_seq && i < 128; i++) { for (int k = 0; k < 3; k++) v[4 * i + k] = raw[3 * i + k]; v[4 * i + 3] = i == 0 ? 0 : (unsigned long long) kt_sites[i - 1]; }
            if (name == "ktrace_solve") return host_out(raw.data() + 3 * 128, 32 * 8, dst, cap, bytes);
            return host_out(v.data(), (size_t) (kt_seq + 1) * 32, dst, cap, bytes); }
        if (name == "enable_dbg") { want_dbg = true; dirty = true; prepared = false; if (bytes) *bytes = 0; return CMLBA_OK; }
        if (dirty || !d_ctrl.p) { set_error("window not built yet (cmlba_prepare / cmlba_run first)"); return CMLBA_ERR_STATE; }
        CK(cudaSetDevice(device));
        CK(cudaStreamSynchronize(stream));
        const int N = dw.N, P = dw.P, R = dw.R, n = dw.n;
        Ctrl c; CK(cudaMemcpy(&c, d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost));
        const int cur = c.cur;
        if (name == "ctrl") return host_out(&c, sizeof(c), dst, cap, bytes);
        if (name == "pt_order") return host_out(pt_order.data(), P * sizeof(int), dst, cap, bytes);
        if (name == "res_point" || name == "res_point_dev" || name == "res_host" || name == "res_target") {   // decoded from the tile-sorted order of the device
            std::vector<uint32_t> pht(R);
            if (R) CK(cudaMemcpy(pht.data(), d_r_pht.p, (size_t) R * 4, cudaMemcpyDeviceToHost));
            if (name == "res_point" || name == "res_point_dev") {
                std::vector<int> v(R);
                for (int i = 0; i < R; i++) { const int dp = (int) (pht[i] & 0xffffffu); v[i] = name == "res_point" ? pt_order[dp] : dp; }
                return host_out(v.data(), R * sizeof(int), dst, cap, bytes);
            }
            std::vector<uint8_t> v(R);
            for (int i = 0; i < R; i++) v[i] = (uint8_t) (name == "res_host" ? ((pht[i] >> 24) & 15u) : (pht[i] >> 28));
            return host_out(v.data(), R, dst, cap, bytes);
        }
        if (name == "lt_trace") return copy_out(d_lt_trace.p, (size_t) dw.lt_grid * 16 * 32, dst, cap, bytes);
        if (name == "res_src") return copy_out(d_r_src.p, R, dst, cap, bytes);
        if (name == "res_job") return copy_out(d_r_job.p, R, dst, cap, bytes);
        if (name == "res_state") return copy_out(cur ? d_r_state1.p : d_r_state0.p, R, dst, cap, bytes);
        if (name == "res_state_cand") return copy_out(cur ? d_r_state0.p : d_r_state1.p, R, dst, cap, bytes);
        if (name == "res_energy") return copy_out(cur ? d_r_energy1.p : d_r_energy0.p, R, dst, cap, bytes);
        if (name == "res_good") return copy_out(cur ? d_r_good1.p : d_r_good0.p, R, dst, cap, bytes);
        if (name == "res_good_cand") return copy_out(cur ? d_r_good0.p : d_r_good1.p, R, dst, cap, bytes);
        if (name == "res_new_state") return copy_out(d_r_new_state.p, R, dst, cap, bytes);
        if (name == "res_new_energy") return copy_out(d_r_new_energy.p, R, dst, cap, bytes);
        if (name == "res_new_energy_wo") return copy_out(d_r_new_energy_wo.p, R, dst, cap, bytes);
        if (name == "res_alive") return copy_out(d_r_alive.p, R, dst, cap, bytes);
        if (name == "res_center") return copy_out(d_r_center.p, (size_t) R * 3, dst, cap, bytes);
        if (name == "rj") {   // candidate records (the last linearization), re-ordered from the host's residual order to the device's tile-sorted one
            std::vector<float> rec((size_t) R * RJ_STRIDE), out((size_t) R * RJ_STRIDE, 0.f);
            std::vector<int> srcv(R);
            if (R) { CK(cudaMemcpy(rec.data(), cur ? d_rj0.p : d_rj1.p, rec.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(srcv.data(), d_r_src.p, (size_t) R * 4, cudaMemcpyDeviceToHost)); }
            for (int i = 0; i < R; i++) if (rec[(size_t) srcv[i] * RJ_STRIDE + 35] != 0.f) memcpy(&out[(size_t) i * RJ_STRIDE], &rec[(size_t) srcv[i] * RJ_STRIDE], 35 * sizeof(float));
            return host_out(out.data(), out.size() * 4, dst, cap, bytes);
        }
        if (name == "dbg") { if (!want_dbg) { set_error("debug dump not enabled (cmlba_read(\"enable_dbg\") first)"); return CMLBA_ERR_STATE; } return copy_out(d_dbg.p, (size_t) R * DBG_STRIDE, dst, cap, bytes); }
        if (name == "T") return copy_out(cur ? d_T1.p : d_T0.p, (size_t) P * N * T_STRIDE, dst, cap, bytes);
        if (name == "T_cand") return copy_out(cur ? d_T0.p : d_T1.p, (size_t) P * N * T_STRIDE, dst, cap, bytes);
        if (name == "pt_idepth") return copy_out(d_pt_idepth.p, P, dst, cap, bytes);
        if (name == "pt_step") return copy_out(d_pt_step.p, P, dst, cap, bytes);
        if (name == "pt_Hdd") return copy_out(d_pt_Hdd.p, P, dst, cap, bytes);
        if (name == "pt_bd") return copy_out(d_pt_bd.p, P, dst, cap, bytes);
        if (name == "pt_Hcd") return copy_out(d_pt_Hcd.p, (size_t) P * 4, dst, cap, bytes);
        if (name == "pt_HdiF") return copy_out(d_pt_HdiF.p, P, dst, cap, bytes);
        if (name == "pt_bdSumF") return copy_out(d_pt_bdSumF.p, P, dst, cap, bytes);
        if (name == "pt_idepth_hessian") return copy_out(d_pt_idh.p, P, dst, cap, bytes);
        if (name == "pt_max_rel_baseline") return copy_out(d_pt_mrb.p, P, dst, cap, bytes);
        if (name == "pt_num_good") return copy_out(d_pt_num_good.p, P, dst, cap, bytes);
        if (name == "pt_colors") return copy_out(d_pt_colors.p, (size_t) P * 8, dst, cap, bytes);
        if (name == "pt_weights") return copy_out(d_pt_weights.p, (size_t) P * 8, dst, cap, bytes);
        if (name == "frames") return copy_out(d_frames.p, N, dst, cap, bytes);
        if (name == "pairs") return copy_out(d_pairs.p, (size_t) N * N, dst, cap, bytes);
        if (name == "AH") return copy_out(d_AH.p, (size_t) N * N * 64, dst, cap, bytes);
        if (name == "AT") return copy_out(d_AT.p, (size_t) N * N * 64, dst, cap, bytes);
        if (name == "Pns") return copy_out(d_Pns.p, (size_t) n * n, dst, cap, bytes);
        if (name == "sys") return copy_out(d_sys.p, (size_t) 2 * n * n + 2 * n, dst, cap, bytes);
        if (name == "x") return copy_out(d_x.p, n, dst, cap, bytes);
        if (name == "xAd") return copy_out(d_xAd.p, (size_t) N * N * 8, dst, cap, bytes);
        if (name == "acc" || name == "acc_cand") {   // per bin (t*N+h) packed 96 doubles: addToHessianTop evaluated from the Jacobian records (host order = bin-major)
            const bool cand = name == "acc_cand";
            std::vector<float> rec((size_t) R * RJ_STRIDE);
            if (R) CK(cudaMemcpy(rec.data(), (cur ^ (cand ? 1 : 0)) ? d_rj1.p : d_rj0.p, rec.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<double> acc((size_t) N * N * ACC_N, 0.0);
            for (int b = 0; b < N * N; b++) for (int r = res_bin_begin[b]; r < res_bin_begin[b + 1]; r++) {
                const float *q = &rec[(size_t) r * RJ_STRIDE];
                if (q[35] == 0.f) continue;
                float Qx[10], Qy[10];
                for (int k = 0; k < 10; k++) { Qx[k] = q[20] * q[k] + q[21] * q[10 + k]; Qy[k] = q[21] * q[k] + q[22] * q[10 + k]; }
                for (int e = 0; e < 91; e++) acc[(size_t) b * ACC_N + e] += (double) acc_entry(e, q, q + 10, Qx, Qy, q + 23, q + 26, q + 29);
            }
            return host_out(acc.data(), acc.size() * 8, dst, cap, bytes);
        }
        if (name == "sc") {   // per host: D[(8N)^2] E[32N] EB[8N] Hcc[16] bc[4] doubles
            const int NB = 8 * N, tot = NB * NB + NB * 5 + 20;
            std::vector<float> part((size_t) n_sc_chunks * dw.sc_stride);
            if (n_sc_chunks) CK(cudaMemcpy(part.data(), d_sc_part.p, part.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<double> s((size_t) N * tot, 0.0);
            for (int h = 0; h < N; h++) for (int ch = h_host_chunk_begin[h]; ch < h_host_chunk_begin[h + 1]; ch++) for (int k = 0; k < tot; k++) s[(size_t) h * tot + k] += part[(size_t) ch * dw.sc_stride + k];
            return host_out(s.data(), s.size() * 8, dst, cap, bytes);
        }
        set_error("unknown buffer name: " + name);
        return CMLBA_ERR_ARG;
    }
};

}  // namespace cmlba

// =================================================================================================
using cmlba::Engine;
struct cmlba_handle { Engine eng; };

static thread_local std::string g_create_error;

extern "C" {

const char *cmlba_version(void) { return "libcmlba 0.1 sm_100a"; }

int cmlba_default_config(cmlba_config *c) {
    if (!c) return CMLBA_ERR_ARG;
    c->iterations = 4; c->huber_threshold = 9.f; c->outlier_th_sum = 2500.f; c->th_opt_iterations = 1.2f;
    c->scale_rotation = 1.f; c->scale_translation = 0.5f; c->scale_light_a = 10.f; c->scale_light_b = 1000.f; c->scale_f = 50.f; c->scale_c = 50.f;
    c->force_accept = 1; c->fix_lambda = 1; c->fixed_lambda = 1e-5f; c->idepth_fix_prior = 2500; c->solver_mode_delta = 1e-5f;
    c->optimize_light_a = 1; c->optimize_light_b = 1; c->disable_marginalization = 1; c->max_frames = 6; c->frame_min_age = 1; c->min_idepth_h_marg = 50.f; c->async_image_upload = 0;
    return CMLBA_OK;
}

int cmlba_create(const cmlba_config *cfg, int device, cmlba_handle **out) {
    if (!out) return CMLBA_ERR_ARG;
    *out = nullptr;
    cmlba_handle *h = new (std::nothrow) cmlba_handle;
    if (!h) return CMLBA_ERR_ARG;
    if (cfg) h->eng.cfg = *cfg; else cmlba_default_config(&h->eng.cfg);
    h->eng.device = device;
    int rc = h->eng.init();
    if (rc) { g_create_error = h->eng.err; delete h; return rc; }
    *out = h;
    return CMLBA_OK;
}

int cmlba_destroy(cmlba_handle *h) { delete h; return CMLBA_OK; }

const char *cmlba_last_error(const cmlba_handle *h) { return h ? h->eng.err.c_str() : g_create_error.c_str(); }

#define HCHK if (!h) return CMLBA_ERR_ARG

int cmlba_set_calib(cmlba_handle *h, double fx, double fy, double cx, double cy, int w, int hh) { HCHK; return h->eng.set_calib(fx, fy, cx, cy, w, hh); }
int cmlba_add_frame(cmlba_handle *h, int64_t id, const double w2c[12], double a, double b, double exposure, const float *grad, int is_init) { HCHK; return h->eng.add_frame(id, w2c, a, b, exposure, grad, is_init); }
int cmlba_add_frame_gray(cmlba_handle *h, int64_t id, const double w2c[12], double a, double b, double exposure, const float *gray, int is_init) { HCHK; return h->eng.add_frame(id, w2c, a, b, exposure, gray, is_init, 1); }
int cmlba_add_frame_device(cmlba_handle *h, int64_t id, const double w2c[12], double a, double b, double exposure, const void *d_texels, int is_init) {
    HCHK; return h->eng.add_frame(id, w2c, a, b, exposure, static_cast<const float *>(d_texels), is_init, 2);
}
int cmlba_add_points(cmlba_handle *h, int n, const int64_t *pid, const int64_t *host, const float *xy, const double *idepth) { HCHK; return h->eng.add_points(n, pid, host, xy, idepth); }
int cmlba_remove_point(cmlba_handle *h, int64_t id) { HCHK; return h->eng.remove_point(id); }
int cmlba_remove_frame(cmlba_handle *h, int64_t id) { HCHK; return h->eng.remove_frame(id); }
int cmlba_flag_frames_for_marginalization(cmlba_handle *h, const double *cams, const int32_t *num_immature, int64_t *ids, int *n) { HCHK; return h->eng.flag_frames(cams, num_immature, ids, n); }
int cmlba_try_marginalize(cmlba_handle *h, int *n_dropped, int *n_to_marg) { HCHK; return h->eng.try_marginalize(n_dropped, n_to_marg); }
int cmlba_marginalize_points(cmlba_handle *h, int64_t *ids, int *n) { HCHK; return h->eng.marginalize_points(ids, n); }
int cmlba_marginalize_frames(cmlba_handle *h, int64_t *ids, int *n) { HCHK; return h->eng.marginalize_frames(ids, n); }
int cmlba_run(cmlba_handle *h, const double *cams, int iterations, int upo, cmlba_run_result *r) { HCHK; return h->eng.run(cams, iterations, upo, r); }
int cmlba_num_frames(const cmlba_handle *h) { return h ? (int) h->eng.frames_.size() : CMLBA_ERR_ARG; }
int cmlba_num_points(const cmlba_handle *h) { return h ? (int) (h->eng.points_.size() - h->eng.n_dead) : CMLBA_ERR_ARG; }
int cmlba_num_residuals(const cmlba_handle *h) {
    if (!h) return CMLBA_ERR_ARG;
    int n = 0;
    for (auto &p : h->eng.points_) if (p.alive) n += __builtin_popcount(p.res_mask);
    return n;
}

int cmlba_get_frames(const cmlba_handle *h, int64_t *id, double *w2c, double *ab, double *state, double *evalpt, double *th) {
    HCHK;
    const auto &fr = h->eng.frames_;
    for (size_t i = 0; i < fr.size(); i++) {
        if (id) id[i] = fr[i].id;
        if (w2c) { memcpy(w2c + 12 * i, fr[i].pre.R, 72); memcpy(w2c + 12 * i + 9, fr[i].pre.t, 24); }
        if (ab) { ab[2 * i] = fr[i].aff_a; ab[2 * i + 1] = fr[i].aff_b; }
        if (state) memcpy(state + 10 * i, fr[i].state, 80);
        if (evalpt) { memcpy(evalpt + 12 * i, fr[i].evalpt.R, 72); memcpy(evalpt + 12 * i + 9, fr[i].evalpt.t, 24); }
        if (th) th[i] = fr[i].energy_th;
    }
    return CMLBA_OK;
}

int cmlba_get_points(const cmlba_handle *h, int64_t *id, double *idepth, double *unc, float *idh, float *mrb, int32_t *ng, int32_t *gft) {
    HCHK;
    const auto &pts = h->eng.points_;
    size_t i = 0;
    for (size_t q = 0; q < pts.size(); q++) {
        const auto &p = pts[q];
        if (!p.alive) continue;
        if (id) id[i] = p.id;
        if (idepth) idepth[i] = p.idepth;
        if (unc) unc[i] = 1.0 / ((double) p.idepth_hessian + 0.01);    // updatePointUncertainty (DSOPoint.h:107-118)
        if (idh) idh[i] = p.idepth_hessian;
        if (mrb) mrb[i] = p.max_rel_bs;
        if (ng) ng[i] = p.num_good;
        if (gft) gft[i] = (p.last_frame[0] >= 0 && p.last_state[0] == CMLBA_RES_IN) ? 1 : 0;   // getGoodPointsForTracking (BA.h:76-85)
        i++;
    }
    return CMLBA_OK;
}

int cmlba_get_outliers(const cmlba_handle *h, int64_t *id, int *n) {
    HCHK; if (!n) return CMLBA_ERR_ARG;
    const auto &o = h->eng.outliers_;
    const int cap = *n;
    *n = (int) o.size();
    if (id) for (int i = 0; i < cap && i < (int) o.size(); i++) id[i] = o[i];
    return CMLBA_OK;
}

int cmlba_get_residuals(const cmlba_handle *h, int64_t *pid, int64_t *tid, int32_t *state, double *energy) {
    HCHK;
    const Engine &e = h->eng;
    // (point id, frame id) -> device residual index of the last run(), decoded from the snapshot on demand (O(R))
    std::unordered_map<int64_t, std::unordered_map<int64_t, int>> last;
    if (e.snap.valid) {
        const int *rp = e.snap.own_map ? e.snap.r_point.data() : reinterpret_cast<const int *>(e.up.h.p + e.up_o_rp);
        const uint8_t *rt = e.snap.own_map ? e.snap.r_target.data() : reinterpret_cast<const uint8_t *>(e.up.h.p + e.up_o_rt);
        for (int i = 0; i < e.snap.R; i++) if (e.snap.alive[i]) last[e.snap.point_id[rp[i]]][e.snap.frame_id[rt[i]]] = i;
    }
    size_t k = 0;
    for (auto &p : e.points_) {
        if (!p.alive) continue;
        auto lp = last.find(p.id);
        for (unsigned m = p.res_mask; m; m &= m - 1, k++) {
            const int64_t f = e.frames_[__builtin_ctz(m)].id;
            int st = CMLBA_RES_IN; double en = 0.0;           // residuals created since the last run (DSOResidual.h:81-86)
            if (lp != last.end()) { auto it = lp->second.find(f); if (it != lp->second.end()) { st = e.snap.state[it->second]; en = e.snap.energy[it->second]; } }
            if (pid) pid[k] = p.id;
            if (tid) tid[k] = f;
            if (state) state[k] = st;
            if (energy) energy[k] = en;
        }
    }
    return CMLBA_OK;
}

// ---- stage entry points
int cmlba_prepare(cmlba_handle *h, const double *cams) { HCHK; return h->eng.prepare(cams); }

int cmlba_linearize(cmlba_handle *h, int fix, double *energy) {
    HCHK; Engine &e = h->eng;
    if (!e.prepared) { e.set_error("cmlba_prepare first"); return CMLBA_ERR_STATE; }
    cudaSetDevice(e.device);
    if (fix) {
        cmlba::set_evalpt_newest_kernel<<<1, 32, 0, e.stream>>>(e.dw, 0);
        cmlba::pairs_kernel<<<(e.dw.N * e.dw.N + 63) / 64, 64, 0, e.stream>>>(e.dw, 0);
        e.launch_linearize(1, 0); e.launch_post(2, 0);
    } else {
        e.launch_linearize(0, 0); e.launch_post(3, 0);
    }
    cmlba::Ctrl c;
    if (cudaMemcpyAsync(&c, e.d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost, e.stream) != cudaSuccess || cudaStreamSynchronize(e.stream) != cudaSuccess) {
        e.set_error(std::string("linearize: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA;
    }
    if (energy) *energy = c.energy_new;
    if (e.launch_rc) { const int rc = e.launch_rc; e.launch_rc = CMLBA_OK; return rc; }
    return e.exchange_status(c);
}

int cmlba_apply(cmlba_handle *h) {
    HCHK; Engine &e = h->eng;
    if (!e.prepared) { e.set_error("cmlba_prepare first"); return CMLBA_ERR_STATE; }
    cudaSetDevice(e.device);
    cmlba::Ctrl c;
    if (cudaMemcpy(&c, e.d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) { e.set_error(std::string("apply: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA; }
    c.cur ^= 1; c.energy_last = c.energy_new;
    if (cudaMemcpy(e.d_ctrl.p, &c, sizeof(c), cudaMemcpyHostToDevice) != cudaSuccess) { e.set_error(std::string("apply: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA; }
    return CMLBA_OK;
}

int cmlba_solve(cmlba_handle *h, int iteration) {
    HCHK; Engine &e = h->eng;
    if (!e.prepared) { e.set_error("cmlba_prepare first"); return CMLBA_ERR_STATE; }
    cudaSetDevice(e.device);
    cmlba::Ctrl c;
    if (cudaMemcpy(&c, e.d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) { e.set_error(std::string("solve: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA; }
    c.iteration = iteration;
    if (cudaMemcpy(e.d_ctrl.p, &c, sizeof(c), cudaMemcpyHostToDevice) != cudaSuccess) { e.set_error(std::string("solve: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA; }
    int rc = e.launch_solve_sequence(0);
    if (rc) return rc;
    if (cudaStreamSynchronize(e.stream) != cudaSuccess || cudaMemcpy(&c, e.d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) {
        e.set_error(std::string("solve: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA;
    }
    return e.exchange_status(c);
}

int cmlba_step(cmlba_handle *h, int update_points_only, int *can_break) {
    HCHK; Engine &e = h->eng;
    (void) update_points_only;   // the step is applied on the device inside cmlba_solve (solve_kernel / point_step_kernel)
    if (!e.prepared) { e.set_error("cmlba_prepare first"); return CMLBA_ERR_STATE; }
    cmlba::Ctrl c;
    if (cudaSetDevice(e.device) != cudaSuccess || cudaMemcpy(&c, e.d_ctrl.p, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) {
        e.set_error(std::string("step: ") + cudaGetErrorString(cudaGetLastError())); return CMLBA_ERR_CUDA;
    }
    if (can_break) *can_break = c.canbreak;
    return CMLBA_OK;
}

int cmlba_reset(cmlba_handle *h) { HCHK; return h->eng.reset(); }
static const char *const kStatNames[19] = {"P Energy ( All residuals )", "R Energy", "L Energy ( Linearized )", "M Energy ( Marginalized )", "Total Energy", "X Norm",
                                           " Hessian P Norm", " Hessian L Norm", " Hessian M Norm", " Hessian SC Norm", "P B Norm", "L B Norm", "M B Norm", "SC B Norm",
                                           "OOB", "In", "InIn", "Nores", "Num Linearized"};
int cmlba_get_statistics(const cmlba_handle *h, double *values) {
    if (!h || !values) return CMLBA_ERR_ARG;
    for (int i = 0; i < 19; i++) values[i] = h->eng.stats_[i];
    return CMLBA_OK;
}
const char *cmlba_statistic_name(int i) { return (i >= 0 && i < 19) ? kStatNames[i] : nullptr; }
int cmlba_bench_pass(cmlba_handle *h, int steps, int warmup, int flush_l2, cmlba_bench_result *out) { HCHK; return h->eng.bench_pass(steps, warmup, flush_l2, out); }

int cmlba_read(cmlba_handle *h, const char *name, void *dst, size_t cap, size_t *bytes) { HCHK; if (!name) return CMLBA_ERR_ARG; return h->eng.read(name, dst, cap, bytes); }

int cmlba_nccl_unique_id(void *uid) {
    std::string err;
    if (!uid) return CMLBA_ERR_ARG;
    if (!cmlba::g_nccl.load(err)) { g_create_error = err; return CMLBA_ERR_UNSUPPORTED; }
    return cmlba::g_nccl.GetUniqueId(uid) == 0 ? CMLBA_OK : CMLBA_ERR_CUDA;
}

int cmlba_comm_ipc_handle(cmlba_handle *h, void *handle_64) { HCHK; if (!handle_64) return CMLBA_ERR_ARG; return h->eng.comm_ipc_handle(handle_64); }
int cmlba_comm_ipc_open(cmlba_handle *h, const void *handles) {
    HCHK;
    if (!handles) { h->eng.p2p_ready = false; h->eng.dirty = true; h->eng.prepared = false; return CMLBA_OK; }   // NULL: back to ncclAllReduce
    return h->eng.comm_ipc_open(handles);
}

int cmlba_comm_init(cmlba_handle *h, const void *uid, int rank, int world) {
    HCHK; Engine &e = h->eng;
    if (!uid || world < 1 || rank < 0 || rank >= world) { e.set_error("bad communicator arguments"); return CMLBA_ERR_ARG; }
    if (world > cmlba::MAXF) { e.set_error("at most 16 ranks: the peer-exchange header and the flag arrays hold 16 entries"); return CMLBA_ERR_ARG; }
    if (world == 1) { e.rank = 0; e.world = 1; return CMLBA_OK; }
    if (!cmlba::g_nccl.load(e.err)) return CMLBA_ERR_UNSUPPORTED;
    cudaSetDevice(e.device);
    cmlba::NcclUid id;
    memcpy(id.b, uid, 128);
    if (cmlba::g_nccl.CommInitRank(&e.comm, world, id, rank) != 0) { e.set_error("ncclCommInitRank failed"); return CMLBA_ERR_CUDA; }
    e.rank = rank; e.world = world;
    return CMLBA_OK;
}

}  // extern "C"
