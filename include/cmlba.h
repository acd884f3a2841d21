/* cmlba.h -- C ABI of libcmlba.so: B200-native photometric bundle adjustment (DSO-style sliding window).
 *
 * Drop-in boundary for ONE path of lizabelos/libCML: CML::Optimization::DSOBundleAdjustment
 * (reference: src/cml/optimization/dso/DSOBundleAdjustment.{h,cpp}, "BA.h"/"BA" below).  The reference
 * exposes a C++ class, not a C ABI; every entry point below cites the class
 * member it replaces.  A handle mirrors one DSOBundleAdjustment instance: it owns the window
 * bookkeeping (frames, points, residuals, FEJ evaluation points, energy thresholds) on the host and all
 * device memory on one GPU.  The caller (CML's Frame/MapPoint graph via the adapter shown in
 * INTEGRATION.md) stays authoritative: it pushes cameras in at run() and reads poses / affine /
 * inverse depths / outliers back out.
 *
 * Conventions: poses are world->camera, R row-major 3x3 + t (double[12] = R(9) | t(3)), as
 * CML::Camera (map/Camera.h:289-315).  All pointers are HOST pointers valid for the duration of the
 * call only.  Every function returns 0 on success or a negative cmlba_status; cmlba_last_error()
 * gives the message.  Nothing aborts.  A handle is single-threaded (BA is not re-entrant: SURVEY 8b).
 * There is NO CPU fallback: without a CUDA device cmlba_create fails with CMLBA_ERR_CUDA.
 */
#ifndef CMLBA_H
#define CMLBA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMLBA_MAX_FRAMES 16 /* window size limit (reference default maxFrames = 6, BA.h:271) */
#define CMLBA_PATTERN 8     /* PredefinedPattern::star8, types.h:1395-1407 */

typedef struct cmlba_handle cmlba_handle;

typedef enum cmlba_status {
    CMLBA_OK = 0,
    CMLBA_ERR_ARG = -1,      /* bad argument (null pointer, unknown id, window too large ...) */
    CMLBA_ERR_CUDA = -2,     /* CUDA runtime error / no device */
    CMLBA_ERR_STATE = -3,    /* call order violated (e.g. run() before set_calib) */
    CMLBA_ERR_NUMERIC = -4,  /* non-finite energy or step: the reference's run() returns false (BA:836-841, 898-905, 1489-1492) */
    CMLBA_ERR_UNSUPPORTED = -5
} cmlba_status;

/* residual states, DSOResidual.h:14-16 */
enum { CMLBA_RES_IN = 0, CMLBA_RES_OOB = 1, CMLBA_RES_OUTLIER = 2 };

/* BA parameters (BA.h:235-288, same meaning and defaults; YAML key dsoBa.<name>) */
typedef struct cmlba_config {
    int iterations;             /* "iterations" 4 */
    float huber_threshold;      /* "Huber threshold" 9 */
    float outlier_th_sum;       /* "outlierTHSumComponent" 2500 */
    float th_opt_iterations;    /* "ThOptIterations" 1.2 */
    float scale_rotation;       /* "Rotation scale" 1 */
    float scale_translation;    /* "translationScale" 0.5 */
    float scale_light_a;        /* "Light A scale" 10 */
    float scale_light_b;        /* "Light B scale" 1000 */
    float scale_f;              /* "Scale F" 50 */
    float scale_c;              /* "Scale C" 50 */
    int force_accept;           /* "forceAccept" true */
    int fix_lambda;             /* "fixLambda" true */
    float fixed_lambda;         /* "fixedLambda" 1e-5 */
    int idepth_fix_prior;       /* "iDepth Fix Prior" 2500 */
    float solver_mode_delta;    /* "Solver mode delta" 1e-5 */
    int optimize_light_a;       /* "optimizeLightA" true */
    int optimize_light_b;       /* "optimizeLightB" true */
    int disable_marginalization;/* "disableMarginalization" true */
    int max_frames;             /* "maxFrames" 6: flagFramesForMarginalization keeps at most this many frames (BA:649) */
    int frame_min_age;          /* "frameMinAge" 1 (BA:658-680) */
    float min_idepth_h_marg;    /* "Minimum iDepth Hessian Marginlaization" 50 (BA:2318) */
    int async_image_upload;     /* not a reference parameter.  0 (default): cmlba_add_frame returns after the image has left the caller's
                                 * buffer.  1: the host->device copy is only enqueued; the caller keeps `grad` valid and unmodified until
                                 * the next cmlba_add_points / cmlba_prepare / cmlba_run returns (CML pins a keyframe's GradientImage for as
                                 * long as the frame is active, capture/CaptureImage.cpp:269-362), so uploads overlap the host bookkeeping.
                                 * Needs page-locked memory to have an effect. */
} cmlba_config;

/* Fills *cfg with the reference defaults. */
int cmlba_default_config(cmlba_config *cfg);

/* new DSOBundleAdjustment(parent) (BA:322-334).  device = CUDA ordinal. */
int cmlba_create(const cmlba_config *cfg, int device, cmlba_handle **out);
int cmlba_destroy(cmlba_handle *h);
const char *cmlba_last_error(const cmlba_handle *h);

/* The level-0 pinhole + size cached by addNewFrame (BA:419-425: mPinhole, mWidth, mHeight). */
int cmlba_set_calib(cmlba_handle *h, double fx, double fy, double cx, double cy, int width, int height);

/* addNewFrame(PFrame, immatureGroup) (BA:417-462): appends a keyframe to the window.
 *   frame_id      caller's id (Frame::getId()); must be larger than every id already in the window
 *   world_to_cam  Frame::getCamera() at insertion -> becomes the FEJ evaluation point (DSOFrame::setEvalPT_scaled)
 *   aff_a,aff_b   Frame::getExposure().getParameters();  exposure_time = CaptureImage::getExposureTime()
 *   grad_aos      getCaptureFrame().getDerivativeImage(0): width*height texels of (I, dI/dx, dI/dy) fp32, row-major
 *   is_init_frame Frame::isGroup(Map::INITFRAME) (-> hasDepthPrior of its points, BA:396)
 * Creates residuals from every point already in the window to this frame (BA:455-460). */
int cmlba_add_frame(cmlba_handle *h, int64_t frame_id, const double world_to_cam[12], double aff_a, double aff_b,
                    double exposure_time, const float *grad_aos, int is_init_frame);

/* Same as cmlba_add_frame, but from the rectified level-0 GRAY image (frame->getCaptureFrame().getGrayImage(0): width*height floats).
 * The derivative image (I, dI/dx, dI/dy) is built on the device exactly as CaptureImageGenerator::generate does
 * (capture/CaptureImage.cpp:249 -> Array2D::gradientImage, image/Array2D.h:288-294, 314-331: central differences * 0.5, zero on the
 * 1-pixel border) -- bit-identical texels, a third of the host->device traffic. */
int cmlba_add_frame_gray(cmlba_handle *h, int64_t frame_id, const double world_to_cam[12], double aff_a, double aff_b,
                         double exposure_time, const float *gray /* [height][width] */, int is_init_frame);
/* Same, from DEVICE memory: d_texels = level-0 float4 (I, dx, dy, *) texels on this handle's device, e.g. cmlimg_device_ptr(img, "texel0")
 * (include/cmlimg.h).  One device-to-device copy, complete when the call returns (the source may be overwritten afterwards). */
int cmlba_add_frame_device(cmlba_handle *h, int64_t frame_id, const double world_to_cam[12], double aff_a, double aff_b, double exposure_time,
                           const void *d_texels, int is_init_frame);

/* addPoints(const PointSet&) (BA:382-415).  n points; host_frame_id[i] = getReferenceFrame()->getId(),
 * xy = getReferenceCorner() (float x,y), idepth = getReferenceInverseDepth().  Reference colours
 * (integer-pixel gray, MapObject.h:398-399) and gradient weights (BA:405-411) are computed on the
 * device from the host frame's image.  One residual per (point, other frame) is created (BA:398-400). */
int cmlba_add_points(cmlba_handle *h, int n, const int64_t *point_id, const int64_t *host_frame_id, const float *xy,
                     const double *idepth);

/* DSOContext::removePoint / frame removal without marginalisation prior (DSOContext.h:94-110, 152-172). */
int cmlba_remove_point(cmlba_handle *h, int64_t point_id);
int cmlba_remove_frame(cmlba_handle *h, int64_t frame_id);

/* ---- window maintenance around run() (Hybrid::directMap, slam/modslam/direct/Mapping.cpp:61-100) ----
 * The decisions (which frames / points leave the window, the per-frame counters they depend on) follow the reference
 * exactly.  The marginalisation PRIOR H_M,b_M that marginalizePointsF / marginalizeFrame fold the leaving variables into
 * (BA:2500-2508, 520-548) is only maintained when disable_marginalization == 0: with the default (true) the reference
 * zeroes it before every solve (BA:1395-1398), so it never reaches a result and is skipped here.
 *
 * cmlba_flag_frames_for_marginalization   flagFramesForMarginalization (BA:603-708).  The reference runs it at the top of
 *      addNewFrame (BA:428), i.e. call it BEFORE cmlba_add_frame of the new keyframe.  cams = current Frame::getCamera()
 *      of the window frames (NULL: last optimised poses); num_immature[i] = frame->getReferenceGroupMapPoints(immature)
 *      .size() (NULL: zeros).  flagged_ids (capacity *n_flagged in, count out) lists every frame flagged so far.
 * cmlba_try_marginalize     tryMarginalize (BA:2240-2363) + isOOB (BA:2515-2554): points to drop go to getOutliers()
 *      and leave the window; points to marginalise are marked (DSOTOMARGINALIZE).
 * cmlba_marginalize_points  marginalizePointsF (BA:2466-2513): marked points leave the window as marginalised
 *      (numMarginalized / numResidualsOut counters of DSOContext.h:99-110, 221-229); ids out.  With a live prior they are
 *      first re-linearised and accumulated in MARGINALIZED mode on the device: H_M += 1/4 (M - M_sc).
 * cmlba_marginalize_frames  marginalizeFrames (BA:710-742) -> marginalizeFrame -> removeFrame: flagged frames leave,
 *      with the points they host and the residuals that target them; ids out.  With a live prior the frame's block is
 *      Schur-complemented out of H_M first (BA:464-548). */
int cmlba_flag_frames_for_marginalization(cmlba_handle *h, const double *cams /*[n][12] or NULL*/, const int32_t *num_immature /*[n] or NULL*/,
                                          int64_t *flagged_ids, int *n_flagged);
int cmlba_try_marginalize(cmlba_handle *h, int *n_dropped, int *n_to_marginalize);
int cmlba_marginalize_points(cmlba_handle *h, int64_t *point_ids, int *n);
int cmlba_marginalize_frames(cmlba_handle *h, int64_t *frame_ids, int *n);

/* bool run(bool updatePointsOnly) (BA:744-910).  cams = the graph's current Frame::getCamera() for every
 * window frame in window order (what updateCamera() reads, BA.h:54-60); NULL = keep current states.
 * iterations <= 0 uses cfg.iterations.  Returns CMLBA_ERR_NUMERIC where the reference returns false. */
typedef struct cmlba_run_result {
    int iterations_done;
    int num_residuals;          /* active residuals linearised per pass */
    int num_dropped;            /* residuals deleted by the final linearizeAll(true) (BA:1623-1640) */
    int num_outliers;           /* points left without residual -> getOutliers() */
    double energy_first;        /* lastEnergy[0] before iteration 0 */
    double energy_last;         /* energy of the final linearizeAll(true) */
    double gpu_ms;              /* device time of the whole run (CUDA events) */
    int kernel_launches;        /* kernels launched by this run */
    int num_rejected;           /* GN steps undone because the energy did not decrease (forceAccept = false, BA:866-876) */
} cmlba_run_result;
int cmlba_run(cmlba_handle *h, const double *cams /* [n_frames][12] or NULL */, int iterations, int update_points_only,
              cmlba_run_result *result);

/* ---- write-back side of the boundary (what run() stores into the graph, BA:934-940, 966-982) ---- */
int cmlba_num_frames(const cmlba_handle *h);
int cmlba_num_points(const cmlba_handle *h);
int cmlba_num_residuals(const cmlba_handle *h);
/* per frame, window order: Frame::setCamera(PRE_worldToCam), setExposureParameters(aff_g2l), + BA-internal state */
int cmlba_get_frames(const cmlba_handle *h, int64_t *frame_id, double *world_to_cam /*[n][12]*/, double *aff_ab /*[n][2]*/,
                     double *state /*[n][10]*/, double *evalpt /*[n][12]*/, double *energy_th /*[n]*/);
/* per point, insertion order (dropped points removed): setReferenceInverseDepth, setUncertainty (DSOPoint.h:107-118) */
int cmlba_get_points(const cmlba_handle *h, int64_t *point_id, double *idepth, double *uncertainty, float *idepth_hessian,
                     float *max_rel_baseline, int32_t *num_good_residuals, int32_t *good_for_tracking);
/* getOutliers() (BA.h:50): ids of points that lost all residuals in the last run(); returns count via *n (capacity in) */
int cmlba_get_outliers(const cmlba_handle *h, int64_t *point_id, int *n);
/* surviving residuals: (point_id, target frame_id, state, energy) */
int cmlba_get_residuals(const cmlba_handle *h, int64_t *point_id, int64_t *target_frame_id, int32_t *state, double *energy);

/* The 19 Statistic series the class declares (BA.h:215-233), latest value of each, in declaration order; cmlba_statistic_name(i) returns
 * the reference's own name of series i ("P Energy ( All residuals )", ..., "Num Linearized").  Energies and norms come from device
 * scalars of the last GN iteration (BA:798-802, 847-851, 1415-1425), OOB / In / InIn / Nores from the last cmlba_try_marginalize
 * (BA:2356-2359).  The four "B Norm" series are declared but never fed by the reference; here they hold |bA|, |bL|, |bM|, |b_sc|. */
#define CMLBA_NUM_STATISTICS 19
int cmlba_get_statistics(const cmlba_handle *h, double *values /* [CMLBA_NUM_STATISTICS] */);
const char *cmlba_statistic_name(int i);

/* ---- stage entry points (protected members of the class; used by parity tests and profiling) ----
 * cmlba_prepare      run() prologue BA:753-782: updateCamera, collect active residuals, computeAdjoints, computeDelta
 * cmlba_linearize    linearizeAll(fixLinearization) BA:1497-1646 (+ setNewFrameEnergyTH); *energy = returned [0]
 * cmlba_apply        applyActiveRes(true) BA:2045-2049
 * cmlba_solve        backupState + solveSystem(iteration, lambda) BA:912-926, 1339-1495 (accumulate, Schur, stitch, LM solve, back-substitute)
 * cmlba_step         doStepFromBackup(updatePointsOnly) BA:948-1028; *can_break = returned bool
 */
int cmlba_prepare(cmlba_handle *h, const double *cams);
int cmlba_linearize(cmlba_handle *h, int fix_linearization, double *energy);
int cmlba_apply(cmlba_handle *h);
int cmlba_solve(cmlba_handle *h, int iteration);
int cmlba_step(cmlba_handle *h, int update_points_only, int *can_break);

/* Empties the window (all frames, points, residuals) but keeps the handle and its device allocations. */
int cmlba_reset(cmlba_handle *h);

/* Benchmark helper (bench.py): `steps` device-timed passes of the Jacobian + Schur accumulation hot path
 * (linearize -> accumulate -> Schur -> stitch) over the prepared window; CUDA events on the launching stream.
 * flush_l2 != 0 evicts the window between passes by writing a buffer larger than L2 outside the timed region. */
typedef struct cmlba_bench_result {
    int steps, residuals, points, frames, launches_per_pass;
    double ms_pass;         /* mean duration of one whole pass */
    double ms_linearize, ms_accumulate, ms_schur, ms_stitch;   /* mean event-to-event intervals of a separate loop in which the kernels of the pass
                                                                * run serialised on one stream: linearize_tile_kernel, accumulate_kernel, schur_kernel,
                                                                * stitch_pair_kernel */
    double ms_assemble;        /* ... assemble_kernel */
    double ms_event_overhead;  /* ... an EMPTY interval: the event-record overhead every interval above carries */
} cmlba_bench_result;
int cmlba_bench_pass(cmlba_handle *h, int steps, int warmup, int flush_l2, cmlba_bench_result *out);

/* Named read-back of internal buffers for parity tests (device -> host copy, residual arrays in the
 * order of cmlba_read("res_point")/("res_target")).  Returns the number of BYTES the buffer holds in
 * *bytes; copies min(capacity, bytes) into dst (dst may be NULL to query the size).  Names are listed in
 * DESIGN.md ("debug buffers"). */
int cmlba_read(cmlba_handle *h, const char *name, void *dst, size_t capacity, size_t *bytes);

/* ---- multi-GPU (points sharded, images/frames replicated) ----
 * Every rank holds the same frames and its own shard of the points.  The reduced system
 * [H_A | b_A | H_sc | b_sc | energies | quantile histogram] is summed across ranks once per GN
 * iteration with ncclAllReduce (libnccl is dlopen'ed on first use).
 * unique_id: 128-byte ncclUniqueId created on rank 0 with cmlba_nccl_unique_id and broadcast by the
 * caller's own plumbing (e.g. torch.distributed). */
int cmlba_nccl_unique_id(void *unique_id_128);
int cmlba_comm_init(cmlba_handle *h, const void *unique_id_128, int rank, int world_size);
/* Optional, after cmlba_comm_init (one process per GPU on one NVLink/NVSwitch node): exchange of the reduced system through
 * cudaIpc-mapped peer memory instead of ncclAllReduce.  Every rank calls cmlba_comm_ipc_handle (64-byte cudaIpcMemHandle_t of
 * its exchange buffer), the caller all-gathers the handles in rank order, every rank calls cmlba_comm_ipc_open with all of
 * them.  From then on assemble_kernel writes the rank's partial system into its buffer and signals the peers, and
 * solve_kernel sums the ranks' buffers with NVLink loads in its prologue (one kernel for the collective + the solve). */
int cmlba_comm_ipc_handle(cmlba_handle *h, void *handle_64);
int cmlba_comm_ipc_open(cmlba_handle *h, const void *handles /* [world_size][64]; NULL switches back to NCCL (every rank must) */);

/* library / build info: "libcmlba <version> sm_100a" */
const char *cmlba_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CMLBA_H */
