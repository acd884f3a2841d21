/* cmltrk.h -- C ABI of the B200 coarse tracker in libcmlba.so (NEXT #1).
 *
 * Drop-in boundary for CML::Optimization::DSOTracker (reference: src/cml/optimization/dso/DSOTracker.h:198-523,
 * DSOTracker.cpp): the coarse-to-fine direct image alignment of a new frame against the inverse-depth map of a
 * reference keyframe.  Each entry point names the reference member it replaces.
 *
 *   cmltrk_make_coarse_depth   DSOTracker::makeCoarseDepthL0(reference, points)        DSOTracker.cpp:494-725
 *   cmltrk_set_frame           CaptureImage pyramid + derivative images of the frame   CaptureImage.cpp:209-262
 *   cmltrk_optimize            DSOTracker::optimize(numTry, frame, reference, camera&, exposure&)  DSOTracker.cpp:15-246
 *                              (K start poses at once: the candidate loop of trackWithMotionModel, DSOTracker.h:240-360)
 *   cmltrk_track               cmltrk_set_frame + cmltrk_optimize in one call (the end-to-end call per frame)
 *
 * Everything runs on the device: the pyramid is built by two kernels, the whole coarse-to-fine Gauss-Newton loop
 * (all levels, all iterations, the 8x8 solves, accept/reject) is ONE kernel launch -- one thread-block cluster per
 * start pose, partial sums exchanged through distributed shared memory.  There is NO CPU fallback.
 * Plain pointers and sizes only; all host pointers may be pageable (they are staged through pinned memory).
 */
#ifndef CMLTRK_H
#define CMLTRK_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMLTRK_MAX_LEVELS 6      /* pyramid levels held; optimize() uses min(levels - 1, 4) as the coarsest (DSOTracker.cpp:24) */
#define CMLTRK_OPT_LEVELS 5
#define CMLTRK_MAX_CANDIDATES 32

typedef struct cmltrk_handle_t *cmltrk_handle;

enum { CMLTRK_OK = 0, CMLTRK_ERR_ARG = -1, CMLTRK_ERR_CUDA = -2, CMLTRK_ERR_STATE = -3 };

/* Parameters (defaults = the reference's createParameter defaults, DSOTracker.h:479-518). */
typedef struct {
    double huber_threshold;            /* "Huber threshold" 9 */
    double cutoff_threshold;           /* "Cutoff threshold" 20 */
    double scale_rotation;             /* "Rotation scale" 1      (applied to increment[0..2]) */
    double scale_translation;          /* "Translation scale" 0.5 (applied to increment[3..5]) */
    double scale_light_a;              /* "Light A scale" 10 */
    double scale_light_b;              /* "Light B scale" 1000 */
    int optimize_a;                    /* "optimizeLightA" true */
    int optimize_b;                    /* "optimizeLightB" true */
    double saturated_ratio_threshold;  /* "saturatedThreshold" 0.33 */
    int levels;                        /* pyramid levels; <= 0: the reference's rule (CaptureImage.cpp:39-72) */
    int cluster_ctas;                  /* CTAs per tracking cluster (1..16); default 16 = the non-portable cluster size, falls back to 8 where it cannot be scheduled */
    int cta_threads;                   /* threads per CTA of the tracking kernel: 256, 384 (default) or 512 */
} cmltrk_config;

/* DSOTracker::Residual (DSOTracker.h:202-236) plus the optimised pose and brightness of one start pose. */
typedef struct {
    double cam[12];                    /* world-to-camera [R row-major | t] of the tracked frame; the start pose if !is_correct */
    double affine[2];                  /* exposure parameters (a, b) as left in `currentExposure` (updated even on failure, DSOTracker.cpp:169) */
    double E[CMLTRK_OPT_LEVELS];
    int32_t num_terms_in_E[CMLTRK_OPT_LEVELS];
    int32_t num_saturated[CMLTRK_OPT_LEVELS];
    int32_t num_robust[CMLTRK_OPT_LEVELS];
    double level_cutoff_repeat[CMLTRK_OPT_LEVELS];
    double flow_vector[3];
    double rel_aff[2];
    double covariance[6];
    int32_t is_correct;
    int32_t too_many_saturated;        /* the reference's (inverted) flag: 1 = saturated ratio is fine (DSOTracker.cpp:238) */
    int32_t iterations;                /* Gauss-Newton steps evaluated (all levels) */
    int32_t levels_used;               /* min(levels, 5) */
    float gpu_ms;                      /* device time of this call's kernels (CUDA events on the tracker stream) */
    int32_t kernel_launches;
} cmltrk_result;

void cmltrk_default_config(cmltrk_config *cfg);
/* Pinhole calibration of level 0 (fx, fy, cx, cy) and the image size; levels follow PinholeUndistorter's pyramid. */
int cmltrk_create(const cmltrk_config *cfg, int device, int width, int height, double fx, double fy, double cx, double cy, cmltrk_handle *out);
void cmltrk_destroy(cmltrk_handle h);
const char *cmltrk_last_error(cmltrk_handle h);   /* h may be NULL: error of the last failed cmltrk_create */

/* makeCoarseDepthL0: projects `num_points` points (host frame index into frame_cams, pixel in the host frame, inverse depth,
 * uncertainty) into the reference keyframe, splats, builds the per-level inverse-depth maps (2x2 sums, dilation, normalisation)
 * and the per-level point lists (u, v, idepth, colour).  ref_gray = level-0 gray image [height][width] float.
 * ref_cam / frame_cams rows = world-to-camera [R(9) | t(3)]; ref_exposure = (exposure time, a, b). */
int cmltrk_make_coarse_depth(cmltrk_handle h, const float *ref_gray, const double ref_cam[12], const double ref_exposure[3], int num_frames,
                             const double *frame_cams, int num_points, const int32_t *pt_frame, const float *pt_xy, const double *pt_idepth,
                             const double *pt_uncertainty);

/* Page-locked staging image [height][width] float owned by the handle.  A producer (camera driver, undistorter) that writes the gray image
 * straight into it and passes this pointer as `gray` to cmltrk_set_frame / cmltrk_track / cmltrk_make_coarse_depth skips the host-side copy;
 * any other pointer is copied into it first.  Valid until cmltrk_destroy; do not write while a call is running. */
float *cmltrk_frame_buffer(cmltrk_handle h);

/* Uploads the frame to track (level-0 gray) and builds its gray pyramid and derivative images on the device. */
int cmltrk_set_frame(cmltrk_handle h, const float *gray, double exposure_time);

/* Device-resident variants (same process, same device; e.g. the levels of include/cmlimg.h):
 *   cmltrk_make_coarse_depth_device: d_gray_levels[l] = device pointer to the fp32 gray image of level l of the reference keyframe (copied);
 *   cmltrk_set_frame_device: d_texel_levels[l] = device pointer to the float4 (I, dx, dy, *) texels of level l of the frame to track; they are
 *   sampled IN PLACE by the following cmltrk_optimize calls and must stay valid and unchanged until then. */
int cmltrk_make_coarse_depth_device(cmltrk_handle h, int levels, const float *const *d_gray_levels, const double ref_cam[12], const double ref_exposure[3], int num_frames,
                                    const double *frame_cams, int num_points, const int32_t *pt_frame, const float *pt_xy, const double *pt_idepth,
                                    const double *pt_uncertainty);
int cmltrk_set_frame_device(cmltrk_handle h, int levels, const void *const *d_texel_levels, double exposure_time);

/* optimize() for `num_candidates` start poses (world-to-camera, [K][12]) and start brightness ([K][2]) at once.
 * last_rmse: NULL, or [CMLTRK_OPT_LEVELS] = mLastResidual.rmse(level) for the rmse sanity check (DSOTracker.cpp:190-196).
 * results [K]. */
int cmltrk_optimize(cmltrk_handle h, int num_candidates, const double *start_cams, const double *start_affine, const double *last_rmse,
                    cmltrk_result *results);

/* set_frame + optimize: one call per tracked frame. */
int cmltrk_track(cmltrk_handle h, const float *gray, double exposure_time, int num_candidates, const double *start_cams, const double *start_affine,
                 const double *last_rmse, cmltrk_result *results);

/* DSOTracker::trackWithMotionModel (DSOTracker.h:240-360): the candidate loop over the poses of Map::multiConstantVelocityMotionModel on the
 * frame given to cmltrk_set_frame / cmltrk_set_frame_device.  The poses are optimised one after the other because every run is gated by the best
 * residual so far (mLastResidual = trackingResult, DSOTracker.cpp:190-196) and the loop stops at the first candidate that is good enough
 * (achievedRes < 1.5 * mLastCoarseRMSE; at most 51 candidates once one succeeded).  failure_mode 1 = when every candidate failed, optimise the
 * first pose again and return that (the reference's "return the best effort" branch).
 * *ok = 1: `best` holds the residual / pose / brightness the reference would leave in the frame (frame->setCamera, setExposureParameters);
 * *ok = 0: tracking failed, `best` is the last kept residual (is_correct = 0 when there was none).  *tried = candidates optimised.
 * The handle keeps mLastCoarseRMSE / mFirstRMSE between calls like the class does (cmltrk_reset_motion_model clears them). */
int cmltrk_track_with_motion_model(cmltrk_handle h, int num_cameras, const double *cameras /* [K][12] */, const double initial_affine[2], int failure_mode,
                                   int *ok, cmltrk_result *best, int *tried);
int cmltrk_reset_motion_model(cmltrk_handle h);
int cmltrk_motion_model_state(cmltrk_handle h, double *last_coarse_rmse, double *first_rmse);

/* Debug / test reads: "pc_n" (int32[levels]), "pc<l>" (float [n][4]), "grad<l>" (float [h][w][4] = I, dx, dy, 0 of the frame to track),
 * "levels_wh" (int32 [levels][2]), "K" (double [levels][4]), "cycles" (int64 [4]: evaluations and SM cycles spent advancing the optimiser,
 * evaluating points and reducing/exchanging sums in the last optimize, start pose 0).  Returns bytes written or a negative error. */
int64_t cmltrk_read(cmltrk_handle h, const char *name, void *dst, int64_t capacity);

/* Device-resident repeat of the last cmltrk_optimize (same start poses) for benchmarks: `repeats` launches, mean device ms per launch. */
int cmltrk_bench_optimize(cmltrk_handle h, int repeats, float *ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif
