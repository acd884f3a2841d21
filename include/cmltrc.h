/* cmltrc.h -- C ABI of the B200 immature-point tracer in libcmlba.so (NEXT #2).
 *
 * Drop-in boundary for CML::Optimization::DSOTracer (reference: src/cml/optimization/dso/DSOTracer.h:34-206, DSOTracer.cpp):
 * creation of immature points, their epipolar-line search in every new frame, and the 1-dof Gauss-Newton that activates them.
 *
 *   cmltrc_add_frame / cmltrc_set_frame_pose / cmltrc_remove_frame   the frame group the tracer works on (Map::getGroupFrames(frameGroup))
 *   cmltrc_make_new_traces       DSOTracer::makeNewTracesFrom(frame, group)                  DSOTracer.cpp:538-583
 *   cmltrc_trace_new_coarse      DSOTracer::traceNewCoarse(frameToTrace, frameGroup) -> trace  DSOTracer.cpp:17-60, 585-832
 *   cmltrc_optimize_immature     DSOTracer::optimizeImmaturePoint(point, minObs, frameGroup) DSOTracer.cpp:280-411 (+ linearizeResidual :413-494)
 *   cmltrc_activate_points       DSOTracer::activatePoints(frameGroup, pointGroup)            DSOTracer.cpp:62-278 (DistanceMap: utils/DistanceMap.h)
 *   cmltrc_get_points            DSOTracerPointPrivate fields (DSOTracer.h:14-32)
 *
 * Points, their state and the frames' images live on the device between calls; one warp per point.  There is NO CPU fallback.
 * The pixel selection of makeNewTraces is include/cmlsel.h; the Map bookkeeping around activatePoints (setMapPoint, removeMapPoint) stays with the caller.
 */
#ifndef CMLTRC_H
#define CMLTRC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMLTRC_MAX_FRAMES 16

typedef struct cmltrc_handle_t *cmltrc_handle;

enum { CMLTRC_OK = 0, CMLTRC_ERR_ARG = -1, CMLTRC_ERR_CUDA = -2, CMLTRC_ERR_STATE = -3 };

/* DSOTracerStatus (DSOPoint.h:12-19) */
enum { CMLTRC_IPS_GOOD = 0, CMLTRC_IPS_OOB, CMLTRC_IPS_OUTLIER, CMLTRC_IPS_SKIPPED, CMLTRC_IPS_BADCONDITION, CMLTRC_IPS_UNINITIALIZED };

/* Parameters (defaults = the reference's createParameter defaults, DSOTracer.h:186-203). */
typedef struct {
    float min_idepth_h_act;        /* "Min iDepth H Act" 100 */
    int gn_iterations;             /* "Iteration on point activation" 3 */
    float huber_threshold;         /* "Huber Threshold" 9 */
    float outlier_th;              /* "Outlier threshold" 144 */
    float outlier_th_sum_component;/* "outlierTHSumComponent" 2500 */
    float max_pix_search;          /* "Max pixel search" 0.027 */
    float max_slack_interval;      /* "Max slack interval" 1.5 */
    float trace_step_size;         /* "Intial Step size" 1 */
    float min_improvement_factor;  /* "Minimum improvement factor" 2 */
    float min_trace_test_radius;   /* "Trace test radius" 2 */
    float extra_slack_on_th;       /* "ExtraSlackOnTH" 1.2 */
} cmltrc_config;

/* State of one immature point (DSOTracerPointPrivate). */
typedef struct {
    int32_t status;                /* lastTraceStatus */
    int32_t host_frame_slot;       /* -1 = removed */
    double idepth_min, idepth_max; /* iDepthMin, iDepthMax (NaN = not yet bounded) */
    double last_trace_uv[2];
    double last_trace_pixel_interval;
    double quality;
    double grad_h[4];
    double energy_th;
} cmltrc_point;

/* Result of optimizeImmaturePoint for one point. */
typedef struct {
    int32_t rc;                    /* the reference's return value: 1 activated, 0 not enough constraint (Hdd gate), -1 drop the point */
    float idepth;                  /* setReferenceInverseDepth value when rc == 1 */
    uint32_t in_mask;              /* bit t set: the residual towards the t-th frame of the window (newest first, host skipped) ended IN
                                      (the frames that receive addDirectApparitions) */
} cmltrc_activation;

void cmltrc_default_config(cmltrc_config *cfg);
int cmltrc_create(const cmltrc_config *cfg, int device, int width, int height, double fx, double fy, double cx, double cy, cmltrc_handle *out);
void cmltrc_destroy(cmltrc_handle h);
const char *cmltrc_last_error(cmltrc_handle h);   /* h may be NULL: error of the last failed cmltrc_create */

/* Frame group.  gray = level-0 gray image [height][width]; cam = world-to-camera [R(9) | t(3)]; exposure = (time, a, b). */
int cmltrc_add_frame(cmltrc_handle h, int64_t frame_id, const float *gray, const double cam[12], const double exposure[3]);
/* Same from DEVICE memory of this handle's device: level-0 fp32 gray and float4 (I, dx, dy, *) texels (e.g. cmlimg_device_ptr "gray0" / "texel0"); copied. */
int cmltrc_add_frame_device(cmltrc_handle h, int64_t frame_id, const float *d_gray, const void *d_texels, const double cam[12], const double exposure[3]);
int cmltrc_set_frame_pose(cmltrc_handle h, int64_t frame_id, const double cam[12], const double exposure[3]);
int cmltrc_remove_frame(cmltrc_handle h, int64_t frame_id);   /* also removes the immature points hosted in it */

/* makeNewTracesFrom: `count` new immature points hosted in frame_id at pixel xy [count][2]; *first_id receives the id of the first one
 * (ids are consecutive).  */
int cmltrc_make_new_traces(cmltrc_handle h, int64_t frame_id, int count, const float *xy, int64_t *first_id);
int cmltrc_remove_points(cmltrc_handle h, int count, const int64_t *ids);
int64_t cmltrc_num_points(cmltrc_handle h);       /* ids issued so far (removed ones included) */

/* traceNewCoarse: traces every live immature point not hosted in frame_id into it.  status_histogram: NULL or int32[6]
 * (trace_good, trace_oob, trace_out, trace_skip, trace_badcondition, trace_uninitialized of DSOTracer.cpp:19-48). */
int cmltrc_trace_new_coarse(cmltrc_handle h, int64_t frame_id, int32_t *status_histogram, float *gpu_ms);

/* optimizeImmaturePoint for `count` points against all frames of the group. */
int cmltrc_optimize_immature(cmltrc_handle h, int count, const int64_t *ids, int min_obs, cmltrc_activation *results, float *gpu_ms);

int cmltrc_get_points(cmltrc_handle h, int64_t first_id, int count, cmltrc_point *out);

/* Counters of one activatePoints call (the statistics the reference publishes, DSOTracer.cpp:106-115, 196-203, 226-262). */
typedef struct {
    double current_minimum_distance;   /* mCurrentMinimumDistance after the adaptation */
    int32_t urgently_need_new_points;  /* mUrgentlyNeedNewPoints */
    int32_t num_deleted_outlier, num_deleted_oob, num_skipped_status, num_skipped_pixel_interval, num_skipped_quality, num_skipped_depth;
    int32_t num_to_optimize, num_mapped, num_non_mapped, num_dropped;
} cmltrc_activate_stats;

/* DSOTracer::activatePoints(frameGroup, pointGroup) DSOTracer.cpp:62-278.  last_frame_id = Map::getLastGroupFrame(frameGroup);
 * active_xy [num_active][2] = the active points of pointGroup projected into that frame (distorted pixels; they seed the distance map and their
 * number drives the minimum-distance adaptation against desired_point_density = "desiredPointDensity"); immature_ids = the immature points in the
 * order the caller's set iterates them (the greedy distance-map gating depends on it), immature_types = my_type of each (NULL: 1).
 * Outputs: activated_ids / activated (rc == 1 entries: inverse depth, frames that receive the apparition) up to `capacity`, removed_ids (points the
 * reference hands to removeMapPoint) up to `capacity`.  Activated and removed points leave the handle's immature set.
 * min_trace_quality = "Min Trace Quality" (3). */
int cmltrc_activate_points(cmltrc_handle h, int64_t last_frame_id, int num_active, const double *active_xy, int desired_point_density, float min_trace_quality,
                           int num_immature, const int64_t *immature_ids, const float *immature_types, int capacity, int64_t *activated_ids,
                           cmltrc_activation *activated, int32_t *num_activated, int64_t *removed_ids, int32_t *num_removed, cmltrc_activate_stats *stats);
/* mCurrentMinimumDistance (starts at 2). */
int cmltrc_set_minimum_distance(cmltrc_handle h, double v);

#ifdef __cplusplus
}
#endif
#endif
