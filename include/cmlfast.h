/* cmlfast.h -- C ABI of the B200 FAST-9 corner detector in libcmlba.so (NEXT #4: the first unit of the ORB extractor).
 *
 * Drop-in boundary for CML::Features::FAST::compute(frame, corners, threshold) (reference: src/cml/features/corner/FAST.{h,cpp}):
 * fast9_detect (9 contiguous circle pixels brighter than p + b or darker than p - b), fast9_score (bisection: the largest b that still
 * detects) and nonmax_suppression (a corner survives iff no 8-neighbour corner scores >= it), corners in raster order.
 * The rest of ORB (per-cell thresholds, octree distribution, orientation, blur, rBRIEF) is not built yet.  There is NO CPU fallback.
 */
#ifndef CMLFAST_H
#define CMLFAST_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmlfast_handle_t *cmlfast_handle;

enum { CMLFAST_OK = 0, CMLFAST_ERR_ARG = -1, CMLFAST_ERR_CUDA = -2 };

int cmlfast_create(int device, int max_width, int max_height, cmlfast_handle *out);
void cmlfast_destroy(cmlfast_handle h);
const char *cmlfast_last_error(cmlfast_handle h);   /* h may be NULL: error of the last failed cmlfast_create */

/* image [height][width] uint8 (host).  xy [capacity][2] and scores [capacity] receive the surviving corners in raster order, *count their number
 * (may exceed capacity: then only `capacity` entries are written).  gpu_ms may be NULL. */
int cmlfast_compute(cmlfast_handle h, const uint8_t *image, int width, int height, int threshold, int capacity, int32_t *xy, int32_t *scores, int32_t *count, float *gpu_ms);

#ifdef __cplusplus
}
#endif
#endif
