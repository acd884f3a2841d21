/* cmlsel.h -- C ABI of the B200 pixel selector in libcmlba.so (NEXT #4: the PixelSelector part).
 *
 * Drop-in boundary for CML::Features::PixelSelector (reference: src/cml/features/corner/PixelSelector.{h,cpp}), the DSO candidate selector
 * DSOTracer::makeNewTraces runs on every keyframe (DSOTracer.cpp:496-503): per-32x32-block gradient-histogram thresholds, the three-level
 * potential-grid selection with random directions, the re-sampling recursion and the random sub-sampling.
 *
 *   cmlsel_create          PixelSelector::PixelSelector(parent, w, h)    PixelSelector.cpp:10-21 (the LCG random pattern, state 777)
 *   cmlsel_compute         PixelSelector::compute(cp, corners, types, density, recursionsLeft, thFactor)   PixelSelector.cpp:367-384
 *                          -> makeMaps :121-213 -> makeHists :41-118, select :217-365
 *   cmlsel_set_potential   PixelSelector::setPotential
 *
 * Input = the device-resident texel levels 0..2 of include/cmlimg.h (float4 (I, dx, dy, weighted gradient norm)).  The reference's select()
 * is one sequential sweep whose random directions depend on a running counter; here every 4*pot block is simulated exactly by one
 * thread and the counter is a prefix sum that is iterated to its fixed point, so corners and types are IDENTICAL to the reference's.
 * There is NO CPU fallback.
 */
#ifndef CMLSEL_H
#define CMLSEL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmlsel_handle_t *cmlsel_handle;

enum { CMLSEL_OK = 0, CMLSEL_ERR_ARG = -1, CMLSEL_ERR_CUDA = -2, CMLSEL_ERR_STATE = -3 };

int cmlsel_create(int device, int width, int height, cmlsel_handle *out);
void cmlsel_destroy(cmlsel_handle h);
const char *cmlsel_last_error(cmlsel_handle h);   /* h may be NULL: error of the last failed cmlsel_create */

int cmlsel_set_potential(cmlsel_handle h, int potential);
int cmlsel_get_potential(cmlsel_handle h);

/* d_texels[l] = device pointer to the float4 texels of level l = 0, 1, 2 (sizes w x h, w/2 x h/2, w/4 x h/4).
 * corners_xy [capacity][2] and types [capacity] receive the selection in the reference's emission order (x outer, y inner);
 * *count the number selected (may exceed capacity: then only `capacity` entries are written).  gpu_ms may be NULL. */
int cmlsel_compute(cmlsel_handle h, const void *const *d_texels, float density, int recursions_left, float th_factor, int capacity, float *corners_xy, float *types,
                   int32_t *count, float *gpu_ms);

/* Debug / test reads: "ths", "ths_smoothed" (float [(h/32)][(w/32)]), "map" (float [h][w], the selection map of the last compute). */
int64_t cmlsel_read(cmlsel_handle h, const char *name, void *dst, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
