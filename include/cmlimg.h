/* cmlimg.h -- C ABI of the B200 image preparation in libcmlba.so (NEXT #3).
 *
 * Drop-in boundary for what CML::CaptureImageGenerator::generate does per frame (reference: src/cml/capture/CaptureImage.cpp:108-262):
 * photometric response LUT (image/LookupTable.h:99-104) -> inverse vignette -> geometric undistortion through the calibration's map
 * (map/InternalCalibration.h:404-437) -> gray pyramid (2x2 means) -> derivative images (image/Array2D.h:288-331) -> weighted gradient norm
 * (image/Array2DProxy.h:198-226).  Two kernel launches per frame; everything stays on the device as `float4 (I, dx, dy, weighted |grad|^2)`
 * texels plus the fp32 gray levels, which is the layout the tracker and the tracer sample.  There is NO CPU fallback.
 * Out of scope: computing the undistortion map from a distortion model (once per session, InternalCalibration.cpp:371-399) -- it is an input.
 */
#ifndef CMLIMG_H
#define CMLIMG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMLIMG_MAX_LEVELS 8

typedef struct cmlimg_handle_t *cmlimg_handle;

enum { CMLIMG_OK = 0, CMLIMG_ERR_ARG = -1, CMLIMG_ERR_CUDA = -2, CMLIMG_ERR_STATE = -3 };

/* in = sensor image size, out = rectified size; levels <= 0: the reference's rule (CaptureImage.cpp:39-72). */
int cmlimg_create(int device, int in_width, int in_height, int out_width, int out_height, int levels, cmlimg_handle *out);
void cmlimg_destroy(cmlimg_handle h);
const char *cmlimg_last_error(cmlimg_handle h);   /* h may be NULL: error of the last failed cmlimg_create */

/* CaptureImageMaker::setLut / setInverseVignette.  lut = 256 response values (NULL: identity, the generator's mDefaultLookupTable);
 * inv_vignette [in_height][in_width] (NULL: none). */
int cmlimg_set_photometric(cmlimg_handle h, const float *lut, const float *inv_vignette);
/* InternalCalibration::mUndistortMap [out_height][out_width][2] = source position of every rectified pixel, NaN = outside (pixel becomes 0).
 * NULL: no pre-undistorter (needs in size == out size). */
int cmlimg_set_undistort_map(cmlimg_handle h, const float *map);

/* Page-locked staging image [in_height][in_width] owned by the handle; a producer writing into it and passing it to cmlimg_prepare skips a host copy. */
float *cmlimg_input_buffer(cmlimg_handle h);

/* generate(): raw gray image [in_height][in_width] float (0..255) -> every level on the device.  gpu_ms (may be NULL): device time of the two kernels. */
int cmlimg_prepare(cmlimg_handle h, const float *raw, float *gpu_ms);

/* Same for an 8-bit sensor image [in_height][in_width] (what the reference's readers deliver before the LUT): a quarter of the upload. */
int cmlimg_prepare_u8(cmlimg_handle h, const uint8_t *raw, float *gpu_ms);

/* levels and their sizes: wh = int32[levels][2]. */
int cmlimg_levels(cmlimg_handle h, int32_t *num_levels, int32_t *wh);

/* Reads a level back: "gray<l>" float [h][w]; "texel<l>" float [h][w][4] = (I, dx, dy, weighted gradient norm).  Returns bytes or a negative error. */
int64_t cmlimg_read(cmlimg_handle h, const char *name, void *dst, int64_t capacity);
/* Device pointer of the same buffers for consumers in this process (valid until the next cmlimg_prepare / destroy). */
const void *cmlimg_device_ptr(cmlimg_handle h, const char *name);

/* Benchmark aid: `repeats` x (L2 flush, then the two kernels on the resident input); mean device ms of the kernels alone. */
int cmlimg_bench(cmlimg_handle h, int repeats, int flush_l2, float *ms_per_frame);

#ifdef __cplusplus
}
#endif
#endif
