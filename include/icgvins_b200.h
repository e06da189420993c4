/*
 * icgvins_b200.h -- C ABI of libicgvins_b200.so: B200-native (sm_100a) replacements for the two compute hot paths
 * of i2Nav-WHU/IC-GVINS.  Plain pointers and sizes only; no C++ / torch / OpenCV / Ceres types cross this boundary.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference checkout,
 * IG/ = ic_gvins/ic_gvins/).  The reference has no FFI of its own: the seams are direct library calls into
 * OpenCV (front end) and Ceres (window solve); INTEGRATION.md shows the shim a maintainer adds at each call site.
 *
 * Conventions
 *   - All functions return 0 on success, a negative ICG_E* code on failure (no exceptions, no aborts).
 *   - "host" pointers are ordinary process memory; "dev" pointers are CUDA device memory of the handle's device.
 *   - A handle is bound to one CUDA device and one stream and is NOT re-entrant (the reference calls each seam
 *     from a single thread: tracking thread IG/ic_gvins.cc:535, optimization thread IG/ic_gvins.cc:434-448).
 *   - Points are interleaved float (x, y) pairs == std::vector<cv::Point2f>::data().
 *   - There is NO CPU fallback: without a CUDA device every create() fails with ICG_ENODEVICE.
 */
#ifndef ICGVINS_B200_H
#define ICGVINS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICG_OK 0
#define ICG_EINVAL (-1)     /* bad argument */
#define ICG_ENODEVICE (-2)  /* no usable CUDA device / driver */
#define ICG_ECUDA (-3)      /* CUDA runtime error (see icg_last_error) */
#define ICG_EUNSUPPORTED (-4) /* parameter combination the sm_100a kernels are not built for */
#define ICG_ENOMEM (-5)
#define ICG_ENCCL (-6)

#define ICG_OPTFLOW_USE_INITIAL_FLOW 4 /* == cv::OPTFLOW_USE_INITIAL_FLOW */

const char *icg_last_error(void);
int icg_version(void);
/* number of kernels launched by this library in this process since load / last reset (bench "gpu_launches") */
uint64_t icg_launch_count(void);
void icg_launch_count_reset(void);

/* ===================================================================================================== *
 *  Path A: pyramidal KLT front end
 * ===================================================================================================== */
typedef struct icg_klt icg_klt;

/*
 * Create a tracker for width x height u8 images.  n_slots device-resident image slots (each holds a 4-level
 * pyramid, levels 0..3 as cv::buildOpticalFlowPyramid(img, winSize 21, maxLevel 3) produces them);
 * max_points = largest n of one call.  stream may be NULL (handle creates its own) or a cudaStream_t.
 */
int icg_klt_create(icg_klt **h, int width, int height, int n_slots, int max_points, int device, void *stream);
void icg_klt_destroy(icg_klt *h);

/*
 * Drop-in for cv::calcOpticalFlowPyrLK as the reference calls it (IG/tracking/tracking.cc:385,390,487,493):
 *   cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), max_level,
 *                            TermCriteria(COUNT+EPS, max_iter, eps), flags)
 * Host buffers, synchronous.  next_xy is in/out (initial flow when flags has ICG_OPTFLOW_USE_INITIAL_FLOW).
 * err may be NULL (the reference never reads it).  Only win == 21 and max_level <= 3 are built
 * (the values at all four call sites); anything else returns ICG_EUNSUPPORTED.
 * Pyramids are cached per image content (the reference passes the same two images to 4 calls per frame).
 */
int icg_klt_calc_optical_flow_pyr_lk(icg_klt *h, const uint8_t *prev, const uint8_t *next, int stride,
                                     const float *prev_xy, float *next_xy, uint8_t *status, float *err, int n,
                                     int win, int max_level, int max_iter, double eps, int flags);

/*
 * Fused replacement for the forward LK + backward LK + gate block (IG/tracking/tracking.cc:385-403 and :487-506):
 *   status[k] = st_fwd && st_bwd && !isOnBorder(fwd) && ptsDistance(bwd, prev) < 0.5
 * next_xy in: predicted positions, out: forward result.  back_xy (may be NULL) receives the backward result.
 */
int icg_klt_track_fb(icg_klt *h, const uint8_t *prev, const uint8_t *next, int stride, const float *prev_xy,
                     float *next_xy, float *back_xy, uint8_t *status, int n);

/* ---- device-resident / batched API (throughput mode: many independent streams per GPU) ---- */

/* async H2D of one frame into slot's level 0 (host memory should be pinned for overlap) + pyramid build */
int icg_klt_upload(icg_klt *h, int slot, const uint8_t *host_img, int stride);
/* async H2D of one frame into slot's level 0 only (no pyramid build; pair with icg_klt_build_pyramids) */
int icg_klt_upload_level0(icg_klt *h, int slot, const uint8_t *host_img, int stride);
/* synchronous D2H of one pyramid level of a slot into a host buffer of row stride `stride` (parity tests) */
int icg_klt_download_level(icg_klt *h, int slot, int level, uint8_t *host_img, int stride);
/* device pointer + pitch of a slot's level-0 plane, so a producer on the same device can write frames in place */
int icg_klt_slot_level0(icg_klt *h, int slot, void **dev_ptr, int *pitch);
/* device pointer + pitch + size of level `level` (read-back for parity tests) */
int icg_klt_slot_level(icg_klt *h, int slot, int level, void **dev_ptr, int *pitch, int *w, int *hgt);
/* build levels 1..3 for slots [first, first+count) from their level-0 planes (one launch) */
int icg_klt_build_pyramids(icg_klt *h, int first_slot, int count);
/*
 * Batched forward+backward tracking of n_total points, all arrays in DEVICE memory:
 *   slots[2k], slots[2k+1] = (prev slot, next slot) of point k;  prev_xy/init_xy/fwd_xy/bwd_xy: float2 per point.
 * Asynchronous on the handle's stream.  mode 0: forward only (status = raw LK status);
 * mode 1: fused forward+backward+gates (as icg_klt_track_fb).
 */
int icg_klt_track_batch_dev(icg_klt *h, int n_total, const int32_t *dev_slots, const float *dev_prev_xy,
                            const float *dev_init_xy, float *dev_fwd_xy, float *dev_bwd_xy, uint8_t *dev_status,
                            int mode);
int icg_klt_sync(icg_klt *h);

#ifdef __cplusplus
}
#endif
#endif /* ICGVINS_B200_H */
