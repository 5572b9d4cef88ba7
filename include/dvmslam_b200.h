/*
 * dvmslam_b200.h -- C-ABI of libdvmslam_b200.so: the B200-native (sm_100a) ORB front end and
 * local back end of DVM-SLAM behind the reference's own operator surface.
 *
 * The reference (proroklab/DVM-SLAM, /root/reference) has no FFI: its hot path sits behind three
 * C++ classes compiled into libORB_SLAM3 -- ORBextractor, ORBmatcher, Optimizer.  Each entry point
 * below names the reference interface it replaces (O3/ = src/slam_system/orb_slam3/).  The C++
 * drop-in adapters that give back the reference's class signatures on top of these calls live in
 * dvmslam_b200/host/ ; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; every function returns DVM_OK (0) or a negative
 * dvm_status; no exceptions cross the boundary; a handle is bound to one GPU and one CUDA stream,
 * is not thread-safe, and different handles are independent.  "host" entry points are synchronous
 * (results are in host memory on return, like the reference's calls); "_device" entry points
 * enqueue on the handle's stream and leave results in HBM for the next stage.
 */
#ifndef DVMSLAM_B200_H
#define DVMSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVM_API __attribute__((visibility("default")))

typedef enum dvm_status {
    DVM_OK = 0,
    DVM_ERR_INVALID = -1,   /* bad argument (null pointer, size out of range, wrong state) */
    DVM_ERR_CUDA = -2,      /* a CUDA runtime call failed; see dvm_last_error() */
    DVM_ERR_CAPACITY = -3,  /* an internal or caller-provided buffer was too small */
    DVM_ERR_NO_DEVICE = -4, /* no usable sm_100 GPU: there is no CPU fallback */
    DVM_ERR_NUMERIC = -5    /* linear solve failed / non-finite values */
} dvm_status;

/* Text of the last error raised on the calling thread ("" if none). */
DVM_API const char* dvm_last_error(void);
/* Library version / build info, and the number of kernels this library has launched so far in the
 * process (used by bench.py for "gpu_launches"). */
DVM_API const char* dvm_version(void);
DVM_API uint64_t dvm_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * ORB extractor   -- replaces ORB_SLAM3::ORBextractor (O3/include/ORBextractor.h:44-96)
 * ---------------------------------------------------------------------------------------------- */

/* Same memory layout as cv::KeyPoint (28 bytes), which is what ORBextractor::operator() fills. */
typedef struct dvm_keypoint {
    float x, y;      /* pt, in level-0 pixel units (already multiplied by the level's scale) */
    float size;      /* (int)(31 * scale[octave]) */
    float angle;     /* degrees in [0,360), cv::fastAtan2 of the intensity centroid */
    float response;  /* cv::FAST score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} dvm_keypoint;

typedef struct dvm_orb dvm_orb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (O3/src/ORBextractor.cc:282-339).  max_width/max_height bound the images later passed to
 * dvm_orb_extract (device buffers are sized once here).  device = CUDA ordinal. */
DVM_API int dvm_orb_create(dvm_orb** out, int device, int nfeatures, float scale_factor, int nlevels,
                           int ini_th_fast, int min_th_fast, int max_width, int max_height);
DVM_API void dvm_orb_destroy(dvm_orb* h);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (O3/include/ORBextractor.h:57-67) plus the per-level feature quota
 * (mnFeaturesPerLevel).  Any pointer may be NULL.  Arrays hold nlevels entries. */
DVM_API int dvm_orb_tables(const dvm_orb* h, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int* features_per_level);
/* Upper bound on the number of keypoints one extract call can return (size kps/desc with it). */
DVM_API int dvm_orb_max_keypoints(const dvm_orb* h);

/* int ORBextractor::operator()(image, mask (ignored), keypoints, descriptors, vLappingArea)
 * (O3/src/ORBextractor.cc:876-955).  gray: CV_8UC1 host image, `stride` bytes per row.
 * lap0/lap1 = vLappingArea[0..1] (the mono caller passes {0,1000}, O3/src/Frame.cc:411).
 * kps[cap] / desc[cap*32] receive *n_out entries in the reference's output order; *mono_index
 * receives the reference's return value (-1 for an empty image, with *n_out = 0). */
DVM_API int dvm_orb_extract(dvm_orb* h, const uint8_t* gray, int width, int height, int stride, int lap0, int lap1,
                            dvm_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index);

/* Same computation with the image already in HBM and the results left in HBM (handle-owned
 * buffers, valid until the next extract on this handle).  Enqueues on the handle's stream and
 * returns without synchronising; dvm_orb_sync() waits.  This is the path the per-frame tracker and
 * bench.py's device-resident `value` use. */
DVM_API int dvm_orb_extract_device(dvm_orb* h, const uint8_t* gray_dev, int width, int height, int stride, int lap0,
                                   int lap1);
DVM_API int dvm_orb_sync(dvm_orb* h);
/* Device pointers to the last result: kps (dvm_keypoint[cap]), desc (u8[cap*32]),
 * counts (int32[2] = {n, mono_index}). */
DVM_API int dvm_orb_result_device(const dvm_orb* h, const dvm_keypoint** kps_dev, const uint8_t** desc_dev,
                                  const int32_t** counts_dev);
/* The CUDA stream (cudaStream_t) the handle launches on, for CUDA-event timing by the caller. */
DVM_API void* dvm_orb_stream(const dvm_orb* h);

/* Per-stage device timing for bench.py's roofline: with profiling enabled every extract call records
 * CUDA events on the handle's stream around its four stages {pyramid resize chain, per-cell FAST,
 * octree + output ordering, orientation + blur + rBRIEF} and synchronises; dvm_orb_get_profile returns
 * the number of profiled frames and the summed milliseconds per stage (and resets both). */
DVM_API int dvm_orb_set_profiling(dvm_orb* h, int enable);
DVM_API int dvm_orb_get_profile(dvm_orb* h, int* n_frames, float* stage_ms_sum /* [4] */);

/* Stage read-back for parity tests (valid after an extract; synchronises).
 *  dvm_orb_debug_level_image: mvImagePyramid[level] (O3/include/ORBextractor.h:69), or with
 *    blurred=1 the GaussianBlur'ed working copy the descriptors are sampled from
 *    (O3/src/ORBextractor.cc:919-920).  out holds w*h bytes, tightly packed.
 *  dvm_orb_debug_level_keypoints: which=0 -> vToDistributeKeys (all per-cell FAST keypoints of the
 *    level, order unspecified); which=1 -> the octree's selection in list order
 *    (O3/src/ORBextractor.cc:697-698).  x,y are level pixels relative to the image origin. */
DVM_API int dvm_orb_debug_level_size(const dvm_orb* h, int level, int* w, int* hgt);
DVM_API int dvm_orb_debug_level_image(dvm_orb* h, int level, int blurred, uint8_t* out);
DVM_API int dvm_orb_debug_level_keypoints(dvm_orb* h, int level, int which, int* xs, int* ys, int* responses,
                                          int cap, int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* DVMSLAM_B200_H */
