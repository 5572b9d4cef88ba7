/*
 * oracle/cvmodels.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatements (plain C) of the four external OpenCV primitives that the
 * reference's ORB front end calls but does not contain:
 *   cv::resize(INTER_LINEAR, 8-bit)   call site  O3/src/ORBextractor.cc:967
 *   cv::FAST(img,kps,th,true)         call sites O3/src/ORBextractor.cc:653,669
 *   cv::GaussianBlur(7x7, sigma 2)    call site  O3/src/ORBextractor.cc:920
 *   cv::fastAtan2                     call site  O3/src/ORBextractor.cc:98
 * (O3/ = /root/reference/src/slam_system/orb_slam3/).  OpenCV itself is an
 * un-vendored dependency of the reference (find_package(OpenCV 4.2),
 * O3/CMakeLists.txt:36); these models are pinned bit-exactly against the
 * cv2 4.13.0 wheel in tests/test_oracle_cv2.py and via the committed golden
 * vectors under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call anything in oracle/.
 */
#ifndef DVM_ORACLE_CVMODELS_H
#define DVM_ORACLE_CVMODELS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* cvRound: round-half-to-even (SSE cvtss2si / lrint in default rounding mode) */
int cvm_round_f(float v);
int cvm_round_d(double v);

/* cv::resize(src, dst, Size(dw,dh), 0, 0, INTER_LINEAR) for CV_8UC1. */
void cvm_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride);

/* Raw FAST-9-16 arc measure m(x,y) = max over the 16 contiguous 9-arcs of
 * max(min(ring - c), min(c - ring)).  A pixel is a corner at threshold th iff
 * m > th and then cv::FAST's response is m - 1.  Undefined (returns 0) when the
 * ring leaves the image. */
int cvm_fast_measure(const uint8_t* img, int stride, int x, int y);

typedef struct { int x, y, response; } cvm_fast_kp;
/* cv::FAST(img(w x h), kps, th, nonmaxSuppression=true), TYPE_9_16.
 * Output row-major.  Returns number of keypoints (writes at most cap). */
int cvm_fast_detect(const uint8_t* img, int w, int h, int stride, int th,
                    cvm_fast_kp* out, int cap);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101), CV_8UC1 */
void cvm_gaussian7_u8(const uint8_t* src, int w, int h, int sstride,
                      uint8_t* dst, int dstride);

/* cv::fastAtan2(y, x) in degrees, [0,360) */
float cvm_fast_atan2(float y, float x);
/* cv::undistortPoints(src, dst, K, distCoeffs(k1,k2,p1,p2,k3), noArray(), P): 5 fixed iterations in double
 * (default TermCriteria(MAX_ITER, 5, 0.01)), R = I; K, P = fx, fy, cx, cy.  xy / out: n interleaved float pairs. */
void cvm_undistort_points(const float* xy, int n, const float* K, const float* dist5, const float* P, float* out);

#ifdef __cplusplus
}
#endif
#endif
