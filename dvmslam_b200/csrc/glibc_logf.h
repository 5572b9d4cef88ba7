/* glibc_logf.h -- logf as glibc computes it (sysdeps/ieee754/flt-32/e_logf.c, the ARM optimized-routines algorithm:
 * 16-entry table of (1/c, log c), degree-3 polynomial in double, one final rounding), restated so that
 * MapPoint::PredictScale (O3/src/MapPoint.cc:557-587: ceil(log(ratio) / mfLogScaleFactor) on floats, i.e. libm's logf)
 * gives the same level on the device as on the reference's host.  glibc's logf is NOT correctly rounded (0.16 % of the
 * inputs in [2^-12, 2^12] differ from (float)log((double)x)), so the algorithm itself has to be followed.  Normal
 * positive inputs only (a distance ratio).  oracle/check_logf.c compares this function with the host's libm over every
 * float in [2^-20, 2^20]: 0 differences (tests/test_host_math.py runs it). */
#ifndef DVM_GLIBC_LOGF_H
#define DVM_GLIBC_LOGF_H
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DVM_LOGF_HD __host__ __device__ inline
#else
#define DVM_LOGF_HD static inline
#endif

DVM_LOGF_HD float dvm_glibc_logf(float x)
{
    const double T[16][2] = {
        { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
        { 0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2 },  { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
        { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3 },
        { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
        { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1p+0, 0x0p+0 },
        { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 },  { 0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4 },
        { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 },  { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3 },
        { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 },  { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 },
    };
    const double Ln2 = 0x1.62e42fefa39efp-1;
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> (23 - 4)) % 16u);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & (0x1ffu << 23));
    float zf;
    memcpy(&zf, &iz, 4);
    const double z = (double)zf;
#if defined(__CUDA_ARCH__)
    const double r = __dsub_rn(__dmul_rn(z, T[i][0]), 1.0);
    const double y0 = __dadd_rn(T[i][1], __dmul_rn((double)k, Ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
#else
    const double r = z * T[i][0] - 1;
    const double y0 = T[i][1] + (double)k * Ln2;
    const double r2 = r * r;
    double y = A1 * r + A2;
    y = A0 * r2 + y;
    y = y * r2 + (y0 + r);
#endif
    return (float)y;
}
#endif
