/* oracle/check_logf.c -- TEST INFRASTRUCTURE.  MapPoint::PredictScale (O3/src/MapPoint.cc:557-587) takes std::log of a
 * float, i.e. the host libm's logf, which is not correctly rounded.  The kernels follow glibc's algorithm
 * (dvmslam_b200/csrc/glibc_logf.h); this program compares that restatement with the libm of the box it runs on over
 * every float in [2^-20, 2^20] and prints the number of differing inputs (exit status 1 if any). */
#include <math.h>
#include <stdio.h>
#include "../dvmslam_b200/csrc/glibc_logf.h"
int main(void)
{
    uint32_t lo, hi;
    float a = ldexpf(1.f, -20), b = ldexpf(1.f, 20);
    memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
    unsigned long long diff = 0, n = 0;
    for (uint32_t u = lo; u <= hi; u++, n++) {
        float x; memcpy(&x, &u, 4);
        volatile float y0 = logf(x);
        volatile float y1 = dvm_glibc_logf(x);
        if (y0 != y1) { if (diff < 5) printf("x=%a logf=%a restated=%a\n", x, y0, y1); diff++; }
    }
    printf("%llu inputs, %llu differ\n", n, diff);
    return diff != 0;
}
