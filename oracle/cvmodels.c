/*
 * oracle/cvmodels.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See cvmodels.h.
 * Compile with -ffp-contract=off: the float expressions below must not be
 * FMA-contracted (cv::fastAtan2's scalar path and the resize coefficient
 * computation are plain IEEE float32 operations).
 */
#include "cvmodels.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

int cvm_round_f(float v) { return (int)lrintf(v); }
int cvm_round_d(double v) { return (int)lrint(v); }

/* ------------------------------------------------------------------ resize
 * Model of OpenCV's 8-bit INTER_LINEAR path (imgproc/resize.cpp: integer
 * coefficients with INTER_RESIZE_COEF_BITS = 11, HResizeLinear then
 * VResizeLinear<uchar,int,short>):
 *   fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), fx -= s
 *   horizontal clamp: s < 0 -> (0, fx = 0); s >= n-1 -> (n-1, fx = 0)
 *   vertical: rows are clipped, coefficients are kept
 *   a = { rint((1-fx)*2048), rint(fx*2048) } in float32
 *   out = (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2
 */
static short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

void cvm_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride)
{
    double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    int* xofs = (int*)malloc(sizeof(int) * dw);
    short* ialpha = (short*)malloc(sizeof(short) * 2 * dw);
    int* rows[2];
    rows[0] = (int*)malloc(sizeof(int) * dw);
    rows[1] = (int*)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = sat_short((int)lrintf((1.f - fx) * 2048.f));
        ialpha[2 * dx + 1] = sat_short((int)lrintf(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = sat_short((int)lrintf((1.f - fy) * 2048.f));
        short b1 = sat_short((int)lrintf(fy * 2048.f));
        for (int k = 0; k < 2; k++) {
            int y = sy + k;
            if (y < 0) y = 0;
            if (y > sh - 1) y = sh - 1;
            const uint8_t* S = src + (size_t)y * sstride;
            int* R = rows[k];
            for (int dx = 0; dx < dw; dx++) {
                int sx = xofs[dx];
                int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                R[dx] = S[sx] * ialpha[2 * dx] + S[sx1] * ialpha[2 * dx + 1];
            }
        }
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (rows[0][dx] >> 4)) >> 16) + ((b1 * (rows[1][dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(ialpha); free(rows[0]); free(rows[1]);
}

/* -------------------------------------------------------------------- FAST
 * Bresenham ring of radius 3, clockwise from (0,3) (features2d/fast.cpp,
 * makeOffsets, patternSize 16). */
static const int ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

int cvm_fast_measure(const uint8_t* img, int stride, int x, int y)
{
    int c = img[(size_t)y * stride + x];
    int d[16];
    for (int k = 0; k < 16; k++)
        d[k] = (int)img[(size_t)(y + ring_dy[k]) * stride + (x + ring_dx[k])] - c;
    int best = -255;
    for (int s = 0; s < 16; s++) {
        int mn = 255, mx = -255;
        for (int k = 0; k < 9; k++) {
            int v = d[(s + k) & 15];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        /* bright arc: min(ring - c);  dark arc: min(c - ring) = -max(ring - c) */
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    return best;
}

int cvm_fast_detect(const uint8_t* img, int w, int h, int stride, int th,
                    cvm_fast_kp* out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int* score = (int*)calloc((size_t)w * h, sizeof(int));
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int m = cvm_fast_measure(img, stride, x, y);
            score[(size_t)y * w + x] = m > th ? m - 1 : 0;
        }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = score[(size_t)y * w + x];
            if (s <= 0) continue;
            const int* p = score + (size_t)y * w + x;
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] &&
                s > p[w - 1] && s > p[w] && s > p[w + 1]) {
                if (n < cap) { out[n].x = x; out[n].y = y; out[n].response = s; }
                n++;
            }
        }
    free(score);
    return n;
}

/* ------------------------------------------------------------ GaussianBlur
 * 8-bit fixed-point separable path (imgproc/smooth.dispatch.cpp, fixed-point
 * kernel for ksize 7 sigma 2): k = {18,34,48,56,48,34,18}/256, row pass in Q8,
 * column pass in Q16, one rounding at the end. */
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

void cvm_gaussian7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int k[7] = { 18, 34, 48, 56, 48, 34, 18 };
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int i = 0; i < 7; i++) acc += k[i] * S[reflect101(x + i - 3, w)];
            tmp[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int j = 0; j < 7; j++) acc += k[j] * tmp[(size_t)reflect101(y + j - 3, h) * w + x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(tmp);
}

/* --------------------------------------------------------------- fastAtan2
 * core/mathfuncs_core: scalar float32 polynomial, no FMA. */
float cvm_fast_atan2(float y, float x)
{
    static const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    static const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    static const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    static const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}
