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

/* Segment test at threshold th: is there a contiguous arc of >= 9 ring pixels all brighter than
 * c + th or all darker than c - th?  Ring comparisons are packed into 16-bit masks; a run of >= 9
 * set bits in the (wrapped) mask is found by shift-and-and. */
static inline int has_run9(uint32_t m16)
{
    uint32_t x = m16 | (m16 << 16);
    uint32_t t = x & (x >> 1);
    t &= t >> 2;
    t &= t >> 4;
    t &= x >> 8;
    return t != 0;
}

static int fast_is_corner(const uint8_t* p, const int* off, int th)
{
    const int c = p[0], hi = c + th, lo = c - th;
    uint32_t b = 0, d = 0;
    for (int k = 0; k < 16; k++) {
        const int v = p[off[k]];
        b |= (uint32_t)(v > hi) << k;
        d |= (uint32_t)(v < lo) << k;
    }
    return has_run9(b) || has_run9(d);
}

/* same value as cvm_fast_measure, with shared partial minima (3-windows, then 9 = 3+3+3) */
static int fast_measure_quick(const uint8_t* p, const int* off)
{
    int r[16], lo3[16], hi3[16];
    for (int k = 0; k < 16; k++) r[k] = p[off[k]];
    for (int i = 0; i < 16; i++) {
        const int a = r[i], b = r[(i + 1) & 15], e = r[(i + 2) & 15];
        int mn = a < b ? a : b; mn = mn < e ? mn : e;
        int mx = a > b ? a : b; mx = mx > e ? mx : e;
        lo3[i] = mn; hi3[i] = mx;
    }
    int best_lo = 0, best_hi = 255;
    for (int i = 0; i < 16; i++) {
        int mn = lo3[i], mx = hi3[i];
        const int l1 = lo3[(i + 3) & 15], l2 = lo3[(i + 6) & 15], h1 = hi3[(i + 3) & 15], h2 = hi3[(i + 6) & 15];
        mn = mn < l1 ? mn : l1; mn = mn < l2 ? mn : l2;
        mx = mx > h1 ? mx : h1; mx = mx > h2 ? mx : h2;
        if (mn > best_lo) best_lo = mn;
        if (mx < best_hi) best_hi = mx;
    }
    const int c = p[0];
    const int a = best_lo - c, b = c - best_hi;
    return a > b ? a : b;
}

/* Row-wise quick reject, written so the compiler vectorises it (u8 saturating arithmetic): a corner
 * needs, on the same side, one pixel of every opposing pair (k, k+8). */
static void fast_row_mask(const uint8_t* p, int w, int stride, int th, uint8_t* mask)
{
    const uint8_t* r[16];
    for (int k = 0; k < 16; k++) r[k] = p + ring_dy[k] * stride + ring_dx[k];
    for (int x = 3; x < w - 3; x++) {
        const int c = p[x];
        const uint8_t hi = (uint8_t)(c + th > 255 ? 255 : c + th), lo = (uint8_t)(c - th < 0 ? 0 : c - th);
        uint8_t br = 1, dk = 1;
        for (int k = 0; k < 8; k++) {
            const uint8_t a = r[k][x], b = r[k + 8][x];
            br &= (uint8_t)((a > hi) | (b > hi));
            dk &= (uint8_t)((a < lo) | (b < lo));
        }
        mask[x] = (uint8_t)(br | dk);
    }
}

int cvm_fast_detect(const uint8_t* img, int w, int h, int stride, int th,
                    cvm_fast_kp* out, int cap)
{
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = ring_dy[k] * stride + ring_dx[k];
    /* three-row rolling score buffer, like cv::FAST: non-corners score 0 */
    uint8_t* buf = (uint8_t*)calloc((size_t)4 * (w + 2), 1);
    uint8_t* rows[3] = { buf + 1, buf + (w + 2) + 1, buf + 2 * (w + 2) + 1 };
    uint8_t* mask = buf + 3 * (w + 2);
    int n = 0;
    for (int y = 3; y <= h - 3; y++) {
        uint8_t* cur = rows[(y - 3) % 3];
        memset(cur - 1, 0, w + 2);
        if (y < h - 3) {
            const uint8_t* p = img + (size_t)y * stride;
            fast_row_mask(p, w, stride, th, mask);
            for (int x = 3; x < w - 3; x++)
                if (mask[x] && fast_is_corner(p + x, off, th))
                    cur[x] = (uint8_t)(fast_measure_quick(p + x, off) - 1);
        }
        if (y == 3) continue;
        /* non-max suppression of row y-1 against rows y-2, y-1, y (strictly greater than all 8) */
        const uint8_t* prev = rows[(y - 4) % 3];
        const uint8_t* pprev = rows[(y - 2) % 3]; /* == (y-5) mod 3: row y-2 */
        for (int x = 3; x < w - 3; x++) {
            const int s = prev[x];
            if (!s) continue;
            if (s > prev[x - 1] && s > prev[x + 1] && s > pprev[x - 1] && s > pprev[x] && s > pprev[x + 1] &&
                s > cur[x - 1] && s > cur[x] && s > cur[x + 1]) {
                if (n < cap) { out[n].x = x; out[n].y = y - 1; out[n].response = s; }
                n++;
            }
        }
    }
    free(buf);
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
    uint16_t* tmp = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * h);
    uint8_t* pad = (uint8_t*)malloc((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int i = 0; i < 3; i++) {
            pad[i] = S[reflect101(i - 3, w)];
            pad[w + 3 + i] = S[reflect101(w + i, w)];
        }
        memcpy(pad + 3, S, w);
        uint16_t* T = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) /* <= 256*255, fits 16 bits */
            T[x] = (uint16_t)(18 * (pad[x] + pad[x + 6]) + 34 * (pad[x + 1] + pad[x + 5]) +
                              48 * (pad[x + 2] + pad[x + 4]) + 56 * pad[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* r[7];
        for (int j = 0; j < 7; j++) r[j] = tmp + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            const uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                                 48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * (uint32_t)r[3][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(tmp);
    free(pad);
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

/* cv::undistortPoints -- call site O3/src/Frame.cc:806 (UndistortKeyPoints) and :836 (ComputeImageBounds).
 * Model of cvUndistortPointsInternal (calib3d/undistort.dispatch.cpp) for the 5-coefficient radial-tangential
 * model: normalise with (u - cx) * (1 / fx), five fixed-point iterations
 *     r2 = x^2 + y^2, icdist = 1 / (1 + ((k3 r2 + k2) r2 + k1) r2),
 *     dx = 2 p1 x y + p2 (r2 + 2 x^2), dy = p1 (r2 + 2 y^2) + 2 p2 x y,  x = (x0 - dx) icdist, y = (y0 - dy) icdist
 * all in double, then x' = (fx' x + cx') * (1 / 1), rounded to float.  Bit-exact against cv2 4.13.0 on 15 000 points
 * and three coefficient sets (tests/golden/undistort.npz).  The k4..k6 / s1..s4 terms are written out with zero
 * coefficients exactly as OpenCV evaluates them, because they change the rounding of icdist and dx, dy. */
void cvm_undistort_points(const float* xy, int n, const float* K, const float* dist5, const float* P, float* out)
{
    double k[12] = { 0 };
    for (int i = 0; i < 5; i++) k[i] = (double)dist5[i];
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double pfx = P[0], pfy = P[1], pcx = P[2], pcy = P[3];
    for (int i = 0; i < n; i++) {
        double x = ((double)xy[2 * i] - cx) * ifx, y = ((double)xy[2 * i + 1] - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = ((double)xy[2 * i] - cx) * ifx; y = ((double)xy[2 * i + 1] - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = pfx * x + 0. * y + pcx, yy = 0. * x + pfy * y + pcy, ww = 1. / (0. * x + 0. * y + 1.);
        out[2 * i] = (float)(xx * ww);
        out[2 * i + 1] = (float)(yy * ww);
    }
}
