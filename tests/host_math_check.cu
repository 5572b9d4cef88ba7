// tests/host_math_check.cu -- CPU-side check of the product's host/device math header
// (dvmslam_b200/csrc/orb_math.cuh) against the oracle (oracle/cvmodels.c) and libstdc++/libm.
// Built and run by tests/test_host_math.py; prints "OK" lines and exits 0 on success.
#include "../dvmslam_b200/csrc/orb_math.cuh"
#include "../oracle/cvmodels.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>

using namespace dvm;

static int fails = 0;
#define CHECK(cond, ...)                         \
    do {                                         \
        if (!(cond)) {                           \
            if (fails < 20) { printf("FAIL: " __VA_ARGS__); printf("\n"); } \
            fails++;                             \
        }                                        \
    } while (0)

int main()
{
    std::mt19937 rng(12345);
    // ---- FAST measure: packed and scalar forms vs the oracle's literal 16-arc definition ----
    {
        const int W = 64, H = 64;
        std::vector<uint8_t> img(W * H);
        static const int dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
        static const int dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };
        for (int trial = 0; trial < 200; trial++) {
            int mode = trial % 4;
            for (auto& p : img) {
                int v = rng() & 255;
                if (mode == 1) v = 100 + (rng() % 40);
                if (mode == 2) v = (rng() & 1) ? 255 : 0;
                if (mode == 3) v = 128 + (int)(rng() % 7) - 3;
                p = (uint8_t)v;
            }
            for (int y = 3; y < H - 3; y++)
                for (int x = 3; x < W - 4; x += 2) {
                    uint32_t ring[16];
                    int ra[16], rb[16];
                    for (int k = 0; k < 16; k++) {
                        ra[k] = img[(y + dy[k]) * W + x + dx[k]];
                        rb[k] = img[(y + dy[k]) * W + x + 1 + dx[k]];
                        ring[k] = ra[k] | (rb[k] << 16);
                    }
                    int mA, mB;
                    fast_measure_x2(ring, img[y * W + x], img[y * W + x + 1], &mA, &mB);
                    int oA = std::max(cvm_fast_measure(img.data(), W, x, y), 0);
                    int oB = std::max(cvm_fast_measure(img.data(), W, x + 1, y), 0);
                    CHECK(mA == oA && mB == oB, "fast_measure_x2 (%d,%d): %d %d vs %d %d", x, y, mA, mB, oA, oB);
                    int sA = std::max(fast_measure_scalar(ra, img[y * W + x]), 0);
                    CHECK(sA == oA, "fast_measure_scalar (%d,%d): %d vs %d", x, y, sA, oA);
                }
        }
        printf("OK fast measure\n");
    }
    // ---- fastAtan2 ----
    {
        std::uniform_int_distribution<int> d(-3000000, 3000000);
        for (int i = 0; i < 2000000; i++) {
            int y = d(rng), x = d(rng);
            if (i % 50 == 0) y = 0;
            if (i % 50 == 1) x = 0;
            if (i % 50 == 2) y = x;
            if (i % 50 == 3) y = -x;
            if (i % 1000 == 4) x = y = 0;
            float a = fast_atan2_deg((float)y, (float)x), b = cvm_fast_atan2((float)y, (float)x);
            CHECK(a == b, "atan2(%d,%d) %a vs %a", y, x, a, b);
        }
        printf("OK fastAtan2\n");
    }
    // ---- sinf / cosf vs this box's libm, strided over [0, 2*pi] plus every angle fastAtan2*pi/180 can hit nearby ----
    {
        long n = 0;
        for (uint32_t u = 0; u < 0x40C90FDCu + 64; u += 97) {
            float a;
            memcpy(&a, &u, 4);
            CHECK(glibc_sincosf(a, 0) == sinf(a), "sinf(%a)", a);
            CHECK(glibc_sincosf(a, 1) == cosf(a), "cosf(%a)", a);
            n++;
        }
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        for (int i = 0; i <= 3600000; i++) {
            float deg = (float)i * 1e-4f, a = deg * factorPI;
            CHECK(glibc_sincosf(a, 0) == sinf(a), "sinf(%a)", a);
            CHECK(glibc_sincosf(a, 1) == cosf(a), "cosf(%a)", a);
        }
        printf("OK sincosf (%ld strided + 3600001 degree-grid samples)\n", n);
    }
    // ---- libstdc++ std::sort emulation: identical permutation on tie-heavy inputs ----
    {
        typedef std::pair<int, int> P; // (key, id): comparator looks at the key only
        auto run = [&](std::vector<P> v) {
            std::vector<unsigned long long> mine(v.size());
            for (size_t i = 0; i < v.size(); i++) mine[i] = ((unsigned long long)(unsigned)v[i].first << 32) | (unsigned)v[i].second;
            std::sort(v.begin(), v.end(), [](P& a, P& b) { return a.first < b.first; });
            std::vector<unsigned long long> coop(mine), tmp(v.size());
            std::vector<int> li(v.size() + 1), ri(v.size() + 1);
            libstdcxx_sort(mine.data(), (int)mine.size(), KeyHi32Less());
            // the stopper-pairing form the GPU runs cooperatively (libstdcxx_sort_cta)
            libstdcxx_sort_stopper_model(coop.data(), tmp.data(), li.data(), ri.data(), (int)coop.size(), KeyHi32Less());
            for (size_t i = 0; i < v.size(); i++)
                if ((int)(mine[i] & 0xffffffffu) != v[i].second || coop[i] != mine[i]) return false;
            return true;
        };
        int cases = 0;
        for (int trial = 0; trial < 4000; trial++) {
            int n = trial < 200 ? trial : (int)(rng() % 3000);
            int keyrange = 1 + (int)(rng() % (trial % 3 == 0 ? 4 : trial % 3 == 1 ? 40 : 100000));
            std::vector<P> v(n);
            for (int i = 0; i < n; i++) v[i] = P((int)(rng() % keyrange), i);
            if (trial % 7 == 0) std::sort(v.begin(), v.end());
            if (trial % 11 == 0) std::reverse(v.begin(), v.end());
            CHECK(run(v), "sort mismatch n=%d keyrange=%d", n, keyrange);
            cases++;
        }
        // median-of-3 killer (forces the heap-sort fallback)
        for (int n : { 64, 500, 1000, 4096 }) {
            std::vector<P> v(n);
            int k = n / 2;
            for (int i = 1; i <= k; i++) {
                if (i % 2 == 1) { v[i - 1] = P(i, i - 1); v[i] = P(k + i, i); }
                v[k + i - 1] = P(2 * i, k + i - 1);
            }
            CHECK(run(v), "sort mismatch on median-of-3 killer n=%d", n);
            cases++;
        }
        printf("OK libstdc++ sort emulation (%d cases)\n", cases);
    }
    // ---- resize coefficients: rebuild a resize from resize_coef and compare with the oracle ----
    {
        for (int trial = 0; trial < 12; trial++) {
            int sw = 67 + (int)(rng() % 900), sh = 67 + (int)(rng() % 600);
            int dw = (int)lrintf((float)sw * (1.0f / 1.2f)), dh = (int)lrintf((float)sh * (1.0f / 1.2f));
            if (trial % 4 == 3) { dw = sw / 2 + 3; dh = sh + 7; }
            std::vector<uint8_t> src(sw * sh), ref(dw * dh), out(dw * dh);
            for (auto& p : src) p = (uint8_t)(rng() & 255);
            cvm_resize_linear_u8(src.data(), sw, sh, sw, ref.data(), dw, dh, dw);
            for (int y = 0; y < dh; y++) {
                int sy0, sy1; short b0, b1;
                resize_coef(y, sh, dh, false, &sy0, &sy1, &b0, &b1);
                for (int x = 0; x < dw; x++) {
                    int sx0, sx1; short a0, a1;
                    resize_coef(x, sw, dw, true, &sx0, &sx1, &a0, &a1);
                    int R0 = src[sy0 * sw + sx0] * a0 + src[sy0 * sw + sx1] * a1;
                    int R1 = src[sy1 * sw + sx0] * a0 + src[sy1 * sw + sx1] * a1;
                    int v = (((b0 * (R0 >> 4)) >> 16) + ((b1 * (R1 >> 4)) >> 16) + 2) >> 2;
                    out[y * dw + x] = (uint8_t)std::min(std::max(v, 0), 255);
                }
            }
            CHECK(out == ref, "resize %dx%d -> %dx%d", sw, sh, dw, dh);
        }
        printf("OK resize coefficients\n");
    }
    if (fails) { printf("%d FAILURES\n", fails); return 1; }
    printf("ALL OK\n");
    return 0;
}
