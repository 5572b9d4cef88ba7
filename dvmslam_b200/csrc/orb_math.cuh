// orb_math.cuh -- bit-exact scalar building blocks of the ORB front end, usable from host and
// device so that the arithmetic can be unit-tested on the CPU (tests/host_math_check.cu) before it
// ever runs on a GPU.  Every float expression is spelled with explicit round-to-nearest
// single operations: the parity convention is "IEEE float32, op by op, no FMA contraction"
// (DESIGN.md), and nvcc would otherwise fuse a*b+c.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace dvm {

#ifdef __CUDA_ARCH__
#define DVM_FMUL(a, b) __fmul_rn((a), (b))
#define DVM_FADD(a, b) __fadd_rn((a), (b))
#define DVM_FSUB(a, b) __fsub_rn((a), (b))
#define DVM_FDIV(a, b) __fdiv_rn((a), (b))
#define DVM_DMUL(a, b) __dmul_rn((a), (b))
#define DVM_DADD(a, b) __dadd_rn((a), (b))
#define DVM_DSUB(a, b) __dsub_rn((a), (b))
#else
// host build of this header must use -ffp-contract=off (nvcc: -Xcompiler -ffp-contract=off)
#define DVM_FMUL(a, b) ((float)(a) * (float)(b))
#define DVM_FADD(a, b) ((float)(a) + (float)(b))
#define DVM_FSUB(a, b) ((float)(a) - (float)(b))
#define DVM_FDIV(a, b) ((float)(a) / (float)(b))
#define DVM_DMUL(a, b) ((double)(a) * (double)(b))
#define DVM_DADD(a, b) ((double)(a) + (double)(b))
#define DVM_DSUB(a, b) ((double)(a) - (double)(b))
#endif

// cvRound(float): round half to even (O3/src/ORBextractor.cc:78,106,110 use it on floats)
__host__ __device__ inline int cv_round(float v)
{
#ifdef __CUDA_ARCH__
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}

// ---- cv::fastAtan2 (degrees), call site O3/src/ORBextractor.cc:98 -------------------------------
__host__ __device__ inline float fast_atan2_deg(float y, float x)
{
    // pK = (float)cK * (float)(180/pi), products taken in float32
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = DVM_FMUL(0.9997878412794807f, scale);
    const float p3 = DVM_FMUL(-0.3258083974640975f, scale);
    const float p5 = DVM_FMUL(0.1555786518463281f, scale);
    const float p7 = DVM_FMUL(-0.04432655554792128f, scale);
    const float eps = 2.220446049250313e-16f; // (float)DBL_EPSILON
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = DVM_FDIV(ay, DVM_FADD(ax, eps));
        c2 = DVM_FMUL(c, c);
        a = DVM_FMUL(DVM_FADD(DVM_FMUL(DVM_FADD(DVM_FMUL(DVM_FADD(DVM_FMUL(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = DVM_FDIV(ax, DVM_FADD(ay, eps));
        c2 = DVM_FMUL(c, c);
        a = DVM_FSUB(90.f,
                     DVM_FMUL(DVM_FADD(DVM_FMUL(DVM_FADD(DVM_FMUL(DVM_FADD(DVM_FMUL(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = DVM_FSUB(180.f, a);
    if (y < 0) a = DVM_FSUB(360.f, a);
    return a;
}

// ---- glibc sinf/cosf (flt-32 sincosf: double polynomial after a fast quadrant reduction) --------
// computeOrbDescriptor calls cos/sin on a float (O3/src/ORBextractor.cc:103-104); libm's sinf/cosf
// are not correctly rounded, so the same algorithm is evaluated here in double, op by op.
// Valid for |x| < 120 (the caller's range is [0, 2*pi]).
__host__ __device__ inline float glibc_sincosf_poly(double x, double x2, int neg_tab, int n)
{
    const double c0 = neg_tab ? -0x1p0 : 0x1p0;
    const double c1 = neg_tab ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2;
    const double c2 = neg_tab ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5;
    const double c3 = neg_tab ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10;
    const double c4 = neg_tab ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = DVM_DMUL(x, x2);
        double t1 = DVM_DADD(s2, DVM_DMUL(x2, s3));
        double x7 = DVM_DMUL(x3, x2);
        double s = DVM_DADD(x, DVM_DMUL(x3, s1));
        return (float)DVM_DADD(s, DVM_DMUL(x7, t1));
    }
    double x4 = DVM_DMUL(x2, x2);
    double u2 = DVM_DADD(c3, DVM_DMUL(x2, c4));
    double u1 = DVM_DADD(c1, DVM_DMUL(x2, c2));
    double x6 = DVM_DMUL(x4, x2);
    double c = DVM_DADD(c0, DVM_DMUL(x2, u1));
    return (float)DVM_DADD(c, DVM_DMUL(x6, u2));
}

__host__ __device__ inline uint32_t f32_top12(float v)
{
#ifdef __CUDA_ARCH__
    return (__float_as_uint(v) >> 20) & 0x7ffu;
#else
    uint32_t u;
    memcpy(&u, &v, 4);
    return (u >> 20) & 0x7ffu;
#endif
}

__host__ __device__ inline float glibc_sincosf(float y, int is_cos)
{
    double x = (double)y;
    if (f32_top12(y) < 0x3f4u /* top12(pi/4) */) {
        if (f32_top12(y) < 0x398u /* top12(2^-12) */) return is_cos ? 1.0f : y;
        return glibc_sincosf_poly(x, DVM_DMUL(x, x), 0, is_cos);
    }
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    double r = DVM_DMUL(x, hpi_inv);
    int n = ((int32_t)r + 0x800000) >> 24;
    x = DVM_DSUB(x, DVM_DMUL((double)n, hpi));
    double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return glibc_sincosf_poly(DVM_DMUL(x, sgn), DVM_DMUL(x, x), (n & 2) ? 1 : 0, n ^ is_cos);
}

// ---- FAST-9-16 arc measure -----------------------------------------------------------------------
// m = max over the 16 contiguous 9-arcs of max(min(ring) - c, c - max(ring)); a pixel is a corner at
// threshold t iff m > t, and cv::FAST's response is then m - 1 (model verified against cv2, see
// oracle/cvmodels.c).  Scalar form: r[k] are the 16 ring pixels clockwise from (0,+3).
__host__ __device__ inline int fast_measure_scalar(const int* r, int c)
{
    int lo3[16], hi3[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int a = r[i], b = r[(i + 1) & 15], d = r[(i + 2) & 15];
        lo3[i] = min(a, min(b, d));
        hi3[i] = max(a, max(b, d));
    }
    int best_lo = 0, best_hi = 255;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int lo9 = min(lo3[i], min(lo3[(i + 3) & 15], lo3[(i + 6) & 15]));
        int hi9 = max(hi3[i], max(hi3[(i + 3) & 15], hi3[(i + 6) & 15]));
        best_lo = max(best_lo, lo9);
        best_hi = min(best_hi, hi9);
    }
    return max(best_lo - c, c - best_hi);
}

// Packed form: two pixels per 32-bit word (u16 lanes).  On sm_90+ the min3/max3 intrinsics are
// single DPX instructions (VIMNMX3).
#ifdef __CUDA_ARCH__
__device__ inline uint32_t pk_min3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ inline uint32_t pk_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
#else
inline uint32_t pk_lane(uint32_t lo, uint32_t hi) { return (lo & 0xffffu) | (hi << 16); }
inline uint32_t pk_min3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t lo = min(a & 0xffffu, min(b & 0xffffu, c & 0xffffu)), hi = min(a >> 16, min(b >> 16, c >> 16));
    return pk_lane(lo, hi);
}
inline uint32_t pk_max3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t lo = max(a & 0xffffu, max(b & 0xffffu, c & 0xffffu)), hi = max(a >> 16, max(b >> 16, c >> 16));
    return pk_lane(lo, hi);
}
#endif

// r[k]: ring value of pixel A in bits 0..15 and of pixel B in bits 16..31; cA/cB the two centres.
// Writes the two measures clamped to [0,255].
__host__ __device__ inline void fast_measure_x2(const uint32_t* r, int cA, int cB, int* mA, int* mB)
{
    uint32_t lo3[16], hi3[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo3[i] = pk_min3(r[i], r[(i + 1) & 15], r[(i + 2) & 15]);
        hi3[i] = pk_max3(r[i], r[(i + 1) & 15], r[(i + 2) & 15]);
    }
    uint32_t lo9[16], hi9[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo9[i] = pk_min3(lo3[i], lo3[(i + 3) & 15], lo3[(i + 6) & 15]);
        hi9[i] = pk_max3(hi3[i], hi3[(i + 3) & 15], hi3[(i + 6) & 15]);
    }
    uint32_t bl = lo9[15], bh = hi9[15];
#pragma unroll
    for (int i = 0; i < 14; i += 2) {
        bl = pk_max3(bl, lo9[i], lo9[i + 1]);
        bh = pk_min3(bh, hi9[i], hi9[i + 1]);
    }
    bl = pk_max3(bl, lo9[14], lo9[14]);
    bh = pk_min3(bh, hi9[14], hi9[14]);
    int a = max((int)(bl & 0xffffu) - cA, cA - (int)(bh & 0xffffu));
    int b = max((int)(bl >> 16) - cB, cB - (int)(bh >> 16));
    *mA = max(a, 0);
    *mB = max(b, 0);
}

// ---- cv::resize INTER_LINEAR 8-bit coefficients (call site O3/src/ORBextractor.cc:967) -----------
// For destination index d on an axis of src_n -> dst_n pixels: source index s, s1 and the Q11
// weights (a0, a1).  `clamp_coef` selects the horizontal rule (weights zeroed at the clamp); the
// vertical rule keeps the weights and only clips the row index.
__host__ inline void resize_coef(int d, int src_n, int dst_n, bool horizontal, int* s0, int* s1, short* a0, short* a1)
{
    double scale = (double)src_n / dst_n;
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (horizontal) {
        if (s < 0) { f = 0; s = 0; }
        if (s >= src_n - 1) { f = 0; s = src_n - 1; }
        *s0 = s;
        *s1 = s + 1 < src_n ? s + 1 : src_n - 1;
    } else {
        *s0 = s < 0 ? 0 : (s > src_n - 1 ? src_n - 1 : s);
        *s1 = s + 1 < 0 ? 0 : (s + 1 > src_n - 1 ? src_n - 1 : s + 1);
    }
    long w0 = lrintf((1.f - f) * 2048.f), w1 = lrintf(f * 2048.f);
    *a0 = (short)(w0 > 32767 ? 32767 : w0 < -32768 ? -32768 : w0);
    *a1 = (short)(w1 > 32767 ? 32767 : w1 < -32768 ? -32768 : w1);
}

// ---- libstdc++ std::sort, restated -------------------------------------------------------------------
// DistributeOctTree sorts (size, node) pairs with a comparator that ties on equal (size, UL.x)
// (O3/src/ORBextractor.cc:402-417,544); which of the tied nodes is split first -- and hence the
// output order -- is whatever libstdc++'s introsort leaves.  This is that algorithm
// (bits/stl_algo.h: __introsort_loop / __unguarded_partition_pivot / __final_insertion_sort, heap
// fallback from bits/stl_heap.h), iterative, on an array of T with a strict-weak `less`.
// comparator on the high 32 bits of a packed (key << 32 | payload) element
struct KeyHi32Less {
    __host__ __device__ bool operator()(unsigned long long a, unsigned long long b) const { return (a >> 32) < (b >> 32); }
};

template <typename T, typename Less>
__host__ __device__ inline void stdsort_adjust_heap(T* first, int hole, int len, T value, Less less)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <typename T, typename Less>
__host__ __device__ inline void stdsort_heapsort(T* first, int len, Less less)
{
    if (len >= 2) { // __make_heap
        int parent = (len - 2) / 2;
        while (true) {
            T v = first[parent];
            stdsort_adjust_heap(first, parent, len, v, less);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len; // __sort_heap
    while (last > 1) {
        --last;
        T v = first[last];
        first[last] = first[0];
        stdsort_adjust_heap(first, 0, last, v, less);
    }
}

template <typename T, typename Less>
__host__ __device__ inline void stdsort_unguarded_linear_insert(T* a, int last, Less less)
{
    T val = a[last];
    int next = last - 1;
    while (less(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

template <typename T, typename Less>
__host__ __device__ inline void stdsort_insertion_sort(T* a, int first, int last, Less less)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            T val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = val;
        } else {
            stdsort_unguarded_linear_insert(a, i, less);
        }
    }
}

template <typename T, typename Less>
__host__ __device__ inline void libstdcxx_sort(T* a, int n, Less less)
{
    if (n <= 0) return;
    // __introsort_loop with an explicit stack (sub-ranges are disjoint, so their order is free)
    int stack_first[64], stack_last[64], stack_depth[64];
    int sp = 0;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    stack_first[sp] = 0; stack_last[sp] = n; stack_depth[sp] = 2 * lg; sp++;
    while (sp > 0) {
        --sp;
        int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                stdsort_heapsort(a + first, last - first, less);
                break;
            }
            --depth;
            // __unguarded_partition_pivot
            int mid = first + (last - first) / 2;
            {
                int ia = first + 1, ib = mid, ic = last - 1, pick;
                if (less(a[ia], a[ib])) {
                    if (less(a[ib], a[ic])) pick = ib;
                    else if (less(a[ia], a[ic])) pick = ic;
                    else pick = ia;
                } else if (less(a[ia], a[ic])) pick = ia;
                else if (less(a[ib], a[ic])) pick = ic;
                else pick = ib;
                T t = a[first]; a[first] = a[pick]; a[pick] = t;
            }
            int lo = first + 1, hi = last;
            while (true) {
                while (less(a[lo], a[first])) ++lo;
                --hi;
                while (less(a[first], a[hi])) --hi;
                if (!(lo < hi)) break;
                T t = a[lo]; a[lo] = a[hi]; a[hi] = t;
                ++lo;
            }
            int cut = lo;
            // recurse on [cut,last), loop on [first,cut)
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        stdsort_insertion_sort(a, 0, 16, less);
        for (int i = 16; i != n; ++i) stdsort_unguarded_linear_insert(a, i, less);
    } else {
        stdsort_insertion_sort(a, 0, n, less);
    }
}


// Scalar model of libstdcxx_sort_cta below (same stopper-pairing partition, same rank-based final pass);
// tests/host_math_check.cu runs it against std::sort on the CPU.
template <typename T, typename Less>
__host__ inline void libstdcxx_sort_stopper_model(T* a, T* tmp, int* li, int* ri, int n, Less less)
{
    if (n <= 0) return;
    if (n > 16) {
        int st_first[64], st_last[64], st_depth[64], sp = 0, lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
        while (sp > 0) {
            --sp;
            int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
            while (last - first > 16) {
                if (depth == 0) { stdsort_heapsort(a + first, last - first, less); break; }
                --depth;
                {
                    const int ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
                    int pick;
                    if (less(a[ia], a[ib])) {
                        if (less(a[ib], a[ic])) pick = ib;
                        else if (less(a[ia], a[ic])) pick = ic;
                        else pick = ia;
                    } else if (less(a[ia], a[ic])) pick = ia;
                    else if (less(a[ib], a[ic])) pick = ic;
                    else pick = ib;
                    const T t = a[first]; a[first] = a[pick]; a[pick] = t;
                }
                const T P = a[first];
                const int lo = first + 1, hi = last;
                int nL = 0, nR = 0;
                for (int i = lo; i < hi; i++) if (!less(a[i], P)) li[nL++] = i;
                for (int i = hi - 1; i >= lo; i--) if (!less(P, a[i])) ri[nR++] = i;
                ri[nR++] = first;
                int K = 0;
                const int kmax = nL < nR ? nL : nR;
                for (int k = 0; k < kmax; k++) K += li[k] < ri[k];
                for (int k = 0; k < K; k++) { const T t = a[li[k]]; a[li[k]] = a[ri[k]]; a[ri[k]] = t; }
                int cut;
                if (K < nL) cut = K > 0 ? (li[K] < ri[K - 1] ? li[K] : ri[K - 1]) : li[K];
                else cut = ri[K - 1];
                st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; sp++;
                last = cut;
            }
        }
    }
    for (int i = 0; i < n; i++) {
        int rank = 0;
        for (int j = 0; j < n; j++) rank += (less(a[j], a[i]) || (!less(a[i], a[j]) && j < i)) ? 1 : 0;
        tmp[rank] = a[i];
    }
    for (int i = 0; i < n; i++) a[i] = tmp[i];
}

#ifdef __CUDACC__
// ---- the same std::sort, cooperatively on one CTA -------------------------------------------------
// Result identical to libstdcxx_sort (element for element, ties included):
//  * __introsort_loop is run by warp 0 with the Hoare partition evaluated in parallel.  The
//    sequential scan "left pointer stops at the next element >= pivot, right pointer at the next element
//    <= pivot, swap, repeat until they cross" pairs the k-th left stopper l_k with the k-th right stopper
//    r_k of the ORIGINAL segment for every k with l_k < r_k (the pointers never re-visit a swapped
//    position before they cross), so the K swaps are independent and the cut is min(l_K, r_{K-1}).
//  * __final_insertion_sort is a stable insertion sort of the whole range, i.e. a stable sort by key:
//    every thread ranks its elements directly.
// a, tmp: n elements each in shared memory; li, ri: n + 1 ints each; stk: 192 ints.  All threads call.
template <typename T, typename Less>
__device__ inline void libstdcxx_sort_cta(T* a, T* tmp, int* li, int* ri, int* stk, int n, Less less)
{
    if (n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32 && n > 16) {
        int* st_first = stk; int* st_last = stk + 64; int* st_depth = stk + 128;
        int sp = 0, lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        if (lane == 0) { st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; }
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            --sp;
            int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
            __syncwarp();
            while (last - first > 16) {
                if (depth == 0) {
                    if (lane == 0) stdsort_heapsort(a + first, last - first, less);
                    __syncwarp();
                    break;
                }
                --depth;
                if (lane == 0) { // __move_median_to_first(first, first + 1, mid, last - 1)
                    const int ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
                    int pick;
                    if (less(a[ia], a[ib])) {
                        if (less(a[ib], a[ic])) pick = ib;
                        else if (less(a[ia], a[ic])) pick = ic;
                        else pick = ia;
                    } else if (less(a[ia], a[ic])) pick = ia;
                    else if (less(a[ib], a[ic])) pick = ic;
                    else pick = ib;
                    const T t = a[first]; a[first] = a[pick]; a[pick] = t;
                }
                __syncwarp();
                const T P = a[first];
                const int lo = first + 1, hi = last;
                int nL = 0, nR = 0;
                for (int c = 0; c < hi - lo; c += 32) {
                    const int il = lo + c + lane, ir = hi - 1 - c - lane;
                    const bool fl = il < hi && !less(a[il], P);
                    const bool fr = ir >= lo && !less(P, a[ir]);
                    const unsigned ml = __ballot_sync(0xffffffffu, fl), mr = __ballot_sync(0xffffffffu, fr);
                    const unsigned below = (1u << lane) - 1u;
                    if (fl) li[nL + __popc(ml & below)] = il;
                    if (fr) ri[nR + __popc(mr & below)] = ir;
                    nL += __popc(ml); nR += __popc(mr);
                }
                if (lane == 0) ri[nR] = first; // the pivot itself stops the right pointer
                nR++;
                __syncwarp();
                int K = 0;
                const int kmax = min(nL, nR);
                for (int c = 0; c < kmax; c += 32) {
                    const int k = c + lane;
                    const bool sw = k < kmax && li[k] < ri[k];
                    K += __popc(__ballot_sync(0xffffffffu, sw));
                }
                for (int k = lane; k < K; k += 32) {
                    const int x = li[k], y = ri[k];
                    const T t = a[x]; a[x] = a[y]; a[y] = t;
                }
                int cut;
                if (K < nL) cut = K > 0 ? min(li[K], ri[K - 1]) : li[K];
                else cut = ri[K - 1];
                __syncwarp();
                if (lane == 0) { st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; }
                sp++;
                last = cut;
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const T v = a[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const T w = a[j];
            rank += (less(w, v) || (!less(v, w) && j < i)) ? 1 : 0;
        }
        tmp[rank] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) a[i] = tmp[i];
    __syncthreads();
}
#endif

} // namespace dvm
