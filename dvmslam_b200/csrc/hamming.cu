// hamming.cu -- exhaustive nearest / second-nearest Hamming search over 256-bit ORB descriptors.
//
// The reference narrows inter-agent loop-closure matching with the DBoW2 vocabulary before it compares
// descriptors (SearchByBoW, O3/src/ORBmatcher.cc:709-834, fed by sendNewKeyFrameBows,
// src/slam_system/src/orb_slam3_wrapper.cpp:457-534).  On a B200 the exhaustive comparison of two
// keyframes (2000 x 2000 descriptors = 32 M popcounts) costs microseconds, so the exchange step
// (config C3) matches every received keyframe against every local one without a prefilter; the
// bookkeeping per query is the matchers' own bestDist1 / bestDist2 (strict '<': the first of equal
// distances wins).  DescriptorDistance = O3/src/ORBmatcher.cc:1900-1914.
//
// Bound: integer/popcount throughput, not HBM -- (na + nb) * 32 B are read once per tile pair while
// na * nb * 8 POPC are issued.  One thread owns kRows query descriptors in registers; the other side is
// streamed through shared memory in tiles that every thread reads as broadcasts.
#include "bow_kernels.cuh"

namespace dvm {

constexpr int kKnnThreads = 128;
constexpr int kKnnRows = 2;                       // query rows per thread
constexpr int kKnnTile = 256;                     // descriptors of b per shared-memory tile (8 KB)
constexpr unsigned kKnnNoKey = 256u << 20;

// popcount(a ^ b) over 256 bits.  POPC issues at a quarter of the ALU rate and bounds this kernel, so the eight
// XOR words are first compressed by a carry-save adder tree (bitwise full adders: 3-input XOR and majority are one
// LOP3 each) into four words of weight 1, 2, 4, 8: four POPC instead of eight, at the price of 14 more LOP3.
__device__ inline void csa(unsigned& sum, unsigned& carry, unsigned a, unsigned b, unsigned c)
{
    sum = a ^ b ^ c;
    carry = (a & b) | (c & (a ^ b));
}
__device__ inline unsigned dist256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    const unsigned x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const unsigned x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    unsigned s1, c1, s2, c2, s3, c3, s5, c5;
    csa(s1, c1, x0, x1, x2);
    csa(s2, c2, x3, x4, x5);
    csa(s3, c3, s1, s2, x6);
    const unsigned ones = s3 ^ x7, c4 = s3 & x7;
    csa(s5, c5, c1, c2, c3);
    const unsigned twos = s5 ^ c4, c6 = s5 & c4;
    const unsigned fours = c5 ^ c6, eights = c5 & c6;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}

// grid = (row groups of a block, splits of the b block, ba * bb)
__global__ void __launch_bounds__(kKnnThreads) hamming_bf_kernel(KnnArgs k, int pair0, int nsplit, int split_len, uint32_t* part1, uint32_t* part2)
{
    __shared__ uint4 tile[kKnnTile * 2];
    __shared__ int s_count;
    const int pair = pair0 + blockIdx.z, ia = pair / k.bb, ib = pair % k.bb;
    const uint4* A = reinterpret_cast<const uint4*>(k.a) + (size_t)ia * k.na * 2;
    const uint4* B = reinterpret_cast<const uint4*>(k.b) + (size_t)ib * k.nb * 2;
    const int row0 = (blockIdx.x * kKnnThreads + threadIdx.x) * kKnnRows;
    uint4 a0[kKnnRows], a1[kKnnRows];
    unsigned k1[kKnnRows], k2[kKnnRows];
#pragma unroll
    for (int r = 0; r < kKnnRows; r++) {
        const int row = min(row0 + r, k.na - 1);
        a0[r] = __ldg(A + (size_t)row * 2); a1[r] = __ldg(A + (size_t)row * 2 + 1);
        k1[r] = kKnnNoKey; k2[r] = kKnnNoKey;
    }
    const int j_begin = blockIdx.y * split_len, j_end = min(j_begin + split_len, k.nb);
    for (int j0 = j_begin; j0 < j_end; j0 += kKnnTile) {
        const int nt = min(kKnnTile, j_end - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < nt * 2; t += kKnnThreads) tile[t] = __ldg(B + (size_t)j0 * 2 + t);
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < nt; t++) {
            const uint4 b0 = tile[2 * t], b1 = tile[2 * t + 1];
#pragma unroll
            for (int r = 0; r < kKnnRows; r++) {
                const unsigned key = (dist256(a0[r], a1[r], b0, b1) << 20) | (unsigned)(j0 + t);
                k2[r] = min(k2[r], max(k1[r], key));   // second smallest key
                k1[r] = min(k1[r], key);
            }
        }
    }
    int accepted = 0;
#pragma unroll
    for (int r = 0; r < kKnnRows; r++) {
        const int row = row0 + r;
        if (row >= k.na) continue;
        const size_t o = ((size_t)pair * k.na + row);
        if (nsplit > 1) {
            part1[o * nsplit + blockIdx.y] = k1[r];
            part2[o * nsplit + blockIdx.y] = k2[r];
        } else {
            k.key1[o] = k1[r]; k.key2[o] = k2[r];
            const unsigned d1 = k1[r] >> 20, d2 = min(k2[r] >> 20, 256u);
            accepted += (d1 <= (unsigned)k.th_low && (float)d1 < __fmul_rn(k.nnratio, (float)d2));
        }
    }
    if (k.counts && nsplit == 1) {
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        accepted = __reduce_add_sync(0xffffffffu, accepted);
        if ((threadIdx.x & 31) == 0 && accepted) atomicAdd(&s_count, accepted);
        __syncthreads();
        if (threadIdx.x == 0 && s_count) atomicAdd(&k.counts[pair], s_count);
    }
}

// merges the per-split nearest / second-nearest keys of a row
__global__ void __launch_bounds__(256) hamming_merge_kernel(KnnArgs k, int nsplit, const uint32_t* part1, const uint32_t* part2)
{
    const size_t o = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t total = (size_t)k.ba * k.bb * k.na;
    if (o >= total) return;
    unsigned k1 = kKnnNoKey, k2 = kKnnNoKey;
    for (int s = 0; s < nsplit; s++) {
        const unsigned p1 = part1[o * nsplit + s], p2 = part2[o * nsplit + s];
        k2 = min(min(k2, p2), max(k1, p1));
        k1 = min(k1, p1);
    }
    k.key1[o] = k1; k.key2[o] = k2;
    if (k.counts) {
        const unsigned d1 = k1 >> 20, d2 = min(k2 >> 20, 256u);
        if (d1 <= (unsigned)k.th_low && (float)d1 < __fmul_rn(k.nnratio, (float)d2)) atomicAdd(&k.counts[o / k.na], 1);
    }
}

int launch_hamming_knn(const KnnArgs& k, KnnScratch& sc, cudaStream_t stream, int mode)
{
    const int pairs = k.ba * k.bb;
    if (pairs <= 0 || k.na <= 0) return DVM_OK;
    if (mode >= 2 || (mode == 0 && hamming_tc_applicable(k))) {
        if (k.na < 1 || k.nb < 1) { set_error("the tensor-core Hamming path needs non-empty blocks"); return DVM_ERR_INVALID; }
        return launch_hamming_knn_tc(k, sc, stream, mode == 3 ? 0 : mode == 4 ? 2 : 1);
    }
    if (k.counts) DVM_CUDA(cudaMemsetAsync(k.counts, 0, (size_t)pairs * sizeof(int), stream));
    const int groups = div_up(k.na, kKnnThreads * kKnnRows);
    // enough CTAs to fill the 148 SMs several times over: split the b block when the batch is small
    int nsplit = 1;
    if (k.nb > kKnnTile) {
        const long ctas = (long)groups * pairs;
        nsplit = (int)((4L * kNumSMs + ctas - 1) / ctas);
        nsplit = max(1, min(nsplit, div_up(k.nb, kKnnTile)));
    }
    const int split_len = div_up(div_up(max(k.nb, 1), nsplit), kKnnTile) * kKnnTile;
    nsplit = max(1, div_up(k.nb, split_len));
    uint32_t *p1 = nullptr, *p2 = nullptr;
    if (nsplit > 1) {
        const size_t need = (size_t)pairs * k.na * nsplit;
        if (need > sc.cap) {
            DVM_CUDA(cudaStreamSynchronize(stream));
            cudaFree(sc.part[0]); cudaFree(sc.part[1]);
            sc.part[0] = sc.part[1] = nullptr; sc.cap = 0;
            DVM_CUDA(cudaMalloc(&sc.part[0], need * sizeof(uint32_t)));
            DVM_CUDA(cudaMalloc(&sc.part[1], need * sizeof(uint32_t)));
            sc.cap = need;
        }
        p1 = sc.part[0]; p2 = sc.part[1];
    }
    for (int pair0 = 0; pair0 < pairs; pair0 += 65535) // gridDim.z limit
        DVM_LAUNCH(hamming_bf_kernel, dim3(groups, nsplit, min(65535, pairs - pair0)), kKnnThreads, 0, stream, k, pair0,
                   nsplit, split_len, p1, p2);
    if (nsplit > 1) {
        const size_t total = (size_t)pairs * k.na;
        DVM_LAUNCH(hamming_merge_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, k, nsplit, p1, p2);
    }
    DVM_CUDA(cudaGetLastError());
    return DVM_OK;
}

} // namespace dvm
