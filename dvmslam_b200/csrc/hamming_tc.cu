// hamming_tc.cu -- the exhaustive Hamming search of config C3 on the 5th-generation tensor cores.
//
// popcount(a ^ b) over 256 bits is a +-1 dot product: with every bit mapped to +1 / -1,  a . b = 256 - 2 * dist.
// So the [na x 256] x [256 x nb] comparison of two keyframes is an int8 GEMM with exact int32 accumulation, and
// the matchers' bookkeeping (bestDist1 / bestDist2 with strict '<', O3/src/ORBmatcher.cc:752-768; DescriptorDistance
// :1900-1914) is an epilogue over the accumulator rows.  Same outputs, bit for bit, as hamming_bf_kernel (hamming.cu),
// which stays the path for small problems.
//
//   expand_desc_kernel   bits -> int8 (+1 / -1), once per call, [blocks][n][256] in HBM (L2-resident: 64 MB at C3)
//   hamming_tc_kernel    persistent, one CTA per SM, warp-specialised:
//       warp 0  TMA producer of the B tiles   (128 descriptors x 256 B, SWIZZLE_128B, 3 stages)
//       warp 1  MMA issuer: tcgen05.mma.cta_group::1.kind::i8, M 128 x N 128 x K 32, 8 per accumulator
//       warp 2  TMA producer of the A block   (256 descriptors, stationary over the 16 column tiles, double-buffered)
//       warp 3  TMEM allocation (512 columns = 4 accumulators of 128 x 128 int32)
//       warps 4..11  epilogue: tcgen05.ld 32 columns at a time, key = acc * -2^19 + column (one IMAD), nearest /
//                    second-nearest keys by three integer min / max, per-row running keys in registers
//   A work item is (keyframe pair, 256-row block of a); its 16 column tiles of 128 pipeline through the four TMEM
//   accumulators (two row halves x two tiles in flight), so the epilogue of tile t overlaps the MMAs of tile t + 1.
//
// Roofline: tensor pipe.  2 * na * nb * 256 integer operations per keyframe pair (C3: 64 x 64 pairs x 2000^2 -> 8.4e15);
// the epilogue issues 4 instructions per accumulator element, which is the same order as the MMA time
// (128 x 128 x 256 MACs = 512 cycles per accumulator at 8192 MAC/clk/SM; 128 x 128 x 4 / 128 lanes = 512 issue slots
// per SM sub-partition) -- the two pipelines are balanced by construction, and the kernel is bound by whichever is slower.
#include "bow_kernels.cuh"
#include <cuda.h>

namespace dvm {

namespace {

constexpr int kTcThreads = 384;
constexpr int kTcRows = 256;          // rows of a per work item (two MMA row halves)
constexpr int kTcCols = 128;          // columns (descriptors of b) per tile
constexpr int kTcBStages = 3;
constexpr uint32_t kTileBytes = 128 * 128;                 // one [128 rows][128 B] swizzle-128B operand tile
constexpr uint32_t kABytes = 4 * kTileBytes;               // 2 row halves x 2 K chunks
constexpr uint32_t kBBytes = 2 * kTileBytes;               // 2 K chunks
constexpr size_t kTcSmemBytes = 1024 /*alignment slack*/ + 2 * kABytes + kTcBStages * kBBytes + 256 /*barriers*/;
constexpr unsigned kNoKey = 256u << 20;

// instruction descriptor of tcgen05.mma.kind::i8 (cute/arch/mma_sm100_desc.hpp, UMMA::InstrDescriptor):
//   [4,6) D format 2 = S32 | [7,10) A format 1 = signed 8 bit | [10,13) B format 1 | [15] A K-major | [16] B K-major
//   | [17,23) N >> 3 | [24,29) M >> 4
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((kTcCols >> 3) << 17) | ((128u >> 4) << 24);

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (UMMA::SmemDescriptor): start address >> 4, leading byte offset
// (unused by swizzled K-major layouts; 1 as CUTLASS encodes it), stride byte offset = 8 rows x 128 B, version 1, layout 2
__device__ inline uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ inline void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Spin on a phase parity; a wait that does not complete within ~2 s traps, so that a protocol bug surfaces as a launch
// failure instead of a hung GPU.
__device__ inline void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 0xfff) == 0xfff) {
            const long long t = clock64();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ inline void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ inline void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ inline void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ inline void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
// mbarrier arrive when every MMA issued so far by this thread has completed (implies tcgen05.fence::before_thread_sync)
__device__ inline void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ inline void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ inline void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ inline void tmem_ld32(uint32_t taddr, int (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// tcgen05.ld is asynchronous: the destination registers are valid only after tcgen05.wait::ld.  The wait takes them as
// in/out operands so that the compiler cannot schedule their uses (or copies) above it.
__device__ inline void tmem_ld_wait(int (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

struct TcArgs {
    int ba, na, bb, nb;
    uint32_t* key1; uint32_t* key2; int* counts;
    int th_low; float nnratio;
    int mblocks, ntiles, items;
    int epilogue;   // 0: skip the accumulator read-out (MMA-only timing probe: no outputs)
};

// One 32-column chunk into the tile-local nearest / second-nearest keys (relative key = acc * -2^19 + column in tile).
// The running pair (k1 <= k2) is a dependency chain, and an epilogue warp has only one partner on its scheduler to hide
// it, so the chunk feeds kChains independent pairs (merged once per tile), and two columns enter a pair at a time:
//   lo/hi = min/max(a, b);  k2' = min3(max(k1, lo), k2, hi);  k1' = min(k1, lo)      -- 7 instructions per two columns
// (the second smallest of two sorted pairs is min(max of the firsts, min of the seconds)).
constexpr int kChains = 4;
__device__ inline int min3i(int a, int b, int c) { return min(min(a, b), c); }   // VIMNMX3
template <bool kMasked>
__device__ inline void fold_chunk(const int (&v)[32], int col0, int valid, int negmul, int (&k1)[kChains], int (&k2)[kChains])
{
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
        const int c = (i >> 1) % kChains;
        int a = v[i] * negmul + (col0 + i);
        int b = v[i + 1] * negmul + (col0 + i + 1);
        if (kMasked) {
            if (col0 + i >= valid) a = 0x7fffffff;
            if (col0 + i + 1 >= valid) b = 0x7fffffff;
        }
        const int lo = min(a, b), hi = max(a, b);
        k2[c] = min3i(max(k1[c], lo), k2[c], hi);
        k1[c] = min(k1[c], lo);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcArgs g)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem_base;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle-128B tiles need 1024-byte alignment
    const uint32_t sA = base, sB = base + 2 * kABytes, sBar = sB + kTcBStages * kBBytes;
    // barriers (8 bytes each)
    const uint32_t a_full = sBar, a_empty = sBar + 16, b_full = sBar + 32, b_empty = sBar + 64, acc_full = sBar + 96,
                   acc_empty = sBar + 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < kTcBStages; i++) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
        for (int i = 0; i < 4; i++) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    if (warp == 0) {
        // ---------------- B tiles ----------------
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
                const int pair = item / g.mblocks, ib = pair % g.bb;
                for (int t = 0; t < g.ntiles; t++) {
                    mbar_wait(b_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(b_full + 8 * stage, kBBytes);
                    const uint32_t dst = sB + stage * kBBytes;
                    tma_load_3d(dst, &map_b, 0, t * kTcCols, ib, b_full + 8 * stage);
                    tma_load_3d(dst + kTileBytes, &map_b, 128, t * kTcCols, ib, b_full + 8 * stage);
                    if (++stage == kTcBStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ---------------- A blocks ----------------
        if (lane == 0) {
            uint32_t buf = 0, phase = 0;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
                const int pair = item / g.mblocks, mb = item % g.mblocks, ia = pair / g.bb;
                mbar_wait(a_empty + 8 * buf, phase ^ 1);
                mbar_expect_tx(a_full + 8 * buf, kABytes);
                const uint32_t dst = sA + buf * kABytes;
                for (int h = 0; h < 2; h++)
                    for (int c = 0; c < 2; c++)
                        tma_load_3d(dst + (h * 2 + c) * kTileBytes, &map_a, c * 128, mb * kTcRows + h * 128, ia, a_full + 8 * buf);
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            uint32_t stage = 0, bphase = 0, buf = 0, aphase = 0, tc = 0;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
                mbar_wait(a_full + 8 * buf, aphase);
                for (int t = 0; t < g.ntiles; t++, tc++) {
                    mbar_wait(b_full + 8 * stage, bphase);
                    for (int h = 0; h < 2; h++) {
                        const uint32_t slot = (tc & 1) * 2 + h;
                        mbar_wait(acc_empty + 8 * slot, ((tc >> 1) & 1) ^ 1);
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const uint64_t ad = umma_desc(sA + buf * kABytes + (h * 2 + c) * kTileBytes);
                            const uint64_t bd = umma_desc(sB + stage * kBBytes + c * kTileBytes);
#pragma unroll
                            for (int j = 0; j < 4; j++)   // K = 32 bytes per MMA: two 16-byte units along the swizzled row
                                umma_i8(tmem + slot * kTcCols, ad + 2 * j, bd + 2 * j, (c | j) != 0);
                        }
                        umma_commit(acc_full + 8 * slot);
                    }
                    umma_commit(b_empty + 8 * stage);
                    if (++stage == kTcBStages) { stage = 0; bphase ^= 1; }
                }
                umma_commit(a_empty + 8 * buf);
                if (++buf == 2) { buf = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue ----------------
        const int half = (warp - 4) >> 2, quarter = warp & 3;   // a warp reads the TMEM lanes 32 * (warp % 4) ..
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int negmul = -(1 << 19);
        uint32_t tc = 0;
        for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
            const int pair = item / g.mblocks, mb = item % g.mblocks;
            const int row = mb * kTcRows + half * 128 + quarter * 32 + lane;
            unsigned k1 = kNoKey, k2 = kNoKey;
            for (int t = 0; t < g.ntiles; t++, tc++) {
                const uint32_t slot = (tc & 1) * 2 + half;
                mbar_wait(acc_full + 8 * slot, (tc >> 1) & 1);
                tc_fence_after();
                if (g.epilogue == 2) {   // timing probe: TMEM read-out only
                    const uint32_t taddr = lane_addr + slot * kTcCols;
                    int va[32], vb[32];
                    tmem_ld32(taddr, va);
                    tmem_ld32(taddr + 32, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    int x = va[0] ^ vb[31];
                    tmem_ld32(taddr + 64, va);
                    tmem_ld32(taddr + 96, vb);
                    tmem_ld_wait(va); tmem_ld_wait(vb);
                    x ^= va[5] ^ vb[7];
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * slot);
                    if (x == 0x12345678) k1 = x;
                } else if (g.epilogue) {
                    const int valid = min(kTcCols, g.nb - t * kTcCols);
                    const uint32_t taddr = lane_addr + slot * kTcCols;
                    int va[32], vb[32];
                    int t1[kChains], t2[kChains];
#pragma unroll
                    for (int c = 0; c < kChains; c++) { t1[c] = 0x7fffffff; t2[c] = 0x7fffffff; }
                    tmem_ld32(taddr, va);
                    tmem_ld_wait(va);
                    tmem_ld32(taddr + 32, vb);
                    if (valid == kTcCols) fold_chunk<false>(va, 0, valid, negmul, t1, t2); else fold_chunk<true>(va, 0, valid, negmul, t1, t2);
                    tmem_ld_wait(vb);
                    tmem_ld32(taddr + 64, va);
                    if (valid == kTcCols) fold_chunk<false>(vb, 32, valid, negmul, t1, t2); else fold_chunk<true>(vb, 32, valid, negmul, t1, t2);
                    tmem_ld_wait(va);
                    tmem_ld32(taddr + 96, vb);
                    if (valid == kTcCols) fold_chunk<false>(va, 64, valid, negmul, t1, t2); else fold_chunk<true>(va, 64, valid, negmul, t1, t2);
                    tmem_ld_wait(vb);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * slot);   // the accumulator is in registers: the MMA warp may reuse it
                    if (valid == kTcCols) fold_chunk<false>(vb, 96, valid, negmul, t1, t2); else fold_chunk<true>(vb, 96, valid, negmul, t1, t2);
                    // tile-local relative keys -> absolute keys (distance << 20 | column); distance 256 never wins
                    const unsigned off = (256u << 19) + (unsigned)(t * kTcCols);
#pragma unroll
                    for (int c = 1; c < kChains; c++) {   // merge the chains: two sorted pairs -> the two smallest
                        t2[0] = min(max(t1[0], t1[c]), min(t2[0], t2[c]));
                        t1[0] = min(t1[0], t1[c]);
                    }
                    const unsigned a1 = t1[0] == 0x7fffffff ? kNoKey : min((unsigned)t1[0] + off, kNoKey);
                    const unsigned a2 = t2[0] == 0x7fffffff ? kNoKey : min((unsigned)t2[0] + off, kNoKey);
                    k2 = min(min(k2, a2), max(k1, a1));
                    k1 = min(k1, a1);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8 * slot);
                }
            }
            if (g.epilogue == 1) {
                int accepted = 0;
                if (row < g.na) {
                    const size_t o = (size_t)pair * g.na + row;
                    g.key1[o] = k1; g.key2[o] = k2;
                    const unsigned d1 = k1 >> 20, d2 = min(k2 >> 20, 256u);
                    accepted = (d1 <= (unsigned)g.th_low && (float)d1 < __fmul_rn(g.nnratio, (float)d2));
                }
                if (g.counts) {
                    accepted = __reduce_add_sync(0xffffffffu, accepted);
                    if (lane == 0 && accepted) atomicAdd(&g.counts[pair], accepted);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// descriptor bits -> +1 / -1 bytes: thread = one 32-bit word of a descriptor -> 32 bytes
__global__ void __launch_bounds__(256) expand_desc_kernel(const uint32_t* __restrict__ src, uint4* __restrict__ dst, size_t nwords)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= nwords) return;
    const uint32_t w = __ldg(src + i);
    uint32_t out[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint32_t nib = (w >> (4 * k)) & 15u;
        const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;   // bit b of the nibble -> byte b (0 / 1)
        out[k] = ~(spread * 0xfeu);                                   // 1 -> 0x01 (+1), 0 -> 0xff (-1)
    }
    dst[2 * i] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[2 * i + 1] = make_uint4(out[4], out[5], out[6], out[7]);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool encode_desc_map(CUtensorMap* map, const uint8_t* base, int n, int blocks)
{
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeFn)fn;
    }();
    if (!encode) return false;
    // [blocks][n][256] int8, box = 128 rows x 128 bytes (one swizzle-128B atom column); rows past n read as zero
    const cuuint64_t dims[3] = { 256, (cuuint64_t)n, (cuuint64_t)blocks };
    const cuuint64_t strides[2] = { 256, (cuuint64_t)n * 256 };
    const cuuint32_t box[3] = { 128, 128, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace

bool hamming_tc_applicable(const KnnArgs& k)
{
    // worth the expansion pass and a persistent grid: at least a few hundred work items' worth of comparisons
    return k.na >= 128 && k.nb >= 128 && k.nb <= (1 << 19) && (double)k.ba * k.bb * k.na * k.nb >= 64.0 * 1024 * 1024;
}

int launch_hamming_knn_tc(const KnnArgs& k, KnnScratch& sc, cudaStream_t stream, int epilogue)
{
    const size_t abytes = (size_t)k.ba * k.na * 256, bbytes = (size_t)k.bb * k.nb * 256;
    const size_t need = ((abytes + 1023) & ~(size_t)1023) + bbytes + 1024;
    if (need > sc.exp_cap) {
        DVM_CUDA(cudaStreamSynchronize(stream));
        cudaFree(sc.expanded); sc.expanded = nullptr; sc.exp_cap = 0;
        DVM_CUDA(cudaMalloc(&sc.expanded, need));
        sc.exp_cap = need;
    }
    uint8_t* ea = sc.expanded;
    uint8_t* eb = sc.expanded + ((abytes + 1023) & ~(size_t)1023);
    int dev = 0, sms = kNumSMs;
    DVM_CUDA(cudaGetDevice(&dev));
    static std::atomic<unsigned long long> attr_set{ 0 };   // function attributes are per device
    if (dev < 64 && !(attr_set.load() >> dev & 1ull)) {
        DVM_CUDA(cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
        attr_set.fetch_or(1ull << dev);
    }
    CUtensorMap map_a, map_b;
    if (!encode_desc_map(&map_a, ea, k.na, k.ba) || !encode_desc_map(&map_b, eb, k.nb, k.bb)) {
        set_error("cuTensorMapEncodeTiled failed for the expanded descriptor arrays");
        return DVM_ERR_CUDA;
    }
    const size_t wa = (size_t)k.ba * k.na * 8, wb = (size_t)k.bb * k.nb * 8;
    DVM_LAUNCH(expand_desc_kernel, (unsigned)((wa + 255) / 256), 256, 0, stream, reinterpret_cast<const uint32_t*>(k.a),
               reinterpret_cast<uint4*>(ea), wa);
    DVM_LAUNCH(expand_desc_kernel, (unsigned)((wb + 255) / 256), 256, 0, stream, reinterpret_cast<const uint32_t*>(k.b),
               reinterpret_cast<uint4*>(eb), wb);
    if (k.counts) DVM_CUDA(cudaMemsetAsync(k.counts, 0, (size_t)k.ba * k.bb * sizeof(int), stream));
    TcArgs g;
    g.ba = k.ba; g.na = k.na; g.bb = k.bb; g.nb = k.nb;
    g.key1 = k.key1; g.key2 = k.key2; g.counts = k.counts; g.th_low = k.th_low; g.nnratio = k.nnratio;
    g.mblocks = div_up(k.na, kTcRows);
    g.ntiles = div_up(k.nb, kTcCols);
    g.items = k.ba * k.bb * g.mblocks;
    g.epilogue = epilogue;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = std::min(g.items, sms);
    DVM_LAUNCH(hamming_tc_kernel, grid, kTcThreads, kTcSmemBytes, stream, map_a, map_b, g);
    DVM_CUDA(cudaGetLastError());
    return DVM_OK;
}

} // namespace dvm
