// orb_kernels.cu -- sm_100a kernels of the ORB extractor (reference: O3/src/ORBextractor.cc).
//
//   pyramid_kernel         cv::resize INTER_LINEAR chain, all levels in one launch (ComputePyramid, :957-976);
//                          resize_linear_kernel = one level (debug / fallback)
//   fast_cells_kernel      per-cell FAST-9-16 + 3x3 NMS + 20->7 threshold fallback, one CTA per
//                          ~35 px cell                           (ComputeKeyPointsOctTree, :612-692)
//   octree_kernel          DistributeOctTree, one CTA per level; the last CTA also lays out the
//                          reference's output order              (:419-610, :929-950)
//   describe_kernel        IC_Angle + 7x7 Gaussian + rBRIEF, one warp per keypoint, fused on a
//                          43x43 window                          (:75-143, :919-925)
#include "orb_kernels.cuh"
#include "orb_math.cuh"
#include <mutex>

namespace dvm {

// ------------------------------------------------------------------------------------------ resize
__global__ void __launch_bounds__(256) resize_linear_kernel(const uint8_t* __restrict__ src, int spitch,
                                                            uint8_t* __restrict__ dst, int dw, int dh, int dpitch,
                                                            const ResizeX* __restrict__ xt,
                                                            const ResizeY* __restrict__ yt)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const ResizeX X = xt[dx];
    const ResizeY Y = yt[dy];
    const uint8_t* S0 = src + (size_t)Y.sy0 * spitch;
    const uint8_t* S1 = src + (size_t)Y.sy1 * spitch;
    const int R0 = S0[X.sx0] * X.a0 + S0[X.sx1] * X.a1;
    const int R1 = S1[X.sx0] * X.a0 + S1[X.sx1] * X.a1;
    int v = (((Y.b0 * (R0 >> 4)) >> 16) + ((Y.b1 * (R1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    dst[(size_t)dy * dpitch + dx] = (uint8_t)v;
}

void launch_resize_level(const OrbCfg& cfg, const OrbBuffers& b, int level, uint8_t* dst, cudaStream_t stream)
{
    const OrbLevel& S = cfg.lv[level - 1];
    const OrbLevel& D = cfg.lv[level];
    dim3 block(32, 8), grid(div_up(D.w, 32), div_up(D.h, 8));
    DVM_LAUNCH(resize_linear_kernel, grid, block, 0, stream, S.img, S.pitch, dst, D.w, D.h, D.pitch,
               b.xtab + D.xtab_off, b.ytab + D.ytab_off);
}

// ---- the whole chain in one launch ----
struct PyrLevels {
    const uint8_t* src0; int pitch0;          // level 0
    uint8_t* dst[kMaxLevels]; int pitch[kMaxLevels];
    int xoff[kMaxLevels], yoff[kMaxLevels];   // resize tables of level l
    int x4off[kMaxLevels];                    // ResizeX4 table of level l
    int w0;                                   // width of level 0
};

constexpr int kPyrThreads = 512;   // two CTAs per SM
constexpr int kPyrMaxSpan = 192;   // widest / tallest region of one tile at level 1 (tile 32 x 16 at level 7 of a 1.2 pyramid: 108 x 60)

__global__ void __launch_bounds__(kPyrThreads, 2) pyramid_kernel(OrbPyrPlan plan, PyrLevels lv, int nlevels, const ResizeX* __restrict__ xtab,
                                                              const ResizeY* __restrict__ ytab, const int* __restrict__ colr,
                                                              const int* __restrict__ rowr)
{
    extern __shared__ uint8_t pyr_smem[];
    __shared__ int s_rng[kMaxLevels][6];
    __shared__ ResizeX s_xt[kPyrMaxSpan];
    __shared__ ResizeY s_yt[kPyrMaxSpan];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < nlevels * 6) {
        const int l = threadIdx.x / 6, k = threadIdx.x % 6;
        s_rng[l][k] = k < 3 ? colr[(blockIdx.x * kMaxLevels + l) * 3 + k] : rowr[(blockIdx.y * kMaxLevels + l) * 3 + k - 3];
    }
    __syncthreads();
    for (int l = 1; l < nlevels; l++) {
        const int x0 = s_rng[l][0], ox1 = s_rng[l][1], nx1 = s_rng[l][2], y0 = s_rng[l][3], oy1 = s_rng[l][4], ny1 = s_rng[l][5];
        const int px0 = s_rng[l - 1][0], py0 = s_rng[l - 1][3];
        // this level's table slices: one dependent load less per pixel
        for (int i = threadIdx.x; i <= nx1 - x0; i += kPyrThreads) s_xt[i] = xtab[lv.xoff[l] + x0 + i];
        for (int i = threadIdx.x; i <= ny1 - y0; i += kPyrThreads) s_yt[i] = ytab[lv.yoff[l] + y0 + i];
        __syncthreads();
        const uint8_t* src = l == 1 ? lv.src0 : pyr_smem + plan.soff[l - 1];
        const int sp = l == 1 ? lv.pitch0 : plan.spitch[l - 1];
        const int sx_org = l == 1 ? 0 : px0, sy_org = l == 1 ? 0 : py0;
        uint8_t* buf = pyr_smem + plan.soff[l];
        const int bp = plan.spitch[l];
        uint8_t* dst = lv.dst[l];
        const int dp = lv.pitch[l];
        const int rw = nx1 - x0 + 1;
        for (int ly = ty; ly <= ny1 - y0; ly += kPyrThreads / 32) {
            const ResizeY Y = s_yt[ly];
            const uint8_t* S0 = src + ((Y.sy0 - sy_org) * sp - sx_org);   // (32-bit offsets: a level is far below 2 GB)
            const uint8_t* S1 = src + ((Y.sy1 - sy_org) * sp - sx_org);
            const int dy = y0 + ly;
            uint8_t* drow = dst + dy * dp + x0;
            // the arithmetic of resize_linear_kernel (cv::resize, 8-bit fixed point, 11-bit coefficients), two pixels per
            // iteration so that their loads are in flight together
            auto pixel = [&](const ResizeX& X) {
                const int R0 = S0[X.sx0] * X.a0 + S0[X.sx1] * X.a1;
                const int R1 = S1[X.sx0] * X.a0 + S1[X.sx1] * X.a1;
                const int v = (((Y.b0 * (R0 >> 4)) >> 16) + ((Y.b1 * (R1 >> 4)) >> 16) + 2) >> 2;
                return (uint8_t)min(max(v, 0), 255);
            };
            for (int lx = tx; lx < rw; lx += 64) {
                const bool second = lx + 32 < rw;
                const ResizeX Xa = s_xt[lx], Xb = s_xt[second ? lx + 32 : lx];
                const uint8_t va = pixel(Xa), vb = pixel(Xb);
                buf[ly * bp + lx] = va;
                if (x0 + lx <= ox1 && dy <= oy1) drow[lx] = va;
                if (second) {
                    buf[ly * bp + lx + 32] = vb;
                    if (x0 + lx + 32 <= ox1 && dy <= oy1) drow[lx + 32] = vb;
                }
            }
        }
        __syncthreads();
    }
}

// The same chain with FOUR destination pixels per thread.  A thread reads the (at most) 12 source bytes its four pixels tap
// as three aligned words per source row, picks the tap bytes of two pixels at a time with one PRMT (selectors from the host-built
// ResizeX4 table), forms tap0 * a0 + tap1 * a1 with one IDP.2A per pixel and row, and stores its four results as one word to the
// level's shared-memory region and -- for the columns this tile owns -- to the level image.  ~20 instructions per pixel against
// ~50 for the byte-gather form above, same arithmetic (cv::resize, 8-bit fixed point), same plan (regions start on multiples
// of 4).  Items = (row, group) pairs dealt over all 512 threads, so the narrow regions of the upper levels still fill the warps.
constexpr int kPyrMaxGroups = kPyrMaxSpan / 4;

__global__ void __launch_bounds__(kPyrThreads, 2) pyramid4_kernel(OrbPyrPlan plan, PyrLevels lv, int nlevels, const ResizeX4* __restrict__ x4tab,
                                                               const ResizeY* __restrict__ ytab, const int* __restrict__ colr,
                                                               const int* __restrict__ rowr)
{
    extern __shared__ __align__(16) uint8_t pyr_smem[];
    __shared__ int s_rng[kMaxLevels][6];
    __shared__ __align__(16) uint4 s_x4[2 * kPyrMaxGroups];
    __shared__ ResizeY s_yt[kPyrMaxSpan];
    if (threadIdx.x < nlevels * 6) {
        const int l = threadIdx.x / 6, k = threadIdx.x % 6;
        s_rng[l][k] = k < 3 ? colr[(blockIdx.x * kMaxLevels + l) * 3 + k] : rowr[(blockIdx.y * kMaxLevels + l) * 3 + k - 3];
    }
    __syncthreads();
    for (int l = 1; l < nlevels; l++) {
        const int x0 = s_rng[l][0], ox1 = s_rng[l][1], nx1 = s_rng[l][2], y0 = s_rng[l][3], oy1 = s_rng[l][4], ny1 = s_rng[l][5];
        const int px0 = s_rng[l - 1][0], py0 = s_rng[l - 1][3];
        const int ng = (nx1 - x0 + 4) >> 2, nrows = ny1 - y0 + 1;
        {
            const uint4* src = reinterpret_cast<const uint4*>(x4tab + lv.x4off[l] + (x0 >> 2));
            for (int i = threadIdx.x; i < 2 * ng; i += kPyrThreads) s_x4[i] = src[i];
            for (int i = threadIdx.x; i < nrows; i += kPyrThreads) s_yt[i] = ytab[lv.yoff[l] + y0 + i];
        }
        __syncthreads();
        const bool from_global = l == 1;
        const int sp = from_global ? lv.pitch0 : plan.spitch[l - 1];
        const int sx_org_w = from_global ? 0 : px0 >> 2, sy_org = from_global ? 0 : py0;
        const int wlast = from_global ? (lv.w0 - 1) >> 2 : (plan.spitch[l - 1] >> 2) - 1;
        const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(lv.src0);
        const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(pyr_smem + plan.soff[l - 1]);
        uint8_t* buf = pyr_smem + plan.soff[l];
        const int bp = plan.spitch[l];
        uint8_t* dst = lv.dst[l];
        const int dp = lv.pitch[l];
        const float inv_ng = 1.0f / (float)ng;
        for (int idx = threadIdx.x; idx < nrows * ng; idx += kPyrThreads) {
            int ly = (int)((float)idx * inv_ng);
            int g = idx - ly * ng;
            if (g < 0) { ly--; g += ng; } else if (g >= ng) { ly++; g -= ng; }
            const ResizeY Y = s_yt[ly];
            const uint4 T0 = s_x4[2 * g], T1 = s_x4[2 * g + 1];   // {sel, base, c0, c1}, {c2, c3, -, -}
            const int wb = (int)(T0.y & 0xffffu) - sx_org_w;
            const int i0 = min(wb, wlast), i1 = min(wb + 1, wlast), i2 = min(wb + 2, wlast);
            const int r0 = ((Y.sy0 - sy_org) * sp) >> 2, r1 = ((Y.sy1 - sy_org) * sp) >> 2;
            uint32_t a0, a1, a2, b0, b1, b2;
            if (from_global) {
                a0 = __ldg(gsrc + r0 + i0); a1 = __ldg(gsrc + r0 + i1); a2 = __ldg(gsrc + r0 + i2);
                b0 = __ldg(gsrc + r1 + i0); b1 = __ldg(gsrc + r1 + i1); b2 = __ldg(gsrc + r1 + i2);
            } else {
                a0 = ssrc[r0 + i0]; a1 = ssrc[r0 + i1]; a2 = ssrc[r0 + i2];
                b0 = ssrc[r1 + i0]; b1 = ssrc[r1 + i1]; b2 = ssrc[r1 + i2];
            }
            const bool winA = (T0.y >> 16) & 1u, winB = (T0.y >> 17) & 1u;
            const uint32_t selA = T0.x & 0xffffu, selB = T0.x >> 16;
            const uint32_t tA0 = __byte_perm(winA ? a1 : a0, winA ? a2 : a1, selA), tB0 = __byte_perm(winB ? a1 : a0, winB ? a2 : a1, selB);
            const uint32_t tA1 = __byte_perm(winA ? b1 : b0, winA ? b2 : b1, selA), tB1 = __byte_perm(winB ? b1 : b0, winB ? b2 : b1, selB);
            const int R00 = (int)__dp2a_lo(T0.z, tA0, 0u), R01 = (int)__dp2a_hi(T0.w, tA0, 0u);
            const int R02 = (int)__dp2a_lo(T1.x, tB0, 0u), R03 = (int)__dp2a_hi(T1.y, tB0, 0u);
            const int R10 = (int)__dp2a_lo(T0.z, tA1, 0u), R11 = (int)__dp2a_hi(T0.w, tA1, 0u);
            const int R12 = (int)__dp2a_lo(T1.x, tB1, 0u), R13 = (int)__dp2a_hi(T1.y, tB1, 0u);
            auto vert = [&](int R0, int R1) {
                const int v = (((Y.b0 * (R0 >> 4)) >> 16) + ((Y.b1 * (R1 >> 4)) >> 16) + 2) >> 2;
                return (uint32_t)min(max(v, 0), 255);
            };
            const uint32_t packed = vert(R00, R10) | vert(R01, R11) << 8 | vert(R02, R12) << 16 | vert(R03, R13) << 24;
            *reinterpret_cast<uint32_t*>(buf + ly * bp + 4 * g) = packed;
            const int dy = y0 + ly;
            if (x0 + 4 * g <= ox1 && dy <= oy1) *reinterpret_cast<uint32_t*>(dst + dy * dp + x0 + 4 * g) = packed;
        }
        __syncthreads();
    }
}

void prepare_pyramid_kernel(int smem_bytes)
{
    cudaFuncSetAttribute(pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaFuncSetAttribute(pyramid4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

void launch_pyramid(const OrbCfg& cfg, const OrbBuffers& b, cudaStream_t stream)
{
    if (cfg.nlevels < 2) return;
    PyrLevels lv;
    memset(&lv, 0, sizeof(lv));
    lv.src0 = cfg.lv[0].img; lv.pitch0 = cfg.lv[0].pitch;
    for (int l = 1; l < cfg.nlevels; l++) {
        lv.dst[l] = const_cast<uint8_t*>(cfg.lv[l].img); lv.pitch[l] = cfg.lv[l].pitch;
        lv.xoff[l] = cfg.lv[l].xtab_off; lv.yoff[l] = cfg.lv[l].ytab_off;
        lv.x4off[l] = cfg.lv[l].x4_off;
    }
    lv.w0 = cfg.lv[0].w;
    // four pixels per thread when the plan allows it and level 0 (possibly the caller's image, read in place) is word-aligned
    static const bool force_scalar = getenv("DVM_PYRAMID_SCALAR") != nullptr;   // (diagnostics)
    if (cfg.pyr.vec_ok && !force_scalar && ((uintptr_t)lv.src0 & 3) == 0 && (lv.pitch0 & 3) == 0) {
        DVM_LAUNCH(pyramid4_kernel, dim3(cfg.pyr.ntx, cfg.pyr.nty), kPyrThreads, cfg.pyr.smem_bytes, stream, cfg.pyr, lv, cfg.nlevels, b.x4tab, b.ytab,
                   b.pyr_col, b.pyr_row);
        return;
    }
    DVM_LAUNCH(pyramid_kernel, dim3(cfg.pyr.ntx, cfg.pyr.nty), kPyrThreads, cfg.pyr.smem_bytes, stream, cfg.pyr, lv, cfg.nlevels, b.xtab, b.ytab,
               b.pyr_col, b.pyr_row);
}

// -------------------------------------------------------------------------------------- FAST cells
// Ring of radius 3, clockwise from (0,+3) (cv::FAST, patternSize 16).
__device__ __constant__ int8_t c_ring_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
__device__ __constant__ int8_t c_ring_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

constexpr int kCellThreads = 128;
constexpr int kScorePitch = 84;
constexpr int kCellListCap = 1700;

// One CTA per cell of the reference's 35-px grid.  The reference runs cv::FAST on the cell's
// sub-image [ini, ini+cell+6) with threshold iniThFAST and, if that returns nothing, again with
// minThFAST.  cv::FAST's 3x3 non-max suppression compares scores strictly, with everything outside
// the sub-image's computed region [3, dim-3) counting as 0; since all scores in one call share one
// threshold this is "strict local maximum of the arc measure m inside the region, and m > th".
__global__ void __launch_bounds__(kCellThreads) fast_cells_kernel(const __grid_constant__ OrbCfg cfg, OrbBuffers b,
                                                                  const __grid_constant__ OrbTmaps tm)
{
    __shared__ __align__(128) uint8_t raw[kCellMaxDim * kCellTilePitch];
    __shared__ __align__(8) unsigned long long tma_bar;
    __shared__ uint8_t score[(kCellMaxDim - 6) * kScorePitch];
    __shared__ uint32_t list[kCellListCap];
    __shared__ int s_nlist, s_nini, s_base, s_rank;

    int level = 0;
#pragma unroll 1
    for (int l = 1; l < cfg.nlevels; l++)
        if ((int)blockIdx.x >= cfg.lv[l].cell_base) level = l;
    const OrbLevel& L = cfg.lv[level];
    const int cell = blockIdx.x - L.cell_base;
    const int ci = cell / L.nCols, cj = cell - ci * L.nCols;

    // cell geometry relative to the border origin (16,16): O3/src/ORBextractor.cc:634-649
    const int iniY = ci * L.hCell, iniX = cj * L.wCell;
    if (iniY >= L.height - 3 || iniX >= L.width - 6) return;
    const int maxY = min(iniY + L.hCell + 6, L.height), maxX = min(iniX + L.wCell + 6, L.width);
    const int sw = maxX - iniX, sh = maxY - iniY;
    if (sw < 7 || sh < 7) return;
    const int cw = sw - 6, ch = sh - 6;

    if (threadIdx.x == 0) { s_nlist = 0; s_nini = 0; s_rank = 0; }
    const uint8_t* tile = raw + (tm.use[level] ? ((kBorder + iniX) & 15) : 0);
    if (tm.use[level]) {
        // the cell's tile (kCellTilePitch x (hCell + 6) bytes, zero-filled past the image) by ONE TMA tensor copy.  The
        // box must start on a 16-byte boundary of the row (an unaligned inner coordinate faults as an illegal instruction),
        // so it starts at the aligned column below the cell and the kernel reads the tile `xoff` bytes in.
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tma_bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)(kCellTilePitch * tm.box_h[level]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(raw)), "l"(reinterpret_cast<uint64_t>(&tm.map[level])),
                           "r"((kBorder + iniX) & ~15), "r"(kBorder + iniY), "r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(0) : "memory");
    } else {
        const uint8_t* src = L.img + (size_t)(kBorder + iniY) * L.pitch + (kBorder + iniX);
        for (int idx = threadIdx.x; idx < sh * sw; idx += kCellThreads) {
            const int r = idx / sw, c = idx - r * sw;
            raw[r * kCellTilePitch + c] = __ldg(src + (size_t)r * L.pitch + c);
        }
        __syncthreads();
    }

    // arc measure of every region pixel, two horizontally adjacent pixels per iteration
    const int pw = (cw + 1) >> 1;
    for (int t = threadIdx.x; t < pw * ch; t += kCellThreads) {
        const int ry = t / pw, rx = (t - ry * pw) * 2;
        const uint8_t* p = tile + (ry + 3) * kCellTilePitch + (rx + 3);
        uint32_t ring[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const uint8_t* q = p + c_ring_dy[k] * kCellTilePitch + c_ring_dx[k];
            ring[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 16);
        }
        int mA, mB;
        fast_measure_x2(ring, p[0], p[1], &mA, &mB);
        score[ry * kScorePitch + rx] = (uint8_t)mA;
        if (rx + 1 < cw) score[ry * kScorePitch + rx + 1] = (uint8_t)mB;
    }
    __syncthreads();

    // strict 3x3 maxima above minThFAST
    for (int t = threadIdx.x; t < cw * ch; t += kCellThreads) {
        const int ry = t / cw, rx = t - ry * cw;
        const int m = score[ry * kScorePitch + rx];
        if (m <= cfg.min_th) continue;
        bool ismax = true;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0) continue;
                const int yy = ry + dy, xx = rx + dx;
                if (yy < 0 || yy >= ch || xx < 0 || xx >= cw) continue;
                ismax = ismax && (m > (int)score[yy * kScorePitch + xx]);
            }
        if (!ismax) continue;
        const int slot = atomicAdd(&s_nlist, 1);
        if (m > cfg.ini_th) atomicAdd(&s_nini, 1);
        if (slot < kCellListCap) list[slot] = pack_kp(iniX + rx + 3, iniY + ry + 3, m - 1);
    }
    __syncthreads();

    const int nlist = min(s_nlist, kCellListCap);
    const bool use_ini = s_nini > 0;
    const int nkeep = use_ini ? s_nini : nlist;
    if (nkeep == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(&b.cand_count[level], nkeep);
    __syncthreads();
    const int base = s_base;
    if (base + nkeep > L.cand_cap || s_nlist > kCellListCap) {
        if (threadIdx.x == 0) atomicOr(b.status, 1);
        return;
    }
    for (int t = threadIdx.x; t < nlist; t += kCellThreads) {
        const uint32_t e = list[t];
        if (!use_ini || kp_resp(e) + 1 > cfg.ini_th) {
            const int r = atomicAdd(&s_rank, 1);
            b.cand[L.cand_off + base + r] = e;
        }
    }
}

void launch_fast_cells(const OrbCfg& cfg, const OrbBuffers& b, const OrbTmaps& tm, cudaStream_t stream)
{
    DVM_LAUNCH(fast_cells_kernel, cfg.total_cells, kCellThreads, 0, stream, cfg, b, tm);
}

bool encode_level_tmap(const OrbCfg& cfg, int level, OrbTmaps& tm)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeFn)fn;
    }();
    const OrbLevel& L = cfg.lv[level];
    tm.use[level] = 0;
    tm.box_h[level] = L.hCell + 6;
    if (!encode || ((uintptr_t)L.img & 15) != 0 || (L.pitch & 15) != 0) return false;
    const cuuint64_t dims[2] = { (cuuint64_t)L.w, (cuuint64_t)L.h };
    const cuuint64_t strides[1] = { (cuuint64_t)L.pitch };
    const cuuint32_t box[2] = { (cuuint32_t)kCellTilePitch, (cuuint32_t)(L.hCell + 6) };
    const cuuint32_t estr[2] = { 1, 1 };
    const CUresult r = encode(&tm.map[level], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(L.img), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    tm.use[level] = r == CUDA_SUCCESS ? 1 : 0;
    return tm.use[level] != 0;
}

// ------------------------------------------------------------------------------------------ octree
struct Node {
    short x0, y0, x1, y1; // [x0,x1) x [y0,y1) relative to the border origin
    int cnt;
};

__device__ inline int node_quadrant(const Node& nd, int x, int y)
{
    // ExtractorNode::DivideNode: halfX = ceil((UR.x-UL.x)/2), O3/src/ORBextractor.cc:349-350,378-389
    const int hx = (nd.x1 - nd.x0 + 1) >> 1, hy = (nd.y1 - nd.y0 + 1) >> 1;
    const bool left = x < nd.x0 + hx, top = y < nd.y0 + hy;
    return left ? (top ? 0 : 2) : (top ? 1 : 3);
}

__device__ inline Node node_child(const Node& nd, int q, int cnt)
{
    const int hx = (nd.x1 - nd.x0 + 1) >> 1, hy = (nd.y1 - nd.y0 + 1) >> 1;
    Node c;
    c.x0 = (q & 1) ? nd.x0 + hx : nd.x0;
    c.x1 = (q & 1) ? nd.x1 : nd.x0 + hx;
    c.y0 = (q & 2) ? nd.y0 + hy : nd.y0;
    c.y1 = (q & 2) ? nd.y1 : nd.y0 + hy;
    c.cnt = cnt;
    return c;
}

constexpr int kOctThreads = 1024;

struct OctSmem { // layout helper for the dynamic shared memory of octree_kernel
    Node* A; Node* B; int* cc; int* proc; int* cands; int* cands2; int* removed; unsigned long long* sortbuf;
};

__host__ __device__ inline size_t oct_smem_bytes_for(int cap)
{
    // A, B (12 B each), cc (16 B), proc, cands, cands2, removed (4 B each), sortbuf (8 B)
    return (size_t)cap * (12 + 12 + 16 + 4 + 4 + 4 + 4 + 8) + 64;
}

int octree_smem_bytes(const OrbCfg& cfg)
{
    int cap = 0;
    for (int l = 0; l < cfg.nlevels; l++) cap = max(cap, cfg.lv[l].node_cap);
    return (int)oct_smem_bytes_for(cap);
}

// thread-local serial scan + one block scan: thread `tid` owns items [tid*ipt, min(n,(tid+1)*ipt))
#define OCT_ITEMS(n)                                           \
    const int ipt__ = div_up(max((n), 1), kOctThreads);        \
    const int it0 = min((int)threadIdx.x * ipt__, (n));        \
    const int it1 = min(it0 + ipt__, (n));

__global__ void __launch_bounds__(kOctThreads, 1)
octree_kernel(const __grid_constant__ OrbCfg cfg, OrbBuffers b, int lap0, int lap1)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int warp_sums[33];
    __shared__ int s_size, s_M, s_nc, s_finish, s_phase, s_Meff, s_ticket;
    __shared__ int s_stk[192];

    const int level = blockIdx.x;
    const OrbLevel& L = cfg.lv[level];
    const int cap = L.node_cap;
    const int tid = threadIdx.x;
    const int N = L.quota;

    // carve shared memory (sized for the largest level)
    int capmax = 0;
    for (int l = 0; l < cfg.nlevels; l++) capmax = max(capmax, cfg.lv[l].node_cap);
    unsigned char* sp = smem_raw;
    unsigned long long* sortbuf = (unsigned long long*)sp; sp += (size_t)capmax * 8;
    int* cc = (int*)sp; sp += (size_t)capmax * 16;
    Node* A = (Node*)sp; sp += (size_t)capmax * 12;
    Node* B = (Node*)sp; sp += (size_t)capmax * 12;
    int* proc = (int*)sp; sp += (size_t)capmax * 4;
    int* cands = (int*)sp; sp += (size_t)capmax * 4;
    int* cands2 = (int*)sp; sp += (size_t)capmax * 4;
    int* removed = (int*)sp;

    const int n = min(b.cand_count[level], L.cand_cap);
    const uint32_t* cand = b.cand + L.cand_off;
    uint16_t* pnode = b.pnode + L.cand_off;

    // ---- root nodes (O3/src/ORBextractor.cc:422-458) ----
    for (int i = tid; i < L.nIni; i += kOctThreads) {
        Node nd;
        nd.x0 = (short)(int)__fmul_rn(L.hX, (float)i);
        nd.x1 = (short)(int)__fmul_rn(L.hX, (float)(i + 1));
        nd.y0 = 0;
        nd.y1 = (short)L.height;
        nd.cnt = 0;
        A[i] = nd;
    }
    __syncthreads();
    for (int p = tid; p < n; p += kOctThreads) {
        const int x = kp_x(cand[p]);
        int id = (int)__fdiv_rn((float)x, L.hX);
        id = min(id, L.nIni - 1);
        pnode[p] = (uint16_t)id;
        atomicAdd(&A[id].cnt, 1);
    }
    __syncthreads();
    if (tid == 0) { // drop empty roots
        int m = 0;
        for (int i = 0; i < L.nIni; i++) {
            cc[4 * i] = m;
            if (A[i].cnt > 0) B[m++] = A[i];
        }
        s_size = m;
        s_phase = 1;
        s_finish = 0;
        s_nc = 0;
    }
    __syncthreads();
    for (int p = tid; p < n; p += kOctThreads) pnode[p] = (uint16_t)cc[4 * pnode[p]];
    { Node* t = A; A = B; B = t; }
    __syncthreads();

    // ---- subdivision rounds ----
    while (true) {
        const int sizeA = s_size;
        const int phase = s_phase;
        const int prevSize = sizeA;
        // processing order
        if (phase == 1) {
            OCT_ITEMS(sizeA);
            int c = 0;
            for (int i = it0; i < it1; i++) c += (A[i].cnt > 1);
            int tot;
            int off = block_exclusive_scan(c, warp_sums, &tot);
            for (int i = it0; i < it1; i++)
                if (A[i].cnt > 1) proc[off++] = i;
            if (tid == 0) s_M = tot;
        } else {
            // sort (size, UL.x) ascending exactly as libstdc++ would, walk it backwards
            // (O3/src/ORBextractor.cc:544-547); cc is dead here and serves as scratch
            const int M0 = s_nc;
            for (int k = tid; k < M0; k += kOctThreads) {
                const Node& nd = A[cands[k]];
                sortbuf[k] = ((unsigned long long)(((unsigned)nd.cnt << 12) | (unsigned)nd.x0) << 32) | (unsigned)cands[k];
            }
            __syncthreads();
            libstdcxx_sort_cta(sortbuf, (unsigned long long*)cc, cc + 2 * capmax, cc + 3 * capmax, s_stk, M0, KeyHi32Less());
            for (int k = tid; k < M0; k += kOctThreads) proc[k] = (int)(unsigned)sortbuf[M0 - 1 - k];
            if (tid == 0) s_M = M0;
            __syncthreads();
        }
        for (int i = tid; i < sizeA * 4; i += kOctThreads) cc[i] = 0;
        for (int i = tid; i < sizeA; i += kOctThreads) removed[i] = 0;
        if (tid == 0) s_Meff = 0x7fffffff;
        __syncthreads();
        int M = s_M;
        if (M == 0) break; // nothing left to split: list size unchanged -> finished

        // child occupancy of every splittable node
        for (int p = tid; p < n; p += kOctThreads) {
            const int id = pnode[p];
            const Node nd = A[id];
            if (nd.cnt > 1) {
                const uint32_t e = cand[p];
                atomicAdd(&cc[4 * id + node_quadrant(nd, kp_x(e), kp_y(e))], 1);
            }
        }
        __syncthreads();

        // early break of the largest-first phase: node t is split iff the list was still short
        // before it (O3/src/ORBextractor.cc:583-584)
        if (phase == 2) {
            OCT_ITEMS(M);
            int c = 0;
            for (int t = it0; t < it1; t++) {
                const int* q = cc + 4 * proc[t];
                c += (q[0] > 0) + (q[1] > 0) + (q[2] > 0) + (q[3] > 0) - 1;
            }
            int tot;
            int before = sizeA + block_exclusive_scan(c, warp_sums, &tot);
            for (int t = it0; t < it1; t++) {
                if (t > 0 && before >= N) { atomicMin(&s_Meff, t); break; }
                const int* q = cc + 4 * proc[t];
                before += (q[0] > 0) + (q[1] > 0) + (q[2] > 0) + (q[3] > 0) - 1;
            }
            __syncthreads();
            M = min(M, s_Meff);
        }
        for (int t = tid; t < M; t += kOctThreads) removed[proc[t]] = 1;
        __syncthreads();

        // offsets: children (push_front order) and expandable children (creation order)
        int childOff, expOff, Stot, Etot;
        {
            OCT_ITEMS(M);
            int c = 0;
            for (int t = it0; t < it1; t++) {
                const int* q = cc + 4 * proc[t];
                c += ((q[0] > 0) + (q[1] > 0) + (q[2] > 0) + (q[3] > 0)) | (((q[0] > 1) + (q[1] > 1) + (q[2] > 1) + (q[3] > 1)) << 16);
            }
            int tot;
            const int off = block_exclusive_scan(c, warp_sums, &tot);
            childOff = off & 0xffff; expOff = off >> 16;
            Stot = tot & 0xffff; Etot = tot >> 16;
        }
        int survOff, nSurv;
        {
            OCT_ITEMS(sizeA);
            int c = 0;
            for (int i = it0; i < it1; i++) c += !removed[i];
            survOff = block_exclusive_scan(c, warp_sums, &nSurv);
        }
        const int newSize = Stot + nSurv;
        if (newSize > cap) { // cannot happen for the reference's bounds; fail loudly
            if (tid == 0) { atomicOr(b.status, 2); s_size = 0; }
            __syncthreads();
            break;
        }
        {
            OCT_ITEMS(M);
            int P = childOff, E = expOff;
            for (int t = it0; t < it1; t++) {
                const int id = proc[t];
                const Node nd = A[id];
                int q[4] = { cc[4 * id], cc[4 * id + 1], cc[4 * id + 2], cc[4 * id + 3] };
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (q[k] > 0) {
                        const int pos = Stot - 1 - P; // k-th push_front ends up here
                        P++;
                        B[pos] = node_child(nd, k, q[k]);
                        if (q[k] > 1) cands2[E++] = pos;
                        cc[4 * id + k] = pos;
                    }
                }
            }
        }
        {
            OCT_ITEMS(sizeA);
            int r = survOff;
            for (int i = it0; i < it1; i++)
                if (!removed[i]) {
                    B[Stot + r] = A[i];
                    cc[4 * i] = Stot + r;
                    r++;
                }
        }
        __syncthreads();
        for (int p = tid; p < n; p += kOctThreads) {
            const int id = pnode[p];
            if (removed[id]) {
                const uint32_t e = cand[p];
                pnode[p] = (uint16_t)cc[4 * id + node_quadrant(A[id], kp_x(e), kp_y(e))];
            } else {
                pnode[p] = (uint16_t)cc[4 * id];
            }
        }
        __syncthreads();
        { Node* t = A; A = B; B = t; }
        { int* t = cands; cands = cands2; cands2 = t; }
        if (tid == 0) {
            s_size = newSize;
            s_nc = Etot;
            if (newSize >= N || newSize == prevSize) s_finish = 1;
            else if (phase == 1 && newSize + 3 * Etot > N) s_phase = 2;
        }
        __syncthreads();
        if (s_finish) break;
    }
    __syncthreads();

    // ---- best response per node, earliest candidate wins ties (O3/src/ORBextractor.cc:592-607) ----
    const int size = s_size;
    unsigned long long* best = (unsigned long long*)cc;
    for (int i = tid; i < size; i += kOctThreads) best[i] = 0ull;
    __syncthreads();
    for (int p = tid; p < n; p += kOctThreads) {
        const uint32_t e = cand[p];
        const int x = kp_x(e), y = kp_y(e);
        // the reference's candidate order: cells row-major, pixels row-major inside a cell
        const unsigned long long cellidx = (unsigned)((y - 3) / L.hCell) * (unsigned)L.nCols + (unsigned)((x - 3) / L.wCell);
        const unsigned long long key = (cellidx << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
        const unsigned long long v = ((unsigned long long)kp_resp(e) << 40) | (0xffffffffffull - key);
        atomicMax(&best[pnode[p]], v);
    }
    __syncthreads();
    for (int i = tid; i < size; i += kOctThreads) {
        const unsigned long long v = best[i];
        const unsigned long long key = 0xffffffffffull - (v & 0xffffffffffull);
        b.sel[L.sel_off + i] = pack_kp((int)(key & 0xfff), (int)((key >> 12) & 0xfff), (int)(v >> 40));
    }
    if (tid == 0) b.sel_count[level] = size;

    // ---- the last CTA to finish lays out the output order (O3/src/ORBextractor.cc:906-950) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = (int)atomicAdd(b.ticket, 1u);
    __syncthreads();
    if (s_ticket != (int)gridDim.x - 1) return;
    __threadfence();
    int lvl_off[kMaxLevels + 1];
    lvl_off[0] = 0;
    for (int l = 0; l < cfg.nlevels; l++) lvl_off[l + 1] = lvl_off[l] + ((volatile int*)b.sel_count)[l];
    const int nk = lvl_off[cfg.nlevels];
    {
        OCT_ITEMS(nk);
        // keypoints whose scaled x lies in [lap0, lap1] are written from the back
        int c = 0, l = 0;
        for (int s = it0; s < it1; s++) {
            while (s >= lvl_off[l + 1]) l++;
            const uint32_t e = ((volatile uint32_t*)b.sel)[cfg.lv[l].sel_off + (s - lvl_off[l])];
            float x = (float)(kp_x(e) + kBorder);
            if (l != 0) x = __fmul_rn(x, cfg.lv[l].scale);
            c += (x >= (float)lap0 && x <= (float)lap1);
        }
        int nStereo;
        int stereoBefore = block_exclusive_scan(c, warp_sums, &nStereo);
        l = 0;
        for (int s = it0; s < it1; s++) {
            while (s >= lvl_off[l + 1]) l++;
            const uint32_t e = ((volatile uint32_t*)b.sel)[cfg.lv[l].sel_off + (s - lvl_off[l])];
            float x = (float)(kp_x(e) + kBorder);
            if (l != 0) x = __fmul_rn(x, cfg.lv[l].scale);
            const bool st = (x >= (float)lap0 && x <= (float)lap1);
            const int dest = st ? (nk - 1 - stereoBefore) : (s - stereoBefore);
            stereoBefore += st;
            b.work_kp[s] = e;
            b.work_meta[s] = ((uint32_t)l << 24) | (uint32_t)dest;
        }
        if (tid == 0) {
            b.counts[0] = nk;
            b.counts[1] = nk - nStereo;
            *b.ticket = 0u;
            for (int l2 = 0; l2 < cfg.nlevels; l2++) { // keep a copy for stage read-back, re-arm for the next frame
                b.cand_count[kMaxLevels + l2] = b.cand_count[l2];
                b.cand_count[l2] = 0;
            }
        }
    }
}

int prepare_octree_kernel(int smem_bytes)
{
    // the attribute is per (function, device) and shared by every handle in the process: only raise it
    static std::mutex mu;
    static int prepared[64] = { 0 };
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && smem_bytes <= prepared[dev]) return DVM_OK;
    cudaError_t e = cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
        set_error("octree kernel needs %d bytes of shared memory: %s", smem_bytes, cudaGetErrorString(e));
        return DVM_ERR_CUDA;
    }
    if (dev >= 0 && dev < 64) prepared[dev] = smem_bytes;
    return DVM_OK;
}

void launch_octree(const OrbCfg& cfg, const OrbBuffers& b, int lap0, int lap1, int smem_bytes, cudaStream_t stream)
{
    DVM_LAUNCH(octree_kernel, cfg.nlevels, kOctThreads, smem_bytes, stream, cfg, b, lap0, lap1);
}

// ---------------------------------------------------------------------------------------- describe
constexpr int kDescWarps = 4;
constexpr int kWin = 43, kWinPitch = 44;   // unblurred window: +-21 around the keypoint
constexpr int kBlr = 37, kBlrPitch = 40;   // blurred window: +-18 (pattern radius 18.4 rounds to <= 18)
constexpr int kHrowPitch = 38;

__device__ __constant__ int c_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

__device__ inline int reflect101(int p, int n)
{
    // BORDER_REFLECT_101; |overshoot| < n here
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(kDescWarps * 32) describe_kernel(const __grid_constant__ OrbCfg cfg, OrbBuffers b)
{
    __shared__ uint8_t s_raw[kDescWarps][kWin * kWinPitch];
    __shared__ uint16_t s_hrow[kDescWarps][kWin * kHrowPitch];
    __shared__ uint8_t s_blr[kDescWarps][kBlr * kBlrPitch];
    __shared__ int8_t s_pat[1024];

    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_pat[i] = b.pattern[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kDescWarps + warp;
    if (s >= b.counts[0]) return;

    const uint32_t e = b.work_kp[s];
    const uint32_t meta = b.work_meta[s];
    const int level = meta >> 24, dest = meta & 0xffffff;
    const OrbLevel& L = cfg.lv[level];
    const int x = kp_x(e) + kBorder, y = kp_y(e) + kBorder;

    uint8_t* raw = s_raw[warp];
    uint16_t* hrow = s_hrow[warp];
    uint8_t* blr = s_blr[warp];

    for (int idx = lane; idx < kWin * kWin; idx += 32) {
        const int r = idx / kWin, c = idx - r * kWin;
        const int yy = reflect101(y - 21 + r, L.h), xx = reflect101(x - 21 + c, L.w);
        raw[r * kWinPitch + c] = __ldg(L.img + (size_t)yy * L.pitch + xx);
    }
    __syncwarp();

    // IC_Angle: first-order moments over the radius-15 disc of the unblurred level (:75-99)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - kHalfPatch;
        const int d = c_umax[v < 0 ? -v : v];
        const uint8_t* rowp = raw + (21 + v) * kWinPitch + 21;
        int sum = 0;
        for (int u = -d; u <= d; ++u) {
            const int val = rowp[u];
            m10 += u * val;
            sum += val;
        }
        m01 = v * sum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // GaussianBlur 7x7 sigma 2, 8-bit fixed point: k = {18,34,48,56,48,34,18}/256 per axis, one
    // rounding at the end (:919-920)
    for (int idx = lane; idx < kWin * kBlr; idx += 32) {
        const int r = idx / kBlr, c = idx - r * kBlr;
        const uint8_t* p = raw + r * kWinPitch + c;
        const int v = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        hrow[r * kHrowPitch + c] = (uint16_t)v;
    }
    __syncwarp();
    for (int idx = lane; idx < kBlr * kBlr; idx += 32) {
        const int r = idx / kBlr, c = idx - r * kBlr;
        const uint16_t* q = hrow + r * kHrowPitch + c;
        const unsigned v = 18u * (q[0] + q[6 * kHrowPitch]) + 34u * (q[kHrowPitch] + q[5 * kHrowPitch]) +
                           48u * (q[2 * kHrowPitch] + q[4 * kHrowPitch]) + 56u * q[3 * kHrowPitch];
        blr[r * kBlrPitch + c] = (uint8_t)((v + 32768u) >> 16);
    }
    __syncwarp();

    // computeOrbDescriptor (:101-143): lane = descriptor byte, 8 tests each
    const float factorPI = (float)(3.14159265358979323846 / 180.0);
    const float rad = __fmul_rn(angle, factorPI);
    const float ca = glibc_sincosf(rad, 1), sb = glibc_sincosf(rad, 0);
    const uint8_t* center = blr + 18 * kBlrPitch + 18;
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int8_t* pt = s_pat + (lane * 8 + k) * 4;
        const float x0 = (float)pt[0], y0 = (float)pt[1], x1 = (float)pt[2], y1 = (float)pt[3];
        const int t0 = center[cv_round(__fadd_rn(__fmul_rn(x0, sb), __fmul_rn(y0, ca))) * kBlrPitch +
                              cv_round(__fsub_rn(__fmul_rn(x0, ca), __fmul_rn(y0, sb)))];
        const int t1 = center[cv_round(__fadd_rn(__fmul_rn(x1, sb), __fmul_rn(y1, ca))) * kBlrPitch +
                              cv_round(__fsub_rn(__fmul_rn(x1, ca), __fmul_rn(y1, sb)))];
        val |= (t0 < t1) << k;
    }
    b.out_desc[(size_t)dest * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        dvm_keypoint kp;
        kp.x = (float)x;
        kp.y = (float)y;
        if (level != 0) { kp.x = __fmul_rn(kp.x, L.scale); kp.y = __fmul_rn(kp.y, L.scale); }
        kp.size = L.size;
        kp.angle = angle;
        kp.response = (float)kp_resp(e);
        kp.octave = level;
        kp.class_id = -1;
        b.out_kps[dest] = kp;
    }
}

void launch_describe(const OrbCfg& cfg, const OrbBuffers& b, cudaStream_t stream)
{
    DVM_LAUNCH(describe_kernel, div_up(cfg.max_kp, kDescWarps), kDescWarps * 32, 0, stream, cfg, b);
}

// -------------------------------------------------------------------------- full-level blur (debug)
__global__ void blur_level_debug_kernel(const uint8_t* __restrict__ src, int w, int h, int spitch,
                                        uint8_t* __restrict__ dst, int dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int k[7] = { 18, 34, 48, 56, 48, 34, 18 };
    unsigned acc = 0;
    for (int j = 0; j < 7; j++) {
        const uint8_t* row = src + (size_t)reflect101(y + j - 3, h) * spitch;
        unsigned r = 0;
        for (int i = 0; i < 7; i++) r += k[i] * row[reflect101(x + i - 3, w)];
        acc += k[j] * r;
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)((acc + 32768u) >> 16);
}

void launch_blur_level_debug(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, cudaStream_t stream)
{
    dim3 block(32, 8), grid(div_up(w, 32), div_up(h, 8));
    DVM_LAUNCH(blur_level_debug_kernel, grid, block, 0, stream, src, w, h, spitch, dst, dpitch);
}

} // namespace dvm
