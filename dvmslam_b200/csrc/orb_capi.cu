// orb_capi.cu -- C-ABI of the ORB extractor (include/dvmslam_b200.h): host-side state that the
// reference keeps in ORB_SLAM3::ORBextractor (scale tables, per-level quotas, pyramid), geometry
// tables, buffer ownership and the launch sequence of one extract call.
#include "orb_kernels.cuh"
#include "orb_math.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

using namespace dvm;

static const int8_t kPatternHost[256 * 4] = {
#include "orb_pattern.inc"
};

struct dvm_orb {
    int device = 0;
    cudaStream_t stream = nullptr;
    // ORBextractor members (O3/include/ORBextractor.h:80-95)
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    double scaleFactor = 0;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> perLevel;
    int max_w = 0, max_h = 0;
    int max_kp = 0;
    // geometry of the current image size
    int cur_w = -1, cur_h = -1;
    OrbCfg cfg;
    OrbBuffers buf;
    OrbTmaps tmaps;                   // TMA descriptors of the levels (FAST cell tiles)
    const uint8_t* tmap0_img = nullptr; int tmap0_pitch = -1; // what tmaps.map[0] was encoded for
    int oct_smem = 0, oct_smem_prepared = 0;
    // device allocations
    uint8_t* d_pyr = nullptr;
    size_t pyr_bytes = 0;
    size_t lvl_off[kMaxLevels] = { 0 };
    uint8_t* d_dbg = nullptr;
    size_t dbg_bytes = 0;
    size_t cand_total = 0;
    ResizeX* d_xtab = nullptr;
    ResizeX4* d_x4tab = nullptr;
    size_t x4_cap = 0;
    ResizeY* d_ytab = nullptr;
    int* d_pyr_col = nullptr; int* d_pyr_row = nullptr; size_t pyr_col_cap = 0, pyr_row_cap = 0;
    bool fused_pyramid = true;
    size_t xtab_cap = 0, ytab_cap = 0;
    uint8_t* d_out = nullptr; // [counts 16 B | status 16 B | kps | desc]
    size_t out_bytes = 0;
    uint8_t* h_out = nullptr; // pinned mirror of d_out
    int8_t* d_pattern = nullptr;
    bool level0_aliased = false;
    // per-stage profiling (bench.py roofline)
    bool profiling = false;
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    int prof_frames = 0;
    float prof_ms[4] = { 0, 0, 0, 0 };
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void level_size(const dvm_orb* h, int w, int hgt, int l, int* lw, int* lh)
{
    // Size sz(cvRound((float)image.cols*scale), cvRound((float)image.rows*scale)), O3/src/ORBextractor.cc:959-960
    *lw = (int)lrintf((float)w * h->invScale[l]);
    *lh = (int)lrintf((float)hgt * h->invScale[l]);
}

static void free_all(dvm_orb* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_pyr); cudaFree(h->d_dbg); cudaFree(h->d_xtab); cudaFree(h->d_x4tab); cudaFree(h->d_ytab); cudaFree(h->d_pyr_col); cudaFree(h->d_pyr_row); cudaFree(h->d_out);
    cudaFree(h->d_pattern);
    cudaFree(h->buf.cand); cudaFree(h->buf.cand_count); cudaFree(h->buf.pnode); cudaFree(h->buf.sel);
    cudaFree(h->buf.sel_count); cudaFree(h->buf.work_kp); cudaFree(h->buf.work_meta); cudaFree(h->buf.ticket);
    if (h->h_out) cudaFreeHost(h->h_out);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Per-size geometry: level sizes, cell grid, octree roots, resize tables.  Mirrors the integer /
// float expressions of ComputePyramid, ComputeKeyPointsOctTree and DistributeOctTree.
constexpr int kPyrMaxSpanHost = 192;   // = kPyrMaxSpan of orb_kernels.cu (table slices in shared memory)
static int configure(dvm_orb* h, int w, int hgt)
{
    if (w == h->cur_w && hgt == h->cur_h) return DVM_OK;
    DVM_REQUIRE(w <= h->max_w && hgt <= h->max_h, "image larger than the size given to dvm_orb_create");
    OrbCfg& c = h->cfg;
    memset(&c, 0, sizeof(c));
    c.nlevels = h->nlevels;
    c.ini_th = h->iniTh;
    c.min_th = h->minTh;
    std::vector<ResizeX> xt;
    std::vector<ResizeY> yt;
    int cells = 0, cand_off = 0, sel_off = 0;
    for (int l = 0; l < h->nlevels; l++) {
        OrbLevel& L = c.lv[l];
        level_size(h, w, hgt, l, &L.w, &L.h);
        L.pitch = (int)align_up(L.w, 128);
        L.img = h->d_pyr + h->lvl_off[l];
        L.width = L.w - 2 * kBorder;
        L.height = L.h - 2 * kBorder;
        DVM_REQUIRE(L.width >= 35 && L.height >= 35, "image too small for this many pyramid levels");
        DVM_REQUIRE(L.width < 4096 && L.height < 4096, "image too large (level dimension must stay below 4096+32)");
        const float W = 35;
        L.nCols = (int)((float)L.width / W);
        L.nRows = (int)((float)L.height / W);
        L.wCell = (int)std::ceil((float)L.width / L.nCols);
        L.hCell = (int)std::ceil((float)L.height / L.nRows);
        DVM_REQUIRE(L.wCell + 6 <= kCellMaxDim && L.hCell + 6 <= kCellMaxDim, "FAST cell larger than the kernel tile");
        L.cell_base = cells;
        cells += L.nCols * L.nRows;
        L.quota = h->perLevel[l];
        L.nIni = (int)std::round((float)L.width / (float)L.height);
        DVM_REQUIRE(L.nIni >= 1 && L.nIni <= 32, "aspect ratio outside the supported 1:2 .. 32:1 range");
        L.hX = (float)L.width / L.nIni;
        L.cand_off = cand_off;
        // strict 3x3 maxima cannot be 8-adjacent: at most ceil(cw/2)*ceil(ch/2) per cell, so this never overflows
        L.cand_cap = ((L.width + 1) / 2 + L.nCols) * ((L.height + 1) / 2 + L.nRows);
        cand_off += L.cand_cap;
        L.node_cap = std::max(L.quota + 4, 4 * L.nIni + 4);
        DVM_REQUIRE(L.node_cap < 65535, "per-level feature quota too large");
        L.sel_off = sel_off;
        sel_off += L.node_cap;
        L.scale = h->scale[l];
        L.size = (float)(int)(31 * h->scale[l]); // const int scaledPatchSize = PATCH_SIZE*mvScaleFactor[level]
        L.xtab_off = (int)xt.size();
        L.ytab_off = (int)yt.size();
        if (l > 0) {
            const OrbLevel& S = c.lv[l - 1];
            for (int d = 0; d < L.w; d++) {
                ResizeX e;
                resize_coef(d, S.w, L.w, true, &e.sx0, &e.sx1, &e.a0, &e.a1);
                xt.push_back(e);
            }
            for (int d = 0; d < L.h; d++) {
                ResizeY e;
                resize_coef(d, S.h, L.h, false, &e.sy0, &e.sy1, &e.b0, &e.b1);
                yt.push_back(e);
            }
        }
    }
    // groups of four destination columns for the vectorised pyramid (see ResizeX4)
    std::vector<ResizeX4> x4;
    bool vec_ok = true;
    for (int l = 1; l < h->nlevels; l++) {
        OrbLevel& L = c.lv[l];
        L.x4_off = (int)x4.size();
        const ResizeX* X = xt.data() + L.xtab_off;
        for (int g = 0; g * 4 < L.w; g++) {
            ResizeX4 e;
            memset(&e, 0, sizeof(e));
            int o0[4], o1[4];
            const int bw = X[std::min(4 * g, L.w - 1)].sx0 >> 2;
            for (int p2 = 0; p2 < 4; p2++) {
                const ResizeX& t = X[std::min(4 * g + p2, L.w - 1)];
                o0[p2] = t.sx0 - 4 * bw; o1[p2] = t.sx1 - 4 * bw;
                e.c[p2] = (uint32_t)(uint16_t)t.a0 | ((uint32_t)(uint16_t)t.a1 << 16);
                vec_ok = vec_ok && t.a0 >= 0 && t.a1 >= 0;
            }
            uint32_t sel = 0, win = 0;
            for (int pr = 0; pr < 2; pr++) {
                const int a = 2 * pr, b2 = 2 * pr + 1;
                const int lo = std::min(std::min(o0[a], o1[a]), std::min(o0[b2], o1[b2]));
                const int hi = std::max(std::max(o0[a], o1[a]), std::max(o0[b2], o1[b2]));
                const int wsh = lo >= 4 ? 4 : 0;
                vec_ok = vec_ok && lo >= 0 && hi - wsh <= 7 && hi <= 11;
                const uint32_t sl = (uint32_t)((o0[a] - wsh) & 7) | (uint32_t)((o1[a] - wsh) & 7) << 4 | (uint32_t)((o0[b2] - wsh) & 7) << 8 |
                                    (uint32_t)((o1[b2] - wsh) & 7) << 12;
                sel |= sl << (16 * pr);
                if (wsh) win |= 1u << (16 + pr);
            }
            e.sel = sel;
            e.base = (uint32_t)bw | win;
            vec_ok = vec_ok && bw < 65536;
            x4.push_back(e);
        }
    }
    DVM_REQUIRE(x4.size() <= h->x4_cap, "resize table too small");
    c.total_cells = cells;
    c.max_kp = sel_off;
    DVM_REQUIRE((size_t)cand_off <= h->cand_total, "candidate buffer too small");
    DVM_REQUIRE(sel_off <= h->max_kp, "keypoint buffer too small");
    DVM_REQUIRE(xt.size() <= h->xtab_cap && yt.size() <= h->ytab_cap, "resize table too small");
    h->oct_smem = octree_smem_bytes(c);
    DVM_REQUIRE(h->oct_smem <= 227 * 1024, "per-level feature quota needs more shared memory than one SM has");
    {
        int rc = prepare_octree_kernel(h->oct_smem);
        if (rc != DVM_OK) return rc;
    }
    // ---- plan of the one-launch pyramid: per tile column / row of the last level, the region of every level ----
    std::vector<int> colr, rowr;
    h->fused_pyramid = c.nlevels >= 2;
    if (c.nlevels >= 2) {
        const int L = c.nlevels - 1;
        OrbPyrPlan& P = c.pyr;
        P.ntx = (c.lv[L].w + kPyrTileW - 1) / kPyrTileW;
        P.nty = (c.lv[L].h + kPyrTileH - 1) / kPyrTileH;
        auto plan_axis = [&](bool is_x, int ntiles, int tile, std::vector<int>& out, int* maxlen) {
            out.assign((size_t)ntiles * kMaxLevels * 3, 0);
            auto dim = [&](int l) { return is_x ? c.lv[l].w : c.lv[l].h; };
            auto s0 = [&](int l, int d) { return is_x ? xt[c.lv[l].xtab_off + d].sx0 : yt[c.lv[l].ytab_off + d].sy0; };   // first tap of pixel d of level l
            auto s1 = [&](int l, int d) { return is_x ? xt[c.lv[l].xtab_off + d].sx1 : yt[c.lv[l].ytab_off + d].sy1; };
            std::vector<int> start((size_t)(ntiles + 1) * kMaxLevels, 0);
            for (int i = 0; i <= ntiles; i++) {   // first pixel of tile i's region at every level (tile ntiles = one past the end)
                int d = std::min(i * tile, dim(L));
                start[(size_t)i * kMaxLevels + L] = d;
                for (int l = L - 1; l >= 0; l--) {
                    // (columns: a region starts on a multiple of 4, so that a thread's four pixels are one aligned word)
                    d = d >= dim(l + 1) ? dim(l) : (is_x ? s0(l + 1, d) & ~3 : s0(l + 1, d));
                    start[(size_t)i * kMaxLevels + l] = d;
                }
            }
            for (int l = 0; l < kMaxLevels; l++) maxlen[l] = 0;
            for (int i = 0; i < ntiles; i++) {
                int need1 = 0;
                for (int l = L; l >= 0; l--) {
                    const int first = start[(size_t)i * kMaxLevels + l];
                    const int own1 = (i + 1 == ntiles ? dim(l) : start[(size_t)(i + 1) * kMaxLevels + l]) - 1;
                    need1 = l == L ? own1 : std::max(own1, s1(l + 1, need1));
                    if (is_x) need1 = std::min(dim(l) - 1, first + ((need1 - first + 4) & ~3) - 1);   // whole groups of four columns
                    int* e = &out[((size_t)i * kMaxLevels + l) * 3];
                    e[0] = first; e[1] = own1; e[2] = need1;
                    maxlen[l] = std::max(maxlen[l], need1 - first + 1);
                }
            }
        };
        int mw[kMaxLevels], mh[kMaxLevels];
        plan_axis(true, P.ntx, kPyrTileW, colr, mw);
        plan_axis(false, P.nty, kPyrTileH, rowr, mh);
        int off = 0;
        for (int l = 1; l <= L; l++) {
            P.spitch[l] = (mw[l] + 3) & ~3;
            P.soff[l] = off;
            off += P.spitch[l] * mh[l];
            off = (off + 15) & ~15;
        }
        P.smem_bytes = off;
        P.vec_ok = vec_ok ? 1 : 0;
        bool span_ok = true;
        for (int l = 1; l <= L; l++) span_ok = span_ok && mw[l] <= kPyrMaxSpanHost && mh[l] <= kPyrMaxSpanHost;
        if (off > 200 * 1024 || !span_ok) h->fused_pyramid = false;   // (other scale factors / level counts: keep the per-level launches)
        else if (off > 48 * 1024) prepare_pyramid_kernel(off);
    }
    DVM_CUDA(cudaStreamSynchronize(h->stream)); // nothing in flight may still read the old tables
    if (!colr.empty()) {
        if (colr.size() > h->pyr_col_cap) { cudaFree(h->d_pyr_col); h->d_pyr_col = nullptr; DVM_CUDA(cudaMalloc(&h->d_pyr_col, colr.size() * 4)); h->pyr_col_cap = colr.size(); }
        if (rowr.size() > h->pyr_row_cap) { cudaFree(h->d_pyr_row); h->d_pyr_row = nullptr; DVM_CUDA(cudaMalloc(&h->d_pyr_row, rowr.size() * 4)); h->pyr_row_cap = rowr.size(); }
        DVM_CUDA(cudaMemcpy(h->d_pyr_col, colr.data(), colr.size() * 4, cudaMemcpyHostToDevice));
        DVM_CUDA(cudaMemcpy(h->d_pyr_row, rowr.data(), rowr.size() * 4, cudaMemcpyHostToDevice));
        h->buf.pyr_col = h->d_pyr_col; h->buf.pyr_row = h->d_pyr_row;
    }
    if (!xt.empty()) {
        DVM_CUDA(cudaMemcpy(h->d_xtab, xt.data(), xt.size() * sizeof(ResizeX), cudaMemcpyHostToDevice));
        DVM_CUDA(cudaMemcpy(h->d_ytab, yt.data(), yt.size() * sizeof(ResizeY), cudaMemcpyHostToDevice));
        if (!x4.empty()) DVM_CUDA(cudaMemcpy(h->d_x4tab, x4.data(), x4.size() * sizeof(ResizeX4), cudaMemcpyHostToDevice));
    }
    // TMA descriptors of the levels (level 0 again per call if the caller's image is read in place)
    memset(&h->tmaps, 0, sizeof(h->tmaps));
    for (int l = 0; l < c.nlevels; l++) encode_level_tmap(c, l, h->tmaps);
    h->tmap0_img = c.lv[0].img; h->tmap0_pitch = c.lv[0].pitch;
    h->cur_w = w;
    h->cur_h = hgt;
    return DVM_OK;
}

extern "C" {

int dvm_orb_create(dvm_orb** out, int device, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                   int min_th_fast, int max_width, int max_height)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    DVM_REQUIRE(nfeatures > 0 && nlevels >= 1 && nlevels <= kMaxLevels, "nfeatures/nlevels out of range");
    DVM_REQUIRE(scale_factor > 1.0f, "scaleFactor must exceed 1");
    DVM_REQUIRE(ini_th_fast >= min_th_fast && min_th_fast >= 1 && ini_th_fast < 255, "FAST thresholds out of range");
    DVM_REQUIRE(max_width >= 67 && max_height >= 67, "max image size too small");
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;

    dvm_orb* h = new dvm_orb;
    h->device = device;
    h->nfeatures = nfeatures; h->nlevels = nlevels; h->iniTh = ini_th_fast; h->minTh = min_th_fast;
    h->max_w = max_width; h->max_h = max_height;
    // scale tables and quotas: ORBextractor::ORBextractor, O3/src/ORBextractor.cc:288-316.
    // scaleFactor is a double member initialised from the float argument (ORBextractor.h:83).
    h->scaleFactor = scale_factor;
    h->scale.assign(nlevels, 1.0f); h->sigma2.assign(nlevels, 1.0f);
    for (int i = 1; i < nlevels; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * h->scaleFactor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    h->invScale.resize(nlevels); h->invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        h->invScale[i] = 1.0f / h->scale[i];
        h->invSigma2[i] = 1.0f / h->sigma2[i];
    }
    h->perLevel.resize(nlevels);
    {
        float factor = (float)(1.0f / h->scaleFactor);
        float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; l++) {
            h->perLevel[l] = (int)lrintf(want);
            sum += h->perLevel[l];
            want *= factor;
        }
        h->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
    }
    int quota_sum = 0;
    for (int q : h->perLevel) quota_sum += q;
    h->max_kp = (quota_sum + nlevels * (4 + 4 * 32) + 3) & ~3; // multiple of 4: keeps the descriptor block 16-byte aligned

#define DVM_CREATE_CUDA(call)                                                                        \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error("%s failed in dvm_orb_create: %s", #call, cudaGetErrorString(e__));            \
            free_all(h);                                                                             \
            return DVM_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

    DVM_CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // pyramid: every level at its maximum size, rows padded to 128 B
    size_t off = 0, xt_n = 0, yt_n = 0, x4_n = 0, cand = 0;
    for (int l = 0; l < nlevels; l++) {
        int lw, lh;
        level_size(h, max_width, max_height, l, &lw, &lh);
        h->lvl_off[l] = off;
        off += align_up(lw, 128) * (size_t)(lh + 1) + 256;
        xt_n += lw + 8; yt_n += lh + 8;
        x4_n += lw / 4 + 4;
        cand += (size_t)(lw / 2 + 64) * (lh / 2 + 64);
    }
    h->pyr_bytes = off;
    h->dbg_bytes = align_up(max_width, 128) * (size_t)max_height;
    h->cand_total = cand;
    h->xtab_cap = xt_n; h->ytab_cap = yt_n; h->x4_cap = x4_n;
    DVM_CREATE_CUDA(cudaMalloc(&h->d_pyr, h->pyr_bytes));
    DVM_CREATE_CUDA(cudaMalloc(&h->d_dbg, h->dbg_bytes));
    DVM_CREATE_CUDA(cudaMalloc(&h->d_xtab, xt_n * sizeof(ResizeX)));
    DVM_CREATE_CUDA(cudaMalloc(&h->d_ytab, yt_n * sizeof(ResizeY)));
    DVM_CREATE_CUDA(cudaMalloc(&h->d_x4tab, x4_n * sizeof(ResizeX4)));
    DVM_CREATE_CUDA(cudaMalloc(&h->d_pattern, sizeof(kPatternHost)));
    DVM_CREATE_CUDA(cudaMemcpy(h->d_pattern, kPatternHost, sizeof(kPatternHost), cudaMemcpyHostToDevice));
    OrbBuffers& b = h->buf;
    memset(&b, 0, sizeof(b));
    DVM_CREATE_CUDA(cudaMalloc(&b.cand, cand * sizeof(uint32_t)));
    DVM_CREATE_CUDA(cudaMalloc(&b.pnode, cand * sizeof(uint16_t)));
    DVM_CREATE_CUDA(cudaMalloc(&b.cand_count, 2 * kMaxLevels * sizeof(int)));
    DVM_CREATE_CUDA(cudaMemset(b.cand_count, 0, 2 * kMaxLevels * sizeof(int)));
    DVM_CREATE_CUDA(cudaMalloc(&b.sel_count, kMaxLevels * sizeof(int)));
    DVM_CREATE_CUDA(cudaMemset(b.sel_count, 0, kMaxLevels * sizeof(int)));
    DVM_CREATE_CUDA(cudaMalloc(&b.sel, h->max_kp * sizeof(uint32_t)));
    DVM_CREATE_CUDA(cudaMalloc(&b.work_kp, h->max_kp * sizeof(uint32_t)));
    DVM_CREATE_CUDA(cudaMalloc(&b.work_meta, h->max_kp * sizeof(uint32_t)));
    DVM_CREATE_CUDA(cudaMalloc(&b.ticket, sizeof(unsigned int)));
    DVM_CREATE_CUDA(cudaMemset(b.ticket, 0, sizeof(unsigned int)));
    h->out_bytes = 32 + (size_t)h->max_kp * (sizeof(dvm_keypoint) + 32);
    DVM_CREATE_CUDA(cudaMalloc(&h->d_out, h->out_bytes));
    DVM_CREATE_CUDA(cudaMemset(h->d_out, 0, h->out_bytes));
    DVM_CREATE_CUDA(cudaHostAlloc(&h->h_out, h->out_bytes, cudaHostAllocDefault));
    b.counts = (int*)h->d_out;
    b.status = (int*)(h->d_out + 16);
    b.out_kps = (dvm_keypoint*)(h->d_out + 32);
    b.out_desc = h->d_out + 32 + (size_t)h->max_kp * sizeof(dvm_keypoint);
    b.xtab = h->d_xtab;
    b.ytab = h->d_ytab;
    b.x4tab = h->d_x4tab;
    b.pattern = h->d_pattern;
#undef DVM_CREATE_CUDA
    *out = h;
    return DVM_OK;
}

int dvm_orb_clone(const dvm_orb* src, dvm_orb** out)
{
    DVM_REQUIRE(src != nullptr && out != nullptr, "null argument");
    return dvm_orb_create(out, src->device, src->nfeatures, (float)src->scaleFactor, src->nlevels, src->iniTh, src->minTh,
                          src->max_w, src->max_h);
}

void dvm_orb_destroy(dvm_orb* h) { free_all(h); }

int dvm_orb_tables(const dvm_orb* h, int* nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                   int* features_per_level)
{
    DVM_REQUIRE(h != nullptr, "null handle");
    if (nlevels) *nlevels = h->nlevels;
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
        if (features_per_level) features_per_level[i] = h->perLevel[i];
    }
    return DVM_OK;
}

int dvm_orb_max_keypoints(const dvm_orb* h) { return h ? h->max_kp : DVM_ERR_INVALID; }
int dvm_orb_max_keypoints_current(const dvm_orb* h)
{
    if (!h) return DVM_ERR_INVALID;
    return h->cur_w > 0 ? std::min(h->max_kp, (h->cfg.max_kp + 3) & ~3) : h->max_kp;
}

// enqueue the whole extractor on the handle's stream; level 0 must already be in place
static int enqueue_pipeline(dvm_orb* h, int lap0, int lap1)
{
    const OrbCfg& c = h->cfg;
    const bool prof = h->profiling;
    if (prof) DVM_CUDA(cudaEventRecord(h->ev[0], h->stream));
    if (h->fused_pyramid) launch_pyramid(c, h->buf, h->stream);
    else for (int l = 1; l < c.nlevels; l++) launch_resize_level(c, h->buf, l, const_cast<uint8_t*>(c.lv[l].img), h->stream);
    if (prof) DVM_CUDA(cudaEventRecord(h->ev[1], h->stream));
    if (c.lv[0].img != h->tmap0_img || c.lv[0].pitch != h->tmap0_pitch) {   // level 0 may be the caller's image
        encode_level_tmap(c, 0, h->tmaps);
        h->tmap0_img = c.lv[0].img; h->tmap0_pitch = c.lv[0].pitch;
    }
    launch_fast_cells(c, h->buf, h->tmaps, h->stream);
    if (prof) DVM_CUDA(cudaEventRecord(h->ev[2], h->stream));
    launch_octree(c, h->buf, lap0, lap1, h->oct_smem, h->stream);
    if (prof) DVM_CUDA(cudaEventRecord(h->ev[3], h->stream));
    launch_describe(c, h->buf, h->stream);
    if (prof) DVM_CUDA(cudaEventRecord(h->ev[4], h->stream));
    DVM_CUDA(cudaGetLastError());
    if (prof) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            DVM_CUDA(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
            h->prof_ms[i] += ms;
        }
        h->prof_frames++;
    }
    return DVM_OK;
}

int dvm_orb_set_profiling(dvm_orb* h, int enable)
{
    DVM_REQUIRE(h != nullptr, "null handle");
    DVM_CUDA(cudaSetDevice(h->device));
    if (enable && !h->ev[0])
        for (auto& e : h->ev) DVM_CUDA(cudaEventCreate(&e));
    h->profiling = enable != 0;
    return DVM_OK;
}

int dvm_orb_get_profile(dvm_orb* h, int* n_frames, float* stage_ms_sum)
{
    DVM_REQUIRE(h != nullptr && n_frames != nullptr && stage_ms_sum != nullptr, "null argument");
    *n_frames = h->prof_frames;
    for (int i = 0; i < 4; i++) { stage_ms_sum[i] = h->prof_ms[i]; h->prof_ms[i] = 0; }
    h->prof_frames = 0;
    return DVM_OK;
}

// where the next extract leaves its result: the handle's own block (default) or caller-provided HBM
static void set_outputs(dvm_orb* h, dvm_keypoint* kps, uint8_t* desc, int32_t* counts)
{
    h->buf.counts = counts ? counts : (int*)h->d_out;
    h->buf.out_kps = kps ? kps : (dvm_keypoint*)(h->d_out + 32);
    h->buf.out_desc = desc ? desc : h->d_out + 32 + (size_t)h->max_kp * sizeof(dvm_keypoint);
}

int dvm_orb_extract_device_to(dvm_orb* h, const uint8_t* gray_dev, int width, int height, int stride, int lap0, int lap1,
                              dvm_keypoint* kps_dev, uint8_t* desc_dev, int32_t* counts_dev)
{
    DVM_REQUIRE(h != nullptr && gray_dev != nullptr, "null argument");
    DVM_REQUIRE(width > 0 && height > 0 && stride >= width, "bad image geometry");
    DVM_REQUIRE((kps_dev == nullptr) == (desc_dev == nullptr) && (kps_dev == nullptr) == (counts_dev == nullptr),
                "kps_dev, desc_dev and counts_dev go together");
    DVM_CUDA(cudaSetDevice(h->device));
    int rc = configure(h, width, height);
    if (rc != DVM_OK) return rc;
    h->cfg.lv[0].img = gray_dev; // level 0 is the caller's image, read in place
    h->cfg.lv[0].pitch = stride;
    h->level0_aliased = true;
    set_outputs(h, kps_dev, desc_dev, counts_dev);
    return enqueue_pipeline(h, lap0, lap1);
}

int dvm_orb_extract_device(dvm_orb* h, const uint8_t* gray_dev, int width, int height, int stride, int lap0, int lap1)
{
    return dvm_orb_extract_device_to(h, gray_dev, width, height, stride, lap0, lap1, nullptr, nullptr, nullptr);
}

int dvm_orb_sync(dvm_orb* h)
{
    DVM_REQUIRE(h != nullptr, "null handle");
    DVM_CUDA(cudaSetDevice(h->device));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    return DVM_OK;
}

int dvm_orb_result_device(const dvm_orb* h, const dvm_keypoint** kps_dev, const uint8_t** desc_dev,
                          const int32_t** counts_dev)
{
    DVM_REQUIRE(h != nullptr, "null handle");
    if (kps_dev) *kps_dev = h->buf.out_kps;
    if (desc_dev) *desc_dev = h->buf.out_desc;
    if (counts_dev) *counts_dev = h->buf.counts;
    return DVM_OK;
}

void* dvm_orb_stream(const dvm_orb* h) { return h ? (void*)h->stream : nullptr; }

int dvm_orb_extract(dvm_orb* h, const uint8_t* gray, int width, int height, int stride, int lap0, int lap1,
                    dvm_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index)
{
    DVM_REQUIRE(h != nullptr && n_out != nullptr && mono_index != nullptr, "null argument");
    *n_out = 0;
    *mono_index = -1;
    if (gray == nullptr || width <= 0 || height <= 0) return DVM_OK; // `if (_image.empty()) return -1;`
    DVM_REQUIRE(stride >= width, "stride smaller than width");
    DVM_REQUIRE(kps != nullptr && desc != nullptr && cap >= 0, "null output buffers");
    DVM_CUDA(cudaSetDevice(h->device));
    int rc = configure(h, width, height);
    if (rc != DVM_OK) return rc;
    OrbLevel& L0 = h->cfg.lv[0];
    L0.img = h->d_pyr + h->lvl_off[0];
    L0.pitch = (int)align_up(L0.w, 128);
    h->level0_aliased = false;
    set_outputs(h, nullptr, nullptr, nullptr);
    DVM_CUDA(cudaMemcpy2DAsync(h->d_pyr + h->lvl_off[0], L0.pitch, gray, stride, width, height, cudaMemcpyHostToDevice,
                               h->stream));
    rc = enqueue_pipeline(h, lap0, lap1);
    if (rc != DVM_OK) return rc;
    // one device->host transfer: {counts, status, keypoints, descriptors} of the quota-bounded block
    DVM_CUDA(cudaMemcpyAsync(h->h_out, h->d_out, h->out_bytes, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    const int* counts = (const int*)h->h_out;
    const int status = *(const int*)(h->h_out + 16);
    if (status != 0) {
        cudaMemsetAsync(h->buf.status, 0, sizeof(int), h->stream);
        cudaMemsetAsync(h->buf.cand_count, 0, kMaxLevels * sizeof(int), h->stream);
        set_error("extractor overflow (status bits %d): %s", status,
                  (status & 1) ? "more FAST candidates than the per-level buffer holds" : "octree node list overflow");
        return DVM_ERR_CAPACITY;
    }
    const int n = counts[0];
    if (n > cap) {
        set_error("caller buffers hold %d keypoints but %d were extracted", cap, n);
        return DVM_ERR_CAPACITY;
    }
    memcpy(kps, h->h_out + 32, (size_t)n * sizeof(dvm_keypoint));
    memcpy(desc, h->h_out + 32 + (size_t)h->max_kp * sizeof(dvm_keypoint), (size_t)n * 32);
    *n_out = n;
    *mono_index = counts[1];
    return DVM_OK;
}

// ---- stage read-back for the parity tests ----
int dvm_orb_debug_level_size(const dvm_orb* h, int level, int* w, int* hgt)
{
    DVM_REQUIRE(h != nullptr && h->cur_w > 0 && level >= 0 && level < h->nlevels, "bad level / no extract yet");
    *w = h->cfg.lv[level].w;
    *hgt = h->cfg.lv[level].h;
    return DVM_OK;
}

int dvm_orb_debug_level_image(dvm_orb* h, int level, int blurred, uint8_t* out)
{
    DVM_REQUIRE(h != nullptr && out != nullptr && h->cur_w > 0 && level >= 0 && level < h->nlevels, "bad argument");
    DVM_CUDA(cudaSetDevice(h->device));
    const OrbLevel& L = h->cfg.lv[level];
    const uint8_t* src = L.img;
    int pitch = L.pitch;
    if (blurred) {
        launch_blur_level_debug(L.img, L.w, L.h, L.pitch, h->d_dbg, (int)align_up(L.w, 128), h->stream);
        src = h->d_dbg;
        pitch = (int)align_up(L.w, 128);
    }
    DVM_CUDA(cudaMemcpy2DAsync(out, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    return DVM_OK;
}

int dvm_orb_debug_level_keypoints(dvm_orb* h, int level, int which, int* xs, int* ys, int* responses, int cap, int* n_out)
{
    DVM_REQUIRE(h != nullptr && n_out != nullptr && h->cur_w > 0 && level >= 0 && level < h->nlevels, "bad argument");
    DVM_CUDA(cudaSetDevice(h->device));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    const OrbLevel& L = h->cfg.lv[level];
    int n = 0;
    std::vector<uint32_t> tmp;
    if (which == 0) {
        DVM_CUDA(cudaMemcpy(&n, h->buf.cand_count + kMaxLevels + level, sizeof(int), cudaMemcpyDeviceToHost));
        n = std::min(n, L.cand_cap);
        tmp.resize(std::max(n, 1));
        DVM_CUDA(cudaMemcpy(tmp.data(), h->buf.cand + L.cand_off, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } else {
        DVM_CUDA(cudaMemcpy(&n, h->buf.sel_count + level, sizeof(int), cudaMemcpyDeviceToHost));
        tmp.resize(std::max(n, 1));
        DVM_CUDA(cudaMemcpy(tmp.data(), h->buf.sel + L.sel_off, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    *n_out = n;
    for (int i = 0; i < n && i < cap; i++) {
        if (xs) xs[i] = kp_x(tmp[i]) + kBorder;
        if (ys) ys[i] = kp_y(tmp[i]) + kBorder;
        if (responses) responses[i] = kp_resp(tmp[i]);
    }
    return DVM_OK;
}

} // extern "C"
