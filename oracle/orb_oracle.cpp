/*
 * oracle/orb_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's ORB extractor
 *   O3/src/ORBextractor.cc   (O3/ = /root/reference/src/slam_system/orb_slam3/)
 * on top of the OpenCV-primitive models in cvmodels.c.  Every function cites the
 * reference lines it follows.  Conventions that the reference leaves to its
 * toolchain and that this oracle fixes (see DESIGN.md "parity conventions"):
 *   - float expressions are evaluated op by op in IEEE float32, no FMA
 *     contraction (build with -ffp-contract=off);
 *   - cosf/sinf are restated explicitly (glibc 2.39 / ARM optimized-routines
 *     sincosf algorithm); orbo_sinf/orbo_cosf equal this box's libm on every
 *     float in [0, 2*pi] (exhaustive check: oracle/check_sincosf.c);
 *   - std::sort is libstdc++'s (the octree's tie order depends on it).
 *
 * Pinning: tests/test_oracle_cv2.py re-runs the whole pipeline with the cv2
 * 4.13.0 primitives substituted for cvmodels.c and demands identical output;
 * oracle/_ref (the reference's own ORBextractor.cc compiled against
 * oracle/cvshim) must agree too when /root/reference is present.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use this file.
 */
#include "cvmodels.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <vector>

namespace {

const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19; // ORBextractor.cc:71-73

const int8_t kPattern[256 * 4] = {
#include "../dvmslam_b200/csrc/orb_pattern.inc"
};

struct KeyPt { // cv::KeyPoint layout
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    uint8_t* row(int y) { return px.data() + (size_t)y * w; }
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
};

/* ---- glibc sinf/cosf restated (sysdeps/ieee754/flt-32/s_sincosf.h) ---- */
struct SinCosTab { double sign[4]; double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; };
const SinCosTab kSC[2] = {
    { { 1.0, -1.0, -1.0, 1.0 }, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2,
      0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
      0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 },
    { { 1.0, -1.0, -1.0, 1.0 }, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2,
      -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
      0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13 } };
inline float sc_poly(double x, double x2, const SinCosTab* p, int n)
{
    if ((n & 1) == 0) {
        double x3 = x * x2, s1 = p->s2 + x2 * p->s3, x7 = x3 * x2, s = x + x3 * p->s1;
        return (float)(s + x7 * s1);
    }
    double x4 = x2 * x2, c2 = p->c3 + x2 * p->c4, c1 = p->c1 + x2 * p->c2, x6 = x4 * x2, c = p->c0 + x2 * c1;
    return (float)(c + x6 * c2);
}
inline uint32_t top12(float v) { uint32_t u; memcpy(&u, &v, 4); return (u >> 20) & 0x7ff; }
inline float sc_eval(float y, int is_cos)
{
    double x = y;
    const SinCosTab* p = &kSC[0];
    if (top12(y) < top12(0x1.921FB6p-1f)) {
        if (top12(y) < top12(0x1p-12f)) return is_cos ? 1.0f : y;
        return sc_poly(x, x * x, p, is_cos);
    }
    double r = x * p->hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * p->hpi;
    double s = p->sign[n & 3];
    if (n & 2) p = &kSC[1];
    return sc_poly(x * s, x * x, p, n ^ is_cos);
}

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor; // ORBextractor.h:83 (double member initialised from a float argument)
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> perLevel;
    int umax[HALF_PATCH_SIZE + 1];
    // last-call intermediates (for stage-wise parity tests)
    std::vector<Image> pyr, blurred;
    std::vector<std::vector<KeyPt>> cand, sel;
};

/* ORBextractor::ORBextractor, ORBextractor.cc:282-339 */
Extractor* make_extractor(int nfeatures, float scaleFactorF, int nlevels, int iniTh, int minTh)
{
    Extractor* e = new Extractor;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniTh; e->minTh = minTh;
    e->scaleFactor = scaleFactorF;
    e->scale.resize(nlevels); e->sigma2.resize(nlevels);
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(nlevels); e->invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    e->perLevel.resize(nlevels);
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->perLevel[l] = cvm_round_f(nDesired);
        sum += e->perLevel[l];
        nDesired *= factor;
    }
    e->perLevel[nlevels - 1] = std::max(nfeatures - sum, 0);

    // circular patch row ends, :324-338
    int v, v0, vmax = (int)std::floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cvm_round_d(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

/* ORBextractor::ComputePyramid, ORBextractor.cc:957-976.  The +19 px reflect
 * border the reference adds is never read by later stages, so it is omitted. */
void compute_pyramid(Extractor* e, const uint8_t* img, int w, int h, int stride)
{
    e->pyr.assign(e->nlevels, Image());
    for (int l = 0; l < e->nlevels; l++) {
        float s = e->invScale[l];
        Image& I = e->pyr[l];
        I.w = cvm_round_f((float)w * s);
        I.h = cvm_round_f((float)h * s);
        I.px.resize((size_t)I.w * I.h);
        if (l == 0) {
            for (int y = 0; y < h; y++) memcpy(I.row(y), img + (size_t)y * stride, w);
        } else {
            const Image& P = e->pyr[l - 1];
            cvm_resize_linear_u8(P.px.data(), P.w, P.h, P.w, I.px.data(), I.w, I.h, I.w);
        }
    }
}

/* ---- DistributeOctTree, ORBextractor.cc:348-610 ---- */
struct Node {
    std::vector<int> keys; // indices into the candidate array, original order preserved
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
};

/* ExtractorNode::DivideNode :348-400 */
void divide(const Node& p, const std::vector<KeyPt>& pts, Node n[4])
{
    const int halfX = (int)std::ceil(static_cast<float>(p.URx - p.ULx) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(p.BRy - p.ULy) / 2);
    n[0].ULx = p.ULx; n[0].ULy = p.ULy;
    n[0].URx = p.ULx + halfX; n[0].URy = p.ULy;
    n[0].BLx = p.ULx; n[0].BLy = p.ULy + halfY;
    n[0].BRx = p.ULx + halfX; n[0].BRy = p.ULy + halfY;
    n[1].ULx = n[0].URx; n[1].ULy = n[0].URy;
    n[1].URx = p.URx; n[1].URy = p.URy;
    n[1].BLx = n[0].BRx; n[1].BLy = n[0].BRy;
    n[1].BRx = p.URx; n[1].BRy = p.ULy + halfY;
    n[2].ULx = n[0].BLx; n[2].ULy = n[0].BLy;
    n[2].URx = n[0].BRx; n[2].URy = n[0].BRy;
    n[2].BLx = p.BLx; n[2].BLy = p.BLy;
    n[2].BRx = n[0].BRx; n[2].BRy = p.BLy;
    n[3].ULx = n[2].URx; n[3].ULy = n[2].URy;
    n[3].URx = n[1].BRx; n[3].URy = n[1].BRy;
    n[3].BLx = n[2].BRx; n[3].BLy = n[2].BRy;
    n[3].BRx = p.BRx; n[3].BRy = p.BRy;
    for (int k : p.keys) {
        const KeyPt& kp = pts[k];
        if (kp.x < n[0].URx) {
            if (kp.y < n[0].BRy) n[0].keys.push_back(k);
            else n[2].keys.push_back(k);
        } else if (kp.y < n[0].BRy) n[1].keys.push_back(k);
        else n[3].keys.push_back(k);
    }
    for (int c = 0; c < 4; c++)
        if (n[c].keys.size() == 1) n[c].noMore = true;
}

typedef std::pair<int, Node*> SizeNode;
/* compareNodes :402-417 */
bool compare_nodes(SizeNode& a, SizeNode& b)
{
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ULx < b.second->ULx;
}

std::vector<KeyPt> distribute_octree(const std::vector<KeyPt>& pts, int minX, int maxX, int minY, int maxY, int N)
{
    std::vector<KeyPt> result;
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> L;
    std::vector<Node*> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        L.push_back(ni);
        ini[i] = &L.back();
    }
    for (size_t i = 0; i < pts.size(); i++) ini[(int)(pts[i].x / hX)]->keys.push_back((int)i);

    for (auto it = L.begin(); it != L.end();) {
        if (it->keys.size() == 1) { it->noMore = true; ++it; }
        else if (it->keys.empty()) it = L.erase(it);
        else ++it;
    }

    bool finish = false;
    std::vector<SizeNode> cands;
    auto push_children = [&](Node n[4], int* nToExpand) {
        for (int c = 0; c < 4; c++) {
            if (n[c].keys.size() > 0) {
                L.push_front(n[c]);
                if (n[c].keys.size() > 1) {
                    if (nToExpand) ++*nToExpand;
                    cands.push_back(std::make_pair((int)n[c].keys.size(), &L.front()));
                    L.front().lit = L.begin();
                }
            }
        }
    };
    while (!finish) {
        int prevSize = (int)L.size();
        int nToExpand = 0;
        cands.clear();
        for (auto it = L.begin(); it != L.end();) {
            if (it->noMore) { ++it; continue; }
            Node n[4];
            divide(*it, pts, n);
            push_children(n, &nToExpand);
            it = L.erase(it);
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) {
            finish = true;
        } else if (((int)L.size() + nToExpand * 3) > N) {
            while (!finish) {
                prevSize = (int)L.size();
                std::vector<SizeNode> prev = cands;
                cands.clear();
                std::sort(prev.begin(), prev.end(), compare_nodes);
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n[4];
                    divide(*prev[j].second, pts, n);
                    push_children(n, nullptr);
                    L.erase(prev[j].second->lit);
                    if ((int)L.size() >= N) break;
                }
                if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
            }
        }
    }
    for (auto& nd : L) { // best response per node, first maximum wins (:596-606)
        int best = nd.keys[0];
        float mx = pts[best].response;
        for (size_t k = 1; k < nd.keys.size(); k++)
            if (pts[nd.keys[k]].response > mx) { best = nd.keys[k]; mx = pts[best].response; }
        result.push_back(pts[best]);
    }
    return result;
}

/* IC_Angle, ORBextractor.cc:75-99 */
float ic_angle(const Image& im, float px, float py, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = im.row(cvm_round_f(py)) + cvm_round_f(px);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    int step = im.w;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * step], vm = center[u - v * step];
            v_sum += (vp - vm);
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    return cvm_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, ORBextractor.cc:101-143 */
const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
void orb_descriptor(const KeyPt& kp, const Image& im, uint8_t* desc)
{
    float angle = (float)kp.angle * factorPI;
    float a = sc_eval(angle, 1), b = sc_eval(angle, 0);
    const uint8_t* center = im.row(cvm_round_f(kp.y)) + cvm_round_f(kp.x);
    const int step = im.w;
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int k = 0; k < 8; k++, pat += 4) {
            int x0 = pat[0], y0 = pat[1], x1 = pat[2], y1 = pat[3];
            int t0 = center[cvm_round_f(x0 * b + y0 * a) * step + cvm_round_f(x0 * a - y0 * b)];
            int t1 = center[cvm_round_f(x1 * b + y1 * a) * step + cvm_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ORBextractor::ComputeKeyPointsOctTree, ORBextractor.cc:612-715 */
void compute_keypoints(Extractor* e)
{
    const float W = 35;
    e->cand.assign(e->nlevels, {});
    e->sel.assign(e->nlevels, {});
    std::vector<cvm_fast_kp> buf(1 << 16);
    for (int level = 0; level < e->nlevels; ++level) {
        const Image& im = e->pyr[level];
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
        std::vector<KeyPt>& toDistribute = e->cand[level];
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                const uint8_t* sub = im.row(y0) + x0;
                int n = cvm_fast_detect(sub, x1 - x0, y1 - y0, im.w, e->iniTh, buf.data(), (int)buf.size());
                if (n == 0) n = cvm_fast_detect(sub, x1 - x0, y1 - y0, im.w, e->minTh, buf.data(), (int)buf.size());
                for (int k = 0; k < n; k++) {
                    KeyPt kp;
                    kp.x = (float)buf[k].x + j * wCell;
                    kp.y = (float)buf[k].y + i * hCell;
                    kp.size = 7.f; kp.angle = -1.f; kp.response = (float)buf[k].response;
                    kp.octave = 0; kp.class_id = -1;
                    toDistribute.push_back(kp);
                }
            }
        }
        std::vector<KeyPt>& kps = e->sel[level];
        kps = distribute_octree(toDistribute, minBorderX, maxBorderX, minBorderY, maxBorderY, e->perLevel[level]);
        const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[level]);
        for (auto& kp : kps) {
            kp.x += minBorderX; kp.y += minBorderY;
            kp.octave = level;
            kp.size = (float)scaledPatchSize;
        }
    }
    for (int level = 0; level < e->nlevels; ++level)
        for (auto& kp : e->sel[level]) kp.angle = ic_angle(e->pyr[level], kp.x, kp.y, e->umax);
}

} // namespace

extern "C" {

void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    return make_extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orbo_destroy(void* h) { delete (Extractor*)h; }

void orbo_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* perLevel, int* umax)
{
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = e->scale[i];
        if (invScale) invScale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (invSigma2) invSigma2[i] = e->invSigma2[i];
        if (perLevel) perLevel[i] = e->perLevel[i];
    }
    if (umax) for (int i = 0; i <= HALF_PATCH_SIZE; i++) umax[i] = e->umax[i];
}

/* ORBextractor::operator(), ORBextractor.cc:876-955.  kps/desc must hold
 * `cap` entries; returns the keypoint count (or -1 on an empty image) and the
 * reference's return value (monoIndex) through mono_index. */
int orbo_extract(void* h, const uint8_t* img, int w, int hgt, int stride, int lap0, int lap1, KeyPt* kps,
                 uint8_t* desc, int cap, int* mono_index)
{
    Extractor* e = (Extractor*)h;
    if (!img || w <= 0 || hgt <= 0) return -1;
    compute_pyramid(e, img, w, hgt, stride);
    compute_keypoints(e);
    int nk = 0;
    for (int l = 0; l < e->nlevels; l++) nk += (int)e->sel[l].size();
    if (nk > cap) return -2;
    e->blurred.assign(e->nlevels, Image());
    int monoIndex = 0, stereoIndex = nk - 1;
    std::vector<uint8_t> d(32);
    for (int level = 0; level < e->nlevels; ++level) {
        std::vector<KeyPt>& v = e->sel[level];
        if (v.empty()) continue;
        const Image& src = e->pyr[level];
        Image& B = e->blurred[level];
        B.w = src.w; B.h = src.h; B.px.resize(src.px.size());
        cvm_gaussian7_u8(src.px.data(), src.w, src.h, src.w, B.px.data(), B.w);
        float scale = e->scale[level];
        for (auto& kp : v) {
            orb_descriptor(kp, B, d.data());
            KeyPt out = kp;
            if (level != 0) { out.x *= scale; out.y *= scale; }
            int dst;
            if (out.x >= lap0 && out.x <= lap1) dst = stereoIndex--;
            else dst = monoIndex++;
            kps[dst] = out;
            memcpy(desc + (size_t)dst * 32, d.data(), 32);
        }
    }
    if (mono_index) *mono_index = monoIndex;
    return nk;
}

/* ---- stage accessors (valid after orbo_extract) ---- */
void orbo_level_size(void* h, int level, int* w, int* hgt)
{
    Extractor* e = (Extractor*)h;
    *w = e->pyr[level].w; *hgt = e->pyr[level].h;
}
void orbo_level_image(void* h, int level, int blurred, uint8_t* out)
{
    Extractor* e = (Extractor*)h;
    const Image& I = blurred ? e->blurred[level] : e->pyr[level];
    if (!I.px.empty()) memcpy(out, I.px.data(), I.px.size());
}
int orbo_level_candidates(void* h, int level, KeyPt* out, int cap)
{
    Extractor* e = (Extractor*)h;
    int n = (int)e->cand[level].size();
    for (int i = 0; i < n && i < cap; i++) out[i] = e->cand[level][i];
    return n;
}
int orbo_level_selected(void* h, int level, KeyPt* out, int cap)
{
    Extractor* e = (Extractor*)h;
    int n = (int)e->sel[level].size();
    for (int i = 0; i < n && i < cap; i++) out[i] = e->sel[level][i];
    return n;
}

/* stand-alone pieces, so the cv2-backed pipeline in tests can reuse the logic */
int orbo_distribute(const KeyPt* pts, int n, int minX, int maxX, int minY, int maxY, int N, KeyPt* out)
{
    std::vector<KeyPt> v(pts, pts + n);
    std::vector<KeyPt> r = distribute_octree(v, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size(); i++) out[i] = r[i];
    return (int)r.size();
}
void orbo_ic_moments(const uint8_t* img, int w, int h, int x, int y, const int* umax, int* m01, int* m10)
{
    int a = 0, b = 0;
    const uint8_t* center = img + (size_t)y * w + x;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) b += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * w], vm = center[u - v * w];
            v_sum += (vp - vm);
            b += u * (vp + vm);
        }
        a += v * v_sum;
    }
    (void)h;
    *m01 = a; *m10 = b;
}
void orbo_descriptor(const uint8_t* blurred, int w, int h, float x, float y, float angle, uint8_t* desc)
{
    Image I; I.w = w; I.h = h; I.px.assign(blurred, blurred + (size_t)w * h);
    KeyPt kp; kp.x = x; kp.y = y; kp.angle = angle;
    orb_descriptor(kp, I, desc);
}
float orbo_sinf(float x) { return sc_eval(x, 0); }
float orbo_cosf(float x) { return sc_eval(x, 1); }

} // extern "C"
