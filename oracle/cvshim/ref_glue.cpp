/* oracle/cvshim/ref_glue.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's own
 * ORB_SLAM3::ORBextractor (compiled from /root/reference, see oracle/Makefile target `ref`). */
#include "ORBextractor.h"
#include <cstring>

struct RefKp { float x, y, size, angle, response; int32_t octave, class_id; };

extern "C" {
void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    return new ORB_SLAM3::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_orb_destroy(void* h) { delete (ORB_SLAM3::ORBextractor*)h; }
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, int lap0, int lap1, RefKp* kps, uint8_t* desc,
                    int cap, int* mono_index)
{
    ORB_SLAM3::ORBextractor* e = (ORB_SLAM3::ORBextractor*)h;
    cv::Mat im(hgt, w, CV_8UC1);
    for (int y = 0; y < hgt; y++) memcpy(im.ptr(y), img + (size_t)y * stride, w);
    std::vector<cv::KeyPoint> k;
    cv::Mat d;
    std::vector<int> lap = { lap0, lap1 };
    int mono = (*e)(im, cv::Mat(), k, d, lap);
    if (mono_index) *mono_index = mono;
    if ((int)k.size() > cap) return -2;
    for (size_t i = 0; i < k.size(); i++) {
        kps[i] = { k[i].pt.x, k[i].pt.y, k[i].size, k[i].angle, k[i].response, k[i].octave, k[i].class_id };
        memcpy(desc + i * 32, d.ptr((int)i), 32);
    }
    return (int)k.size();
}
void ref_orb_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2)
{
    ORB_SLAM3::ORBextractor* e = (ORB_SLAM3::ORBextractor*)h;
    int n = e->GetLevels();
    for (int i = 0; i < n; i++) {
        scale[i] = e->GetScaleFactors()[i]; invScale[i] = e->GetInverseScaleFactors()[i];
        sigma2[i] = e->GetScaleSigmaSquares()[i]; invSigma2[i] = e->GetInverseScaleSigmaSquares()[i];
    }
}
}
