// orb_extractor_adapter.cc -- compiled INSTEAD of O3/src/ORBextractor.cc (see orb_extractor_adapter.h).
#include "orb_extractor_adapter.h"

#include <cassert>
#include <cstring>

namespace ORB_SLAM3 {

ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
    : orb_(new dvm_host::OrbHandle), levels_(nlevels), scale_(scaleFactor)
{
    // the tables of O3/src/ORBextractor.cc:282-322 are computed by the library with the same float arithmetic
    dvm_host::check(dvm_orb_create(&orb_->h, dvm_host::device_from_env(), nfeatures, scaleFactor, nlevels, iniThFAST,
                                   minThFAST, dvm_host::max_width(), dvm_host::max_height()),
                    "ORBextractor::ORBextractor");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    int n = 0;
    dvm_host::check(dvm_orb_tables(orb_->h, &n, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                   mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()),
                    "ORBextractor::ORBextractor");
    assert(n == nlevels);
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() = default;

int ORBextractor::operator()(cv::InputArray image, cv::InputArray /*mask: ignored by the reference too*/,
                             std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors,
                             std::vector<int>& vLappingArea)
{
    if (image.empty()) return -1;                      // O3/src/ORBextractor.cc:879-880
    cv::Mat im = image.getMat();
    assert(im.type() == CV_8UC1);                      // :883
    const int cap = dvm_orb_max_keypoints(orb_->h);
    keypoints.resize(static_cast<size_t>(cap));
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0, mono = 0;
    dvm_host::check(dvm_orb_extract(orb_->h, im.data, im.cols, im.rows, static_cast<int>(im.step), vLappingArea[0],
                                    vLappingArea[1], reinterpret_cast<dvm_keypoint*>(keypoints.data()), desc.data, cap,
                                    &n, &mono),
                    "ORBextractor::operator()");
    keypoints.resize(static_cast<size_t>(n));
    if (n == 0) {
        descriptors.release();                         // :897-898
    } else {
        descriptors.create(n, 32, CV_8U);              // :900
        cv::Mat out = descriptors.getMat();
        for (int i = 0; i < n; i++) std::memcpy(out.ptr(i), desc.ptr(i), 32);
    }
    return mono;
}

void ORBextractor::FillImagePyramid()
{
    for (int l = 0; l < levels_; l++) {
        int w = 0, h = 0;
        dvm_host::check(dvm_orb_debug_level_size(orb_->h, l, &w, &h), "ORBextractor::FillImagePyramid");
        if (w <= 0 || h <= 0) continue;
        mvImagePyramid[l].create(h, w, CV_8UC1);
        cv::Mat tight(h, w, CV_8UC1);
        dvm_host::check(dvm_orb_debug_level_image(orb_->h, l, 0, tight.data), "ORBextractor::FillImagePyramid");
        tight.copyTo(mvImagePyramid[l]);
    }
}

} // namespace ORB_SLAM3
