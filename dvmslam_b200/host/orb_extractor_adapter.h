// orb_extractor_adapter.h -- ORB_SLAM3::ORBextractor with the reference's public signature
// (O3/include/ORBextractor.h:44-96) on top of dvm_orb_* .  Include it INSTEAD of the reference's
// ORBextractor.h and compile orb_extractor_adapter.cc INSTEAD of O3/src/ORBextractor.cc.
//
// Callers and what they touch (all preserved):
//   Tracking::Tracking / ParseORBParamFile   new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST,
//                                            fMinThFAST), twice per agent (O3/src/Tracking.cc:575-581)
//   Frame::ExtractORB                        monoLeft = (*mpORBextractorLeft)(im, cv::Mat(), mvKeys,
//                                            mDescriptors, vLapping)          (O3/src/Frame.cc:410-417)
//   Frame::Frame                             GetLevels / GetScaleFactor(s) / GetInverseScaleFactors /
//                                            GetScaleSigmaSquares / GetInverseScaleSigmaSquares (Frame.cc:399-405)
//   mvImagePyramid                           public member, read by the stereo matcher only; filled on demand
//                                            by FillImagePyramid()
#pragma once
#include <memory>
#include <opencv2/opencv.hpp>
#include <vector>

#include "dvm_host.h"

static_assert(sizeof(cv::KeyPoint) == sizeof(dvm_keypoint), "cv::KeyPoint must have dvm_keypoint's 28-byte layout");

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // keypoints in the reference's output order, descriptors CV_8U [N x 32]; returns monoIndex (-1: empty image)
    int operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                   cv::OutputArray descriptors, std::vector<int>& vLappingArea);

    int GetLevels() { return levels_; }
    float GetScaleFactor() { return static_cast<float>(scale_); }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;
    // copies the pyramid of the last extracted image out of HBM into mvImagePyramid (the mono path never reads it)
    void FillImagePyramid();

    // the C handle, for the device-resident paths (dvm_frame_construct_device, dvm_tracker_create)
    dvm_orb* handle() const { return orb_->h; }

protected:
    std::unique_ptr<dvm_host::OrbHandle> orb_;
    int levels_;
    double scale_;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

} // namespace ORB_SLAM3
