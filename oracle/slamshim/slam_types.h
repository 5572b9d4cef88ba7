/* oracle/slamshim/slam_types.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-ins for ORB_SLAM3::Frame / KeyFrame / MapPoint / GeometricCamera / Pinhole with the member NAMES, types and
 * default arguments of O3/include/{Frame,KeyFrame,MapPoint}.h and O3/include/CameraModels/{GeometricCamera,Pinhole}.h
 * -- data members and trivial accessors only.  Every member function with arithmetic or control flow in it
 * (GetFeaturesInArea, PosInGrid, AssignFeaturesToGrid, isInFrustum, UpdatePoseMatrices, IsInImage, PredictScale,
 * Get{Min,Max}DistanceInvariance, Pinhole::project, Pinhole::epipolarConstrain) is NOT written here: its body is
 * taken from the reference's own .cc file at build time (oracle/slamshim/extract_ref.py -> oracle/_ref/gen/*.inc)
 * and compiled against these declarations together with the reference's unmodified ORBmatcher.cc. */
#ifndef DVM_SLAMSHIM_SLAM_TYPES_H
#define DVM_SLAMSHIM_SLAM_TYPES_H
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>
#include "sophus/sim3.hpp"
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

using namespace std;   /* the reference's headers do this too (Frame.h) and its .cc bodies rely on it */

class Frame;
class KeyFrame;
class MapPoint;

class GeometricCamera {
public:
    virtual ~GeometricCamera() { }
    virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    virtual bool epipolarConstrain(GeometricCamera* otherCamera, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                                   const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12, const float sigmaLevel,
                                   const float unc) = 0;
    float getParameter(const int i) { return mvParameters[i]; }
    std::vector<float> mvParameters;
};

class Pinhole : public GeometricCamera {
public:
    Pinhole(float fx, float fy, float cx, float cy) { mvParameters = { fx, fy, cx, cy }; }
    Eigen::Vector2f project(const Eigen::Vector3f& v3D);                                  /* Pinhole.cpp (extracted) */
    Eigen::Matrix3f toK_()
    {   /* Pinhole.cpp: K << fx, 0, cx, 0, fy, cy, 0, 0, 1 */
        Eigen::Matrix3f K;
        K(0, 0) = mvParameters[0]; K(0, 1) = 0.f; K(0, 2) = mvParameters[2];
        K(1, 0) = 0.f; K(1, 1) = mvParameters[1]; K(1, 2) = mvParameters[3];
        K(2, 0) = 0.f; K(2, 1) = 0.f; K(2, 2) = 1.f;
        return K;
    }
    bool epipolarConstrain(GeometricCamera* pCamera2, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2,
                           const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12, const float sigmaLevel,
                           const float unc);                                                /* Pinhole.cpp (extracted) */
};

class MapPoint {
public:
    /* --- data the glue fills --- */
    Eigen::Vector3f mWorldPos, mNormalVector;
    cv::Mat mDescriptor;
    float mfMinDistance = 0, mfMaxDistance = 0;
    int nObs = 0;
    bool mbBad = false;
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    MapPoint* mpReplaced = nullptr;
    int mnFlatIndex = -1;                /* index of this point in the caller's flat arrays */
    std::mutex mMutexPos;
    /* --- members of the reference class --- */
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF)
    {
        const auto it = mObservations.find(pKF);
        return it == mObservations.end() ? std::tuple<int, int>(-1, -1) : it->second;
    }
    void AddObservation(KeyFrame* pKF, int idx) { mObservations[pKF] = std::tuple<int, int>(idx, -1); nObs++; }
    void Replace(MapPoint* pMP) { mpReplaced = pMP; mbBad = true; }
    float GetMinDistanceInvariance();                                                       /* MapPoint.cc (extracted) */
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);
};

class Frame {
public:
    Frame() { }
    int N = 0;
    int Nleft = -1, Nright = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    float mbf = 0, mb = 0;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    int mnScaleLevels = 0;
    float mfScaleFactor = 0, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    Sophus::SE3<float> mTcw;
    Eigen::Matrix<float, 3, 3> mRwc;
    Eigen::Matrix<float, 3, 1> mOw;
    Eigen::Matrix<float, 3, 3> mRcw;
    Eigen::Matrix<float, 3, 1> mtcw;
    Sophus::SE3f mTrl;
    void SetPose(const Sophus::SE3<float>& Tcw) { mTcw = Tcw; UpdatePoseMatrices(); }       /* Frame.cc:518-524 */
    inline Sophus::SE3<float> GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    inline Eigen::Vector3f GetCameraCenter() { return mOw; }
    /* bodies from Frame.cc (extracted) */
    void UpdatePoseMatrices();
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    bool isInFrustum(MapPoint* pMP, float viewingCosLimit);
    bool isInFrustumChecks(MapPoint*, float, bool = false) { return false; }                /* stereo-fisheye only */
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                     const int maxLevel = -1, const bool bRight = false) const;
};

class KeyFrame {
public:
    KeyFrame(int gridCols, int gridRows, float wInv, float hInv, int nLevels, float logScale, int minX, int minY, int maxX, int maxY)
        : mnGridCols(gridCols), mnGridRows(gridRows), mfGridElementWidthInv(wInv), mfGridElementHeightInv(hInv),
          mnScaleLevels(nLevels), mfLogScaleFactor(logScale), mnMinX(minX), mnMinY(minY), mnMaxX(maxX), mnMaxY(maxY) { }
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0;
    int N = 0;
    int NLeft = -1, NRight = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    const int mnGridCols, mnGridRows;
    const float mfGridElementWidthInv, mfGridElementHeightInv;
    const int mnScaleLevels;
    const float mfLogScaleFactor;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    const int mnMinX, mnMinY, mnMaxX, mnMaxY;            /* ints in the reference (O3/include/KeyFrame.h:406-409) */
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<std::vector<std::vector<size_t>>> mGrid, mGridRight;
    Sophus::SE3f mTcw, mTwc;
    Eigen::Vector3f mOw;
    /* KeyFrame.cc:97-113 SetPose: mTcw = Tcw; mTwc = mTcw.inverse(); Rwc/Ow from mTwc */
    void SetPose(const Sophus::SE3f& Tcw) { mTcw = Tcw; mTwc = mTcw.inverse(); mOw = mTwc.translation(); }
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTwc; }
    Eigen::Vector3f GetCameraCenter() { return mOw; }
    Sophus::SE3f GetRightPose() { return mTcw; }
    Sophus::SE3f GetRightPoseInverse() { return mTwc; }
    Eigen::Vector3f GetRightCameraCenter() { return mOw; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints()
    {
        std::set<MapPoint*> s;
        for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
        return s;
    }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    /* bodies from KeyFrame.cc (extracted) */
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
};

} // namespace ORB_SLAM3
#endif
