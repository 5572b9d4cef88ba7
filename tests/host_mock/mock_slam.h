// mock_slam.h -- TEST INFRASTRUCTURE.  Stand-ins for the reference's Frame / KeyFrame / MapPoint / Map /
// Sophus::SE3f types with exactly the members dvmslam_b200/host/*.h read or write (names and meaning as in
// O3/include/{Frame,KeyFrame,MapPoint,Map}.h), so the adapters can be compiled and driven without Eigen,
// Sophus, DBoW2 and the rest of ORB-SLAM3.
#pragma once
#include <algorithm>
#include <map>
#include <mutex>
#include <opencv2/opencv.hpp>
#include <set>
#include <tuple>
#include <vector>

namespace mock {

struct Vec3 {
    float v[3] = { 0, 0, 0 };
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
struct Mat3 {
    float m[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float operator()(int r, int c) const { return m[3 * r + c]; }
};
struct Quat {
    float q[4] = { 0, 0, 0, 1 };
    float& x() { return q[0]; } float& y() { return q[1]; } float& z() { return q[2]; } float& w() { return q[3]; }
    float x() const { return q[0]; } float y() const { return q[1]; } float z() const { return q[2]; } float w() const { return q[3]; }
};
// Sophus::SE3<float>: unit_quaternion(), translation(), rotationMatrix(), SE3(quaternion, translation)
struct SE3f {
    Quat q; Vec3 t;
    SE3f() { }
    SE3f(const Quat& q_, const Vec3& t_) : q(q_), t(t_) { }
    Quat unit_quaternion() const { return q; }
    Vec3 translation() const { return t; }
    Mat3 rotationMatrix() const
    {   // Eigen's Quaternion::toRotationMatrix
        const float x = q.q[0], y = q.q[1], z = q.q[2], w = q.q[3];
        const float tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x,
                    txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
        Mat3 R;
        R.m[0] = 1 - (tyy + tzz); R.m[1] = txy - twz; R.m[2] = txz + twy;
        R.m[3] = txy + twz; R.m[4] = 1 - (txx + tzz); R.m[5] = tyz - twx;
        R.m[6] = txz - twy; R.m[7] = tyz + twx; R.m[8] = 1 - (txx + tyy);
        return R;
    }
};

// Sophus::Sim3<float>: quaternion() (squared norm = scale), translation()
struct Sim3f {
    Quat q; Vec3 t;
    Quat quaternion() const { return q; }
    Vec3 translation() const { return t; }
};

// g2o::Sim3: rotation() (Eigen::Quaterniond), translation() (Vector3d), scale(), all with mutable access
struct QuatD {
    double q[4] = { 0, 0, 0, 1 };
    double& x() { return q[0]; } double& y() { return q[1]; } double& z() { return q[2]; } double& w() { return q[3]; }
};
struct Vec3D {
    double v[3] = { 0, 0, 0 };
    double& operator()(int i) { return v[i]; }
};
struct Sim3 {
    QuatD r; Vec3D t; double s = 1;
    QuatD& rotation() { return r; }
    Vec3D& translation() { return t; }
    double& scale() { return s; }
};
struct Mat77 {
    double m[49];
    Mat77() { for (double& x : m) x = -1; }
    void setZero() { for (double& x : m) x = 0; }
};

struct Map;
struct KeyFrame;

struct MapPoint {
    static std::mutex mGlobalMutex;
    unsigned long mnId = 0;
    Vec3 pos;
    cv::Mat desc = cv::Mat(1, 32, CV_8U);
    int nObs = 1;
    bool bad = false;
    Map* map = nullptr;
    // tracking fields written by Frame::isInFrustum
    bool mbTrackInView = false, mbTrackInViewR = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackViewCos = 0, mTrackDepth = 0;
    int mnTrackScaleLevel = 0;
    unsigned long mnBALocalForKF = ~0ul, mnBAGlobalForKF = 0, mnBALocalForMerge = ~0ul;
    Vec3 mPosGBA;
    std::map<KeyFrame*, std::tuple<int, int>> observations;
    int normalUpdates = 0;

    Vec3 normal;
    float minDistInv = 0, maxDistInv = 0;   // GetMin/MaxDistanceInvariance()
    MapPoint* replacedBy = nullptr;
    Vec3 GetNormal() const { return normal; }
    float GetMinDistanceInvariance() const { return minDistInv; }
    float GetMaxDistanceInvariance() const { return maxDistInv; }
    float minDist = 0, maxDist = 0;         // mfMinDistance / mfMaxDistance and the getters INTEGRATION.md adds
    float GetMinDistance() const { return minDist; }
    float GetMaxDistance() const { return maxDist; }
    bool IsInKeyFrame(KeyFrame* kf) const { return observations.count(kf) != 0; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* kf) const
    {
        const auto it = observations.find(kf);
        return it == observations.end() ? std::make_tuple(-1, -1) : it->second;
    }
    void AddObservation(KeyFrame* kf, int idx) { observations[kf] = std::make_tuple(idx, -1); nObs++; }
    void Replace(MapPoint* other) { replacedBy = other; bad = true; }
    Vec3 GetWorldPos() const { return pos; }
    void SetWorldPos(const Vec3& p) { pos = p; }
    cv::Mat GetDescriptor() const { return desc; }
    int Observations() const { return nObs; }
    bool isBad() const { return bad; }
    Map* GetMap() const { return map; }
    std::map<KeyFrame*, std::tuple<int, int>> GetObservations() const { return observations; }
    void EraseObservation(KeyFrame* kf) { observations.erase(kf); }
    void UpdateNormalAndDepth() { normalUpdates++; }
    // loop-closing bookkeeping read by OptimizeEssentialGraph
    unsigned long mnCorrectedByKF = ~0ul, mnCorrectedReference = 0;
    KeyFrame* refKF = nullptr;
    KeyFrame* GetReferenceKeyFrame() const { return refKF; }
};

typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;   // DBoW2::FeatureVector

struct Frame {
    static float fx, fy, cx, cy, mnMinX, mnMinY, mnMaxX, mnMaxY;
    long unsigned int mnId = 0;
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvScaleFactors, mvInvLevelSigma2;
    FeatureVector mFeatVec;
    SE3f mTcw;
    SE3f GetPose() const { return mTcw; }
    void SetPose(const SE3f& T) { mTcw = T; }
};

struct KeyFrame {
    float fx = 0, fy = 0, cx = 0, cy = 0;   // per-object constants in the reference (O3/include/KeyFrame.h)
    long unsigned int mnId = 0;
    unsigned long mnBALocalForKF = ~0ul, mnBAFixedForKF = ~0ul, mnBAGlobalForKF = 0, mnBALocalForMerge = ~0ul;
    SE3f mTcwGBA;
    int N = 0, NLeft = -1;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight, mvInvLevelSigma2, mvScaleFactors, mvLevelSigma2;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mapPoints;
    MapPoint* GetMapPoint(size_t i) const { return mapPoints[i]; }
    void AddMapPoint(MapPoint* mp, size_t i) { mapPoints[i] = mp; }
    std::vector<KeyFrame*> covisible;
    FeatureVector mFeatVec;
    SE3f Tcw;
    bool bad = false;
    Map* map = nullptr;
    std::vector<MapPoint*> GetMapPointMatches() const { return mapPoints; }
    std::set<MapPoint*> GetMapPoints() const
    {
        std::set<MapPoint*> s;
        for (MapPoint* p : mapPoints) if (p) s.insert(p);
        return s;
    }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() const { return covisible; }
    SE3f GetPose() const { return Tcw; }
    Mat3 GetRotation() const { return Tcw.rotationMatrix(); }
    Vec3 GetTranslation() const { return Tcw.t; }
    void SetPose(const SE3f& T) { Tcw = T; }
    bool isBad() const { return bad; }
    Map* GetMap() const { return map; }
    void EraseMapPointMatch(MapPoint* mp)
    {
        for (auto& p : mapPoints) if (p == mp) p = nullptr;
    }
    // spanning tree / loop edges / covisibility graph as OptimizeEssentialGraph reads them
    KeyFrame* parent = nullptr;
    std::set<KeyFrame*> children, loopEdges;
    std::map<KeyFrame*, int> weights;
    KeyFrame* GetParent() const { return parent; }
    std::set<KeyFrame*> GetLoopEdges() const { return loopEdges; }
    bool hasChild(KeyFrame* k) const { return children.count(k) != 0; }
    int GetWeight(KeyFrame* k) const { const auto it = weights.find(k); return it == weights.end() ? 0 : it->second; }
    std::vector<KeyFrame*> GetCovisiblesByWeight(int w) const
    {
        std::vector<KeyFrame*> v;
        for (const auto& kv : weights) if (kv.second >= w) v.push_back(kv.first);
        return v;
    }
};

struct Map {
    std::mutex mMutexMapUpdate;
    unsigned long initKFid = 0;
    int changeIndex = 0;
    unsigned long GetInitKFid() const { return initKFid; }
    KeyFrame* originKF = nullptr;
    KeyFrame* GetOriginKF() const { return originKF; }
    void IncreaseChangeIndex() { changeIndex++; }
    std::vector<KeyFrame*> keyframes;
    std::vector<MapPoint*> mappoints;
    std::vector<KeyFrame*> GetAllKeyFrames() const { return keyframes; }
    std::vector<MapPoint*> GetAllMapPoints() const { return mappoints; }
    unsigned long GetMaxKFid() const { unsigned long m = 0; for (auto* k : keyframes) m = std::max(m, (unsigned long)k->mnId); return m; }
};

} // namespace mock
