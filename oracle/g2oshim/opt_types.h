/* oracle/g2oshim/opt_types.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-ins for ORB_SLAM3::Frame / KeyFrame / MapPoint / Map / GeometricCamera / Pinhole / LoopClosing / Optimizer with
 * the member NAMES and types that the optimisation functions of O3/src/Optimizer.cc use (O3/include/{Frame,KeyFrame,
 * MapPoint,Map,LoopClosing,Optimizer}.h, O3/include/CameraModels/{GeometricCamera,Pinhole}.h) -- data members and
 * trivial accessors only.  The functions under test themselves (Optimizer::PoseOptimization, LocalBundleAdjustment x2,
 * BundleAdjustment, OptimizeSim3, OptimizeEssentialGraph), the edge types of O3/src/OptimizableTypes.cpp and
 * Pinhole::project / projectJac are NOT written here: their bodies are taken from the reference's own files at build time
 * (oracle/g2oshim/extract_opt.py) and compiled, unmodified, against these declarations and the reference's vendored g2o. */
#ifndef DVM_G2OSHIM_OPT_TYPES_H
#define DVM_G2OSHIM_OPT_TYPES_H
#include <list>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include <Eigen/Dense>
#include <sophus/se3.hpp>
#include "Thirdparty/g2o/g2o/types/sim3.h"

namespace cv {
struct Point2f { float x = 0, y = 0; Point2f() { } Point2f(float x_, float y_) : x(x_), y(y_) { } };
struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    KeyPoint() { }
    KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) { }
};
} // namespace cv

namespace ORB_SLAM3 {
using namespace std;   /* the reference's headers do this too and its .cc bodies rely on it */

class Map;
class KeyFrame;
class MapPoint;

struct Verbose {
    enum eLevel { VERBOSITY_QUIET = 0, VERBOSITY_NORMAL = 1, VERBOSITY_VERBOSE = 2, VERBOSITY_VERY_VERBOSE = 3, VERBOSITY_DEBUG = 4 };
    static void PrintMess(std::string, eLevel) { }
};

class GeometricCamera {
public:
    virtual ~GeometricCamera() { }
    virtual Eigen::Vector2d project(const Eigen::Vector3d& v3D) = 0;
    virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D) = 0;
    std::vector<float> mvParameters;
    float getParameter(const int i) { return mvParameters[i]; }              /* GeometricCamera.h:96-100 */
    void setParameter(const float p, const size_t i) { mvParameters[i] = p; }
    size_t size() { return mvParameters.size(); }
};
class Pinhole : public GeometricCamera {
public:
    Pinhole(float fx, float fy, float cx, float cy) { mvParameters = { fx, fy, cx, cy }; }
    Eigen::Vector2d project(const Eigen::Vector3d& v3D);                 /* Pinhole.cpp (extracted) */
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d& v3D);  /* Pinhole.cpp (extracted) */
};

class MapPoint {
public:
    static std::mutex mGlobalMutex;
    long unsigned int mnId = 0;
    Eigen::Vector3f mWorldPos = Eigen::Vector3f::Zero();
    bool mbBad = false;
    Map* mpMap = nullptr;
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    KeyFrame* mpRefKF = nullptr;
    int nNormalUpdates = 0;
    /* members of the reference class that the functions write */
    long unsigned int mnBALocalForKF = 0, mnBAGlobalForKF = 0, mnBALocalForMerge = 0, mnCorrectedByKF = 0, mnCorrectedReference = 0;
    Eigen::Vector3f mPosGBA = Eigen::Vector3f::Zero();
    int mnTrackScaleLevel = 0;
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    void SetWorldPos(const Eigen::Vector3f& p) { mWorldPos = p; }
    bool isBad() { return mbBad; }
    Map* GetMap() { return mpMap; }
    std::map<KeyFrame*, std::tuple<int, int>> GetObservations() { return mObservations; }
    void EraseObservation(KeyFrame* pKF) { mObservations.erase(pKF); }
    void UpdateNormalAndDepth() { nNormalUpdates++; }
    KeyFrame* GetReferenceKeyFrame() { return mpRefKF; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame* pKF)
    {
        const auto it = mObservations.find(pKF);
        return it == mObservations.end() ? std::tuple<int, int>(-1, -1) : it->second;
    }
};

class Frame {
public:
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvInvLevelSigma2;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    Sophus::SE3<float> mTcw, mTrl;
    Sophus::SE3<float> GetPose() const { return mTcw; }
    void SetPose(const Sophus::SE3<float>& Tcw) { mTcw = Tcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
};

class KeyFrame {
public:
    long unsigned int mnId = 0;
    int N = 0, NLeft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight;
    std::vector<float> mvInvLevelSigma2;
    std::vector<MapPoint*> mvpMapPoints;
    GeometricCamera* mpCamera = nullptr;
    GeometricCamera* mpCamera2 = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0;
    bool mbBad = false, bImu = false;
    Map* mpMap = nullptr;
    KeyFrame* mPrevKF = nullptr;
    KeyFrame* mpParent = nullptr;
    std::set<KeyFrame*> mspChildrens, mspLoopEdges;
    std::vector<KeyFrame*> mvpOrderedConnectedKeyFrames;   /* covisible keyframes, best first */
    std::map<KeyFrame*, int> mConnectedKeyFrameWeights;
    Sophus::SE3f mTcw, mTcwGBA, mTrl;
    long unsigned int mnBALocalForKF = 0, mnBAFixedForKF = 0, mnBAGlobalForKF = 0, mnBALocalForMerge = 0;
    int nEraseCalls = 0;
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    void SetPose(const Sophus::SE3f& Tcw) { mTcw = Tcw; }
    Eigen::Matrix3f GetRotation() { return mTcw.rotationMatrix(); }
    Eigen::Vector3f GetTranslation() { return mTcw.translation(); }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    bool isBad() { return mbBad; }
    Map* GetMap() { return mpMap; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint*> GetMapPoints()
    {   /* KeyFrame.cc:326-337: the non-bad matches */
        std::set<MapPoint*> s;
        for (MapPoint* p : mvpMapPoints) if (p && !p->isBad()) s.insert(p);
        return s;
    }
    void EraseMapPointMatch(MapPoint* pMP)
    {   /* KeyFrame.cc:303-314 */
        const std::tuple<int, int> idx = pMP->GetIndexInKeyFrame(this);
        if (std::get<0>(idx) != -1) mvpMapPoints[std::get<0>(idx)] = nullptr;
        nEraseCalls++;
    }
    std::vector<KeyFrame*> GetVectorCovisibleKeyFrames() { return mvpOrderedConnectedKeyFrames; }
    std::vector<KeyFrame*> GetCovisiblesByWeight(const int& w)
    {   /* KeyFrame.cc:270-289: connected keyframes with weight >= w, best first */
        std::vector<KeyFrame*> out;
        for (KeyFrame* k : mvpOrderedConnectedKeyFrames) if (mConnectedKeyFrameWeights[k] >= w) out.push_back(k);
        return out;
    }
    int GetWeight(KeyFrame* pKF) { const auto it = mConnectedKeyFrameWeights.find(pKF); return it == mConnectedKeyFrameWeights.end() ? 0 : it->second; }
    KeyFrame* GetParent() { return mpParent; }
    bool hasChild(KeyFrame* pKF) { return mspChildrens.count(pKF) != 0; }
    std::set<KeyFrame*> GetLoopEdges() { return mspLoopEdges; }
};

class Map {
public:
    std::mutex mMutexMapUpdate;
    std::vector<KeyFrame*> mvpKeyFrames;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<KeyFrame*> mvpKeyFrameOrigins;
    long unsigned int mnInitKFid = 0, mnMaxKFid = 0;
    int mnMapChange = 0;
    bool mbIsInertial = false;
    std::set<long unsigned int> msOptKFs, msFixedKFs;
    std::vector<KeyFrame*> GetAllKeyFrames() { return mvpKeyFrames; }
    std::vector<MapPoint*> GetAllMapPoints() { return mvpMapPoints; }
    long unsigned int GetInitKFid() { return mnInitKFid; }
    long unsigned int GetMaxKFid() { return mnMaxKFid; }
    KeyFrame* GetOriginKF() { return mvpKeyFrameOrigins.empty() ? nullptr : mvpKeyFrameOrigins[0]; }
    bool IsInertial() { return mbIsInertial; }
    void IncreaseChangeIndex() { mnMapChange++; }
};

class LoopClosing {
public:
    typedef map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>, Eigen::aligned_allocator<std::pair<KeyFrame* const, g2o::Sim3>>> KeyFrameAndPose;
};

class Optimizer {
public:
    void static BundleAdjustment(const std::vector<KeyFrame*>& vpKF, const std::vector<MapPoint*>& vpMP, int nIterations = 5,
                                 bool* pbStopFlag = NULL, const unsigned long nLoopKF = 0, const bool bRobust = true);
    void static LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF, int& num_OptKF, int& num_MPs,
                                      int& num_edges);
    int static PoseOptimization(Frame* pFrame);
    void static OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
                                       const LoopClosing::KeyFrameAndPose& NonCorrectedSim3,
                                       const LoopClosing::KeyFrameAndPose& CorrectedSim3,
                                       const map<KeyFrame*, set<KeyFrame*>>& LoopConnections, const bool& bFixScale);
    static int OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12, const float th2,
                            const bool bFixScale, Eigen::Matrix<double, 7, 7>& mAcumHessian, const bool bAllPoints = false);
    void static LocalBundleAdjustment(KeyFrame* pMainKF, vector<KeyFrame*> vpAdjustKF, vector<KeyFrame*> vpFixedKF, bool* pbStopFlag);
};

} // namespace ORB_SLAM3
#endif
