// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in DECLARATIONS of the ORB-SLAM2 classes that the matching functions touch
// (Frame, MapPoint, KeyFrame, ORBmatcher, DBoW2::FeatureVector), reduced to the members those functions read or write.
// The function BODIES are the reference's own: oracle/Makefile (target _ref) streams the line ranges of
// thirdparty/orb-slam2/src/ORBmatcher.cc and src/Frame.cc that hold them to the compiler behind this header, checking
// the first line of every range.  The real class headers drag in the whole SLAM system (threads, g2o, DBoW2, Pangolin),
// which cannot be built here.  Not part of the product.
#pragma once
#include <climits>
#include <cmath>
#include <cstdint>
#include <map>
#include <mutex>
#include <vector>

#include "pgo_opencv_shim.h"

#define FRAME_GRID_ROWS 48   // thirdparty/orb-slam2/include/Frame.h:37-38
#define FRAME_GRID_COLS 64

namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2

namespace ORB_SLAM2 {

class KeyFrame;

class MapPoint {
 public:
  cv::Mat mWorldPos, mDescriptor;
  std::mutex mMutexFeatures;
  std::map<KeyFrame*, size_t> mObservations;   // MapPoint.h: keyframe -> index of the observing feature
  void ComputeDistinctiveDescriptors();
  int nObs = 1;
  bool mbBad = false;
  // Tracking::SearchLocalPoints fills these through Frame::isInFrustum (MapPoint.h)
  float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1;
  bool mbTrackInView = false;
  int mnTrackScaleLevel = 0;
  cv::Mat GetWorldPos() { return mWorldPos.clone(); }
  cv::Mat GetDescriptor() { return mDescriptor.clone(); }
  int Observations() { return nObs; }
  bool isBad() { return mbBad; }
};

class Frame {
 public:
  cv::Mat mTcw;
  float fx = 1, fy = 1, cx = 0, cy = 0, mb = 0, mbf = 0;
  int N = 0;
  std::vector<cv::KeyPoint> mvKeys, mvKeysUndistorted;
  std::vector<float> mvuRight;
  cv::Mat mDescriptors;
  std::vector<MapPoint*> mvpMapPoints;
  std::vector<bool> mvbOutlier;
  std::vector<float> mvScaleFactors;
  DBoW2::FeatureVector mFeatVec;
  static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
  std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

  std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1,
                                        const int maxLevel = -1) const;
  bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
  void AssignFeaturesToGrid();
};

class KeyFrame {
 public:
  std::vector<MapPoint*> mvpMapPoints;
  DBoW2::FeatureVector mFeatVec;
  cv::Mat mDescriptors;
  std::vector<cv::KeyPoint> mvKeysUn;
  std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
  bool mbBad = false;
  bool isBad() { return mbBad; }
};

class ORBmatcher {   // thirdparty/orb-slam2/include/ORBmatcher.h:38-103, the members the compiled functions use
 public:
  ORBmatcher(float nnratio = 0.6, bool checkOri = true);
  static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
  int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
  int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
  int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
  int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                              int windowSize = 10);
  static const int TH_LOW;
  static const int TH_HIGH;
  static const int HISTO_LENGTH;

 protected:
  float RadiusByViewingCos(const float& viewCos);
  void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM2
