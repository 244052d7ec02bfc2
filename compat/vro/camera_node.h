// vro stand-in (see matching_result.h): CCameraNode, one keyframe of the feature front end.
#pragma once
#include <map>
#include <vector>
#include <Eigen/Core>
#include <Eigen/StdVector>
#include "opencv2/opencv.hpp"
#include "matching_result.h"
#include "cam_model.h"
typedef std::vector<Eigen::Vector4f, Eigen::aligned_allocator<Eigen::Vector4f> > std_vector_of_eigen_vector4f;
class CCameraNode {
 public:
  int m_id = -1;        // graph id
  int m_seq_id = -1;    // frame sequence id
  std::vector<cv::KeyPoint> m_feature_loc_2d;
  std_vector_of_eigen_vector4f m_feature_loc_3d;
  cv::Mat m_feature_descriptors;
  CCameraNode() {}
  virtual ~CCameraNode() {}
  // RANSAC feature matching against an older node: front end, not available here -- reports "no match"
  virtual MatchingResult matchNodePair(CCameraNode* /*older*/) { return MatchingResult(); }
  static void set_cam_cov(const CamModel&) {}
  // covariance of the VRO estimate from the inlier matches (front-end numerics): not available here
  template <class Helper, class Cov> void computeCov(CCameraNode* /*old*/, std::vector<cv::DMatch>& /*inliers*/, Helper /*h*/, Cov& cov) { cov.setIdentity(); }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
Eigen::Matrix4f getTransformFromMatches(const CCameraNode* newer, const CCameraNode* older, const std::vector<cv::DMatch>& matches);
inline Eigen::Matrix4f getTransformFromMatches(const CCameraNode*, const CCameraNode*, const std::vector<cv::DMatch>&) { return Eigen::Matrix4f::Identity(); }
