// Stand-ins for the sibling catkin package `visual_odometry` (vro), which is not part of the reference repository
// (CMakeLists.txt:22-23): only the members the reference's graph wrapper and offline drivers touch.
// The feature front end itself (extraction, RANSAC matching) is out of scope; these classes carry data, they do not match.
#pragma once
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include "opencv2/opencv.hpp"
struct LoadedEdge3D {          // rgbdslam's edge record: used as mr.edge.{id1,id2,transform,informationMatrix}
  int id1 = -1, id2 = -1;
  Eigen::Isometry3d transform;
  Eigen::Matrix<double, 6, 6> informationMatrix;
  LoadedEdge3D() : informationMatrix(Eigen::Matrix<double, 6, 6>::Identity()) {}
};
class MatchingResult {
 public:
  std::vector<cv::DMatch> inlier_matches;
  std::vector<cv::DMatch> all_matches;
  LoadedEdge3D edge;
  float rmse = 0;
  Eigen::Matrix4f final_trafo;
  Eigen::Matrix4f icp_trafo;
  unsigned int inlier_points = 0, outlier_points = 0, occluded_points = 0, all_points = 0;
  bool succeed_match = false;
  MatchingResult() : final_trafo(Eigen::Matrix4f::Identity()), icp_trafo(Eigen::Matrix4f::Identity()) { edge.transform.setIdentity(); }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
