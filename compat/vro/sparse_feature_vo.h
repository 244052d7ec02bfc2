// vro stand-in (see matching_result.h): CSparseFeatureVO, the feature extractor of the front end (out of scope: no-ops).
#pragma once
#include <vector>
#include "camera_node.h"
#include "camera_node_ba.h"
#include "cam_model.h"
class CSparseFeatureVO {
 public:
  explicit CSparseFeatureVO(const CamModel& = CamModel()) {}
  void featureExtraction(cv::Mat&, cv::Mat&, float, CCameraNode&) {}
  template <class A, class B> void generatePointCloud(cv::Mat&, cv::Mat&, int, float, A&, B&) {}
};
