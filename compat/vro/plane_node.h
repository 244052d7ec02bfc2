// stand-in for the sibling package `plane` (see plane.h): CPlaneNode, the planes extracted from one depth frame.
#pragma once
#include <vector>
#include "plane.h"
#include "cam_model.h"
#include "opencv2/opencv.hpp"
class CPlaneNode {
 public:
  std::vector<CPlane*> mv_planes;
  std::vector<std::vector<int> > mv_indices;     // pixel indices of every plane
  std::vector<int> mv_landmark_id;               // landmark (L) id per plane, -1 = not associated yet
  cv::Mat m_dpt;                                 // the depth frame the planes were extracted from
  CPlaneNode() {}
  ~CPlaneNode() { for (CPlane* p : mv_planes) delete p; }
  // plane segmentation of a depth frame / a point cloud: front end, not available here (no planes found)
  int extractPlanes(cv::Mat&, cv::Mat&, CamModel*) { return 0; }
  int extractPlanes(CloudPtr&, CamModel*) { return 0; }
  bool mergeOverlappedPlanes(int) { return false; }
  bool empty() const { return m_dpt.empty(); }
  void setDpt(cv::Mat& d) { m_dpt = d.clone(); }
};
