// vro stand-in (see matching_result.h): CCameraNodeBA, a keyframe that remembers the landmark id of every feature.
#pragma once
#include "camera_node.h"
class CCameraNodeBA : public CCameraNode {
 public:
  std::vector<int> mv_feature_qid;      // landmark (Q) id per feature, -1 = not assigned yet
  CCameraNodeBA() {}
  virtual ~CCameraNodeBA() {}
  // feature matching under a predicted transform: front end, the caller overrides it (tests) or gets no matches
  virtual std::map<int, int> matchNodePairBA(CCameraNodeBA* /*older*/, Eigen::Matrix4f& /*Tji*/, CamModel* /*cam*/) { return std::map<int, int>(); }
};
