// gtsam_graph.h -- host-side mirror of the reference's wrapper API for the hot path, over the C ABI.
//
// Same class / method names, argument meaning and error behaviour as the reference (paths relative to its root):
//   CGraphGT      gtsam/gtsam_graph.h:46-150      (graph builder + optimise triggers)
//   CImuBase      gtsam/imu_base.h:31-77          (IMU log + preintegration driver)
//   CImuVn100     gtsam/imu_vn100.h:19-37         (VN-100 noise spec + log reader)
//   MatchingResult / CCameraNode: the minimal member surface the wrapper touches; the real
//   classes live in the sibling `visual_odometry` package, which is not part of the reference repository.
// Only the members on the solver path are mirrored; front-end members (feature matching, plane segmentation,
// PLY/trajectory writers) are out of scope.
#pragma once
#include <array>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "gtsam_lite.h"

#define D2R(d) (((d) * M_PI) / 180.)
#define R2D(r) (((r) * 180.) / M_PI)
#ifndef ROS_INFO
#define ROS_INFO(...) do { fprintf(stdout, "[INFO] " __VA_ARGS__); fprintf(stdout, "\n"); } while (0)
#define ROS_WARN(...) do { fprintf(stderr, "[WARN] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { fprintf(stderr, "[ERROR] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#endif

// ---- sibling-package surface
struct MatchingEdge {
  int id1 = 0, id2 = 0;
  gtsam::Pose3 transform;               // Eigen::Isometry3d in the reference
  gtsam::Matrix6 informationMatrix;     // order [rot, trans] as GTSAM consumes it (gtsam_graph.cpp:676,689)
};
class MatchingResult {
 public:
  MatchingEdge edge;
  gtsam::Matrix4 final_trafo;
  bool succeed_match = true;
};
class CCameraNode {
 public:
  int m_id = -1, m_seq_id = -1;
  virtual ~CCameraNode() {}
};

// BA node / camera model: the members the BA builders touch (camera_node_ba.h, cam_model.h in the sibling packages).
// Feature matching itself (matchNodePairBA) is front end and out of scope: the base returns no matches, a caller that
// has correspondences overrides it.
struct KeyPoint2f { struct { float x = 0, y = 0; } pt; };
class CamModel {
 public:
  CamModel(double fx_, double fy_, double cx_, double cy_, double k1_ = 0, double k2_ = 0) : fx(fx_), fy(fy_), cx(cx_), cy(cy_), k1(k1_), k2(k2_) {}
  double fx, fy, cx, cy, k1, k2;
};
class CCameraNodeBA : public CCameraNode {
 public:
  std::vector<std::array<float, 4>> m_feature_loc_3d;    // camera-frame (x, y, z, 1)
  std::vector<KeyPoint2f> m_feature_loc_2d;              // pixel measurements
  std::vector<int> mv_feature_qid;                       // landmark id of every feature, -1 = none yet
  virtual std::map<int, int> matchNodePairBA(CCameraNodeBA* /*older*/, const gtsam::Matrix4& /*Tji*/, CamModel*) { return std::map<int, int>(); }
};

typedef enum { SUCC_KF, FAIL_NOT_KF, FAIL_KF } ADD_RET;   // gtsam/gtsam_graph.h:43

namespace CG {                                             // gtsam/color.h
typedef enum { RED = 0, GREEN, BLUE, PURPLE, WHITE, YELLOW, DARK } COLOR;
extern unsigned char g_color[][3];
}

class CGraphGT {
 public:
  CGraphGT();
  virtual ~CGraphGT();

  void firstNode(CCameraNode*, bool online = true);              // gtsam_graph.cpp:320-368
  void fakeOdoNode(CCameraNode*);                                // :697-722
  void optimizeGraph();                                          // :1779-1782
  void optimizeGraphBatch();                                     // :1784-1788
  void optimizeGraphIncremental();                               // :1768-1776
  bool addToGTSAM(MatchingResult&, bool set_estimate);           // :630-695
  bool addToGTSAM(gtsam::NavState&, int vid, bool add_pose);     // :613-628
  bool addToGTSAM(CCameraNodeBA* ni, CCameraNodeBA* nj, std::map<int, int>& matches, CamModel* pcam);   // :370-448 (multi-frame BA)
  bool bundleAdjust(MatchingResult* pm, CCameraNode* pNewNode, CamModel* pcam);                         // :500-610 (two-view BA -> edge)
  // :1118-1298 with the CPlane argument replaced by what the factor needs: plane (nx,ny,nz,d) in the IMU frame
  // and its 3x3 covariance in the OrientedPlane3 tangent (S_upj after the reference's conditioning)
  bool addPlaneFactor(const gtsam::Vector4& plane_imu, const gtsam::Matrix3& S_upj, int pose_id, int landmark);
  double error();                                                // :173-176
  size_t camnodeSize() { return m_graph_map.size(); }

  int m_sequence_id = 0;
  int m_vertex_id = 0;
  std::map<int, CCameraNode*> m_graph_map;                       // graph owns the nodes (dtor deletes, :152-156)
  gtsam::NonlinearFactorGraph* mp_fac_graph;
  gtsam::Values* mp_node_values;
  void setWorld2Original(double p);                              // :178-208
  void setCamera2IMU(double p);                                  // :218-254
  void setCamera2IMUTranslation(double px, double py, double pz);// :210-216
  gtsam::Pose3* mp_w2o;
  gtsam::Pose3* mp_u2c;
  gtsam::imuBias::ConstantBias* mp_prev_bias;
  gtsam::NavState* mp_prev_state;

  void printVROResult(std::ostream& ouf, MatchingResult& m);     // :1560-1572
  void readVRORecord(std::string inf);                           // :1505-1508
  void readVRORecord(std::string inf, std::vector<MatchingResult*>& mv);   // :1510-1558
  std::vector<MatchingResult*> mv_vro_res;
  bool addNodeOffline(CCameraNode*, MatchingResult*, bool only_vo = false);   // :1593-1623
  void addEdgeOffline(MatchingResult*);                          // :1652-1668
  void correctMatchingID(MatchingResult* mr);                    // :1626-1649

  gtsam::ISAM2* mp_isam2;
  gtsam::ISAM2Params* mp_isam2_param;
  gtsam::NonlinearFactorGraph* mp_new_fac;
  gtsam::Values* mp_new_node;
  void initISAM2Params();                                        // :93-99

  std::map<int, int> mv_plane_num;
  std::map<int, int> mv_plane_last_seen;
  int m_plane_landmark_id = 0;
  int m_sift_landmark_id = 0;
  bool writeTrajectory(std::string ouf);                         // :1819-1840
  bool trajectoryPLY(std::string ouf, CG::COLOR c);              // :1842-1864
  void headerPLY(std::ofstream&, int vertex_number);             // :1927-1939
  void writeG2O(std::string ouf);                                // :1941-1945 (gtsam::writeG2o: Pose3 vertices, BetweenFactor<Pose3> edges)
};

// ---- IMU
typedef std::vector<std::array<double, 6>> stdv_eigen_vector6d;   // [gx gy gz ax ay az] (imu_vn100.cpp:96)

class CImuBase {
 public:
  CImuBase(double delta_t, gtsam::imuBias::ConstantBias prior_bias);
  virtual ~CImuBase();
  virtual void setStartPoint(double t);                          // imu_base.cpp:108-121
  virtual bool readImuData(std::string f);
  virtual int findIndexAt(double t);                             // :123-154
  virtual bool predictNextFlag(double t, gtsam::NavState&);      // :39-48
  virtual bool predictNextFlag(int next_i, gtsam::NavState&);    // :63-70
  virtual gtsam::NavState predictNext(int next_i);               // :72-87
  virtual gtsam::NavState predictNext(double t);                 // :50-61
  int m_curr_i;
  double getLastTimeStamp();
  void resetGravity(double gx, double gy, double gz);            // :251-256
  static std::shared_ptr<gtsam::PreintegratedCombinedMeasurements::Params> getParam();   // :258-263
  virtual std::shared_ptr<gtsam::PreintegratedCombinedMeasurements::Params> getIMUParams() = 0;
  virtual gtsam::NavState predictBetween(int i, int j, gtsam::NavState& state_i, gtsam::imuBias::ConstantBias bias_i);   // :156-170
  virtual void resetPreintegrationAndBias(gtsam::imuBias::ConstantBias bias);   // :89-93
  virtual void resetPreintegrationAndBias();                     // :95-99
  void setState(gtsam::NavState&);                               // :180-183

  int m_syn_start_id;
  gtsam::imuBias::ConstantBias m_prior_imu_bias;
  gtsam::imuBias::ConstantBias m_prev_imu_bias;
  stdv_eigen_vector6d mv_measurements;
  std::vector<double> mv_timestamps;
  float m_dt;
  gtsam::NavState m_prev_state;
  gtsam::PreintegrationType* mp_combined_pre_imu;
};

class CImuVn100 : public CImuBase {
 public:
  CImuVn100(double dt, gtsam::imuBias::ConstantBias prior_bias);
  virtual ~CImuVn100();
  virtual std::shared_ptr<gtsam::PreintegratedCombinedMeasurements::Params> getIMUParams();   // imu_vn100.cpp:24-67
  virtual bool readImuData(std::string f);                       // imu_vn100.cpp:78-105
  std::vector<std::array<double, 3>> mv_rpy;
};
