// plane_driver.cpp -- the plane-aided VIO branch (BASELINE config 3) through the reference's OWN CGraphGT::addPlaneFactor
// (gtsam/gtsam_graph.cpp:1118-1298, built unchanged over compat/ + the gtsam facade): camera-frame plane -> IMU frame, the
// covariance J blkdiag(B^T S_n B, S_d) J^T with its conditioning (off-diagonal dropped, float-truncated diagonal), new
// plane landmarks initialised in the world frame, OrientedPlane3Factor per observation -- on top of the VRO + IMU graph the
// offline driver builds (test_vro_imu_graph.cpp:158-357 without images).
// Plane file, one observation per line:  pose_id landmark_id  nx ny nz d  S(16, row-major 4x4 covariance of (n, d)), camera frame.
//   usage: plane_driver vro.log imu.log times.log planes.txt out_poses.txt out_planes.txt
#include <cstdio>
#include <fstream>
#include <map>
#include <ros/ros.h>
#include <gtsam/navigation/CombinedImuFactor.h>
#include <gtsam/geometry/OrientedPlane3.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/inference/Symbol.h>
#include "gtsam_graph.h"
#include "imu_vn100.h"
#include "camera_node.h"
#include "matching_result.h"
#include "plane.h"

using namespace gtsam;
using symbol_shorthand::B;
using symbol_shorthand::L;
using symbol_shorthand::V;
using symbol_shorthand::X;

struct Obs { int pose, lm; CPlane* p; };

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s vro.log imu.log times.log planes.txt out_poses.txt out_planes.txt\n", argv[0]); return 2; }
  try {
    CGraphGT gt_graph;
    gt_graph.readVRORecord(argv[1]);
    gt_graph.setCamera2IMU(0);
    std::map<int, double> img_times;
    { std::ifstream inf(argv[3]); int id; double t; while (inf >> id >> t) img_times[id] = t; }
    std::multimap<int, Obs> obs;
    {
      std::ifstream inf(argv[4]);
      int pose, lm;
      while (inf >> pose >> lm) {
        CPlane* p = new CPlane();
        inf >> p->nx_ >> p->ny_ >> p->nz_ >> p->d1_;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inf >> p->m_CP[i][j];
        obs.insert({pose, Obs{pose, lm, p}});
      }
    }
    imuBias::ConstantBias prior_bias;
    CImuVn100* imu = new CImuVn100(0.005, prior_bias);
    if (!imu->readImuData(argv[2])) return 1;
    CCameraNode* first = new CCameraNode();
    first->m_seq_id = 0;
    gt_graph.firstNode(first, false);
    imu->setStartPoint(img_times[0]);
    int added = 0, rejected = 0;
    auto add_planes = [&](int node_id) {
      auto r = obs.equal_range(node_id);
      for (auto it = r.first; it != r.second; ++it) {
        if (gt_graph.addPlaneFactor(it->second.p, node_id, it->second.lm)) ++added; else ++rejected;
        if (it->second.lm >= gt_graph.m_plane_landmark_id) gt_graph.m_plane_landmark_id = it->second.lm + 1;
      }
    };
    add_planes(0);
    int cur_frame_id = 0;
    for (size_t i = 0; i < gt_graph.mv_vro_res.size(); i++) {
      MatchingResult* pm = gt_graph.mv_vro_res[i];
      if (pm->edge.id2 > cur_frame_id) {
        CCameraNode* node = new CCameraNode();
        bool valid_match = gt_graph.addNodeOffline(node, pm);
        if (!valid_match) gt_graph.m_graph_map[node->m_id] = node;
        NavState cur_p;
        bool imu_available = imu->predictNextFlag(img_times[pm->edge.id2], cur_p);
        PreintegratedCombinedMeasurements* preint = dynamic_cast<PreintegratedCombinedMeasurements*>(imu->mp_combined_pre_imu);
        const int id = node->m_id;
        if (imu_available) {
          CombinedImuFactor imu_factor(X(id - 1), V(id - 1), X(id), V(id), B(id - 1), B(id), *preint);
          gt_graph.mp_fac_graph->add(imu_factor);
          gt_graph.mp_new_fac->add(imu_factor);
          gt_graph.addToGTSAM(cur_p, id, !valid_match);
          NavState st(gt_graph.mp_node_values->at<Pose3>(X(id)), gt_graph.mp_node_values->at<Vector3>(V(id)));
          imu->setState(st);
          imu->resetPreintegrationAndBias(*gt_graph.mp_prev_bias);
        }
        add_planes(id);
        cur_frame_id = pm->edge.id2;
      } else {
        gt_graph.addEdgeOffline(pm);
      }
    }
    const double e0 = gt_graph.error();
    {   // the initial plane landmarks, as addPlaneFactor inserted them
      std::ofstream ouf(std::string(argv[6]) + ".init");
      ouf.precision(17);
      for (int l = 0; l < gt_graph.m_plane_landmark_id; ++l)
        if (gt_graph.mp_node_values->exists(L(l))) { Vector4 c = gt_graph.mp_node_values->at<OrientedPlane3>(L(l)).planeCoefficients(); ouf << l << " " << c(0) << " " << c(1) << " " << c(2) << " " << c(3) << "\n"; }
    }
    gt_graph.optimizeGraphBatch();
    const double e1 = gt_graph.error();
    printf("RESULT nodes %zu planes %d factors_added %d rejected %d error_before %.17g error_after %.17g\n", gt_graph.camnodeSize(), gt_graph.m_plane_landmark_id, added, rejected, e0, e1);
    std::ofstream ouf(argv[5]);
    ouf.precision(17);
    for (auto& kv : gt_graph.m_graph_map) {
      double T[12];
      gt_graph.mp_node_values->at<Pose3>(X(kv.first)).toArray12(T);
      ouf << kv.first;
      for (int k = 0; k < 12; ++k) ouf << " " << T[k];
      ouf << "\n";
    }
    std::ofstream opl(argv[6]);
    opl.precision(17);
    for (int l = 0; l < gt_graph.m_plane_landmark_id; ++l)
      if (gt_graph.mp_node_values->exists(L(l))) { Vector4 c = gt_graph.mp_node_values->at<OrientedPlane3>(L(l)).planeCoefficients(); opl << l << " " << c(0) << " " << c(1) << " " << c(2) << " " << c(3) << "\n"; }
    delete imu;
  } catch (const std::exception& e) {
    fprintf(stderr, "plane_driver failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
