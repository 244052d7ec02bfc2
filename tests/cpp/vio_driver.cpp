// vio_driver.cpp -- test program in the shape of the reference's offline VIO driver
// (gtsam/test_vro_imu_graph.cpp:76-360, without images and planes): VRO edge log + IMU log + image time log ->
// CGraphGT / CImuVn100 -> optimizeGraphBatch.  It is compiled against the REFERENCE's own headers and linked with the
// reference's own gtsam_graph.cpp / imu_base.cpp / imu_vn100.cpp built unchanged over compat/ + the gtsam facade
// (compat/build_ref.py); nothing of CGraphGT is re-implemented in this repository.
//   usage: vio_driver <vro.log> <imu.log> <times.log> <out_poses.txt> [incremental]
// With "incremental" the loop also does what the reference does at the end of every frame
// (test_vro_imu_graph.cpp:344-350): optimizeGraphIncremental(), then the integrator is re-seeded from the estimated
// bias, pose and velocity of the current frame.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <ros/ros.h>
#include <gtsam/navigation/CombinedImuFactor.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/nonlinear/ISAM2.h>
#include <gtsam/inference/Symbol.h>
#include "gtsam_graph.h"
#include "imu_vn100.h"
#include "camera_node.h"
#include "matching_result.h"

using namespace gtsam;
using symbol_shorthand::B;
using symbol_shorthand::V;
using symbol_shorthand::X;

static bool loadImgTime(const char* f, std::map<int, double>& m) {   // `img_id timestamp` (test_vro_imu_graph.cpp:425-442)
  std::ifstream inf(f);
  if (!inf.is_open()) return false;
  int id; double t;
  while (inf >> id >> t) m[id] = t;
  return true;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s vro.log imu.log times.log out.txt\n", argv[0]); return 2; }
  try {
    CGraphGT gt_graph;
    gt_graph.readVRORecord(argv[1]);
    gt_graph.setCamera2IMU(0);
    std::map<int, double> img_times;
    if (!loadImgTime(argv[3], img_times)) { ROS_ERROR("failed to read time file %s", argv[3]); return 1; }
    imuBias::ConstantBias prior_bias;
    double dt = 0.005;   // 200 Hz
    CImuVn100* imu = new CImuVn100(dt, prior_bias);
    if (!imu->readImuData(argv[2])) { ROS_ERROR("failed to load imu data from file %s", argv[2]); return 1; }
    const int g_f_start = 0;
    CCameraNode* pNewNode = new CCameraNode();
    pNewNode->m_seq_id = g_f_start;
    gt_graph.firstNode(pNewNode, false);
    imu->setStartPoint(img_times[g_f_start]);
    int cur_frame_id = g_f_start;
    const bool incremental = argc > 5 && !strcmp(argv[5], "incremental");
    double inc_ms_sum = 0, inc_ms_max = 0, inc_dev_ms = 0; long inc_frames = 0, inc_relin = 0;
    for (size_t i = 0; i < gt_graph.mv_vro_res.size(); i++) {
      MatchingResult* pm = gt_graph.mv_vro_res[i];
      if (pm->edge.id2 <= g_f_start) continue;
      if (pm->edge.id2 > cur_frame_id) {           // a new frame: incremental edge + IMU factor
        int cur_imu_id = pm->edge.id2;
        CCameraNode* node = new CCameraNode();
        bool valid_match = gt_graph.addNodeOffline(node, pm);
        if (!valid_match) gt_graph.m_graph_map[node->m_id] = node;
        NavState cur_p;
        bool imu_available = imu->predictNextFlag(img_times[cur_imu_id], cur_p);
        PreintegratedCombinedMeasurements* preint = dynamic_cast<PreintegratedCombinedMeasurements*>(imu->mp_combined_pre_imu);
        const int cur_node_id = node->m_id;
        if (imu_available) {
          CombinedImuFactor imu_factor(X(cur_node_id - 1), V(cur_node_id - 1), X(cur_node_id), V(cur_node_id),
                                       B(cur_node_id - 1), B(cur_node_id), *preint);
          gt_graph.mp_fac_graph->add(imu_factor);
          gt_graph.mp_new_fac->add(imu_factor);
          gt_graph.addToGTSAM(cur_p, cur_node_id, !valid_match);
          if (!incremental) {
            // re-seed the integrator from the dead-reckoned state
            NavState st(gt_graph.mp_node_values->at<Pose3>(X(cur_node_id)), gt_graph.mp_node_values->at<Vector3>(V(cur_node_id)));
            imu->setState(st);
            imu->resetPreintegrationAndBias(*gt_graph.mp_prev_bias);
          }
        }
        cur_frame_id = pm->edge.id2;
        if (incremental) {
          // the look-back / loop-closure edges of this frame, then one ISAM2 update (test_vro_imu_graph.cpp:323-350)
          size_t j = i + 1;
          while (j < gt_graph.mv_vro_res.size() && gt_graph.mv_vro_res[j]->edge.id2 <= cur_frame_id) gt_graph.addEdgeOffline(gt_graph.mv_vro_res[j++]);
          i = j - 1;
          const auto t0 = std::chrono::steady_clock::now();
          gt_graph.optimizeGraphIncremental();
          const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
          inc_ms_sum += ms; inc_ms_max = ms > inc_ms_max ? ms : inc_ms_max; ++inc_frames;
          inc_relin += gt_graph.mp_isam2->report.n_relinearized; inc_dev_ms += gt_graph.mp_isam2->report.ms_update;
          imu->resetPreintegrationAndBias(gt_graph.mp_node_values->at<imuBias::ConstantBias>(B(cur_node_id)));
          NavState pre_state(gt_graph.mp_node_values->at<Pose3>(X(cur_node_id)), gt_graph.mp_node_values->at<Vector3>(V(cur_node_id)));
          imu->setState(pre_state);
        }
      } else {
        gt_graph.addEdgeOffline(pm);               // look-back / loop-closure edge
      }
    }
    if (incremental) {
      printf("INCREMENTAL frames %ld mean_ms %.4f max_ms %.4f device_ms_mean %.4f relinearized %ld\n", inc_frames, inc_ms_sum / (inc_frames ? inc_frames : 1),
             inc_ms_max, inc_dev_ms / (inc_frames ? inc_frames : 1), inc_relin);
      std::ofstream est(std::string(argv[4]) + ".isam2");
      est.precision(17);
      for (auto& kv : gt_graph.m_graph_map) {
        double T[12];
        gt_graph.mp_node_values->at<Pose3>(X(kv.first)).toArray12(T);
        est << kv.first;
        for (int k = 0; k < 12; ++k) est << " " << T[k];
        est << "\n";
      }
    }
    double e0 = gt_graph.error();
    gt_graph.optimizeGraphBatch();
    double e1 = gt_graph.error();
    printf("RESULT nodes %zu error_before %.17g error_after %.17g\n", gt_graph.camnodeSize(), e0, e1);
    std::ofstream ouf(argv[4]);
    ouf.precision(17);
    for (auto& kv : gt_graph.m_graph_map) {
      double T[12];
      gt_graph.mp_node_values->at<Pose3>(X(kv.first)).toArray12(T);
      ouf << kv.first;
      for (int k = 0; k < 12; ++k) ouf << " " << T[k];
      Vector3 v = gt_graph.mp_node_values->at<Vector3>(V(kv.first));
      ouf << " " << v[0] << " " << v[1] << " " << v[2] << "\n";
    }
    delete imu;
  } catch (const std::exception& e) {
    fprintf(stderr, "vio_driver failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
