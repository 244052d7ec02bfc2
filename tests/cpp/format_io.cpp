// Host-only exercise of the text formats on the path through the reference's own CGraphGT
// (gtsam/gtsam_graph.cpp built unchanged over compat/ + the gtsam facade, compat/build_ref.py):
// VRO edge log round trip (printVROResult -> readVRORecord), trajectory log, trajectory PLY, g2o export.  No GPU call.
#include <cstdio>
#include <fstream>
#include <ros/ros.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/inference/Symbol.h>
#include "gtsam_graph.h"
#include "camera_node.h"
#include "matching_result.h"
using namespace gtsam;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::string dir = argv[1];
  CGraphGT g;
  g.firstNode(new CCameraNode);
  // three VRO records: 1 <- 0, 2 <- 1 and a "failed match" record (information(0,0) == 10000)
  std::vector<MatchingResult> recs(3);
  for (int k = 0; k < 3; ++k) {
    Vector6 r; for (int i = 0; i < 6; ++i) r(i) = 0.01 * (i + 1) * (k + 1);
    Pose3 p = Pose3::ChartAtOrigin::Retract(r);
    recs[k].final_trafo = p.matrix().cast<float>(); recs[k].edge.transform = p.matrix();
    recs[k].edge.id1 = k; recs[k].edge.id2 = k + 1;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) recs[k].edge.informationMatrix(i, j) = (i == j ? 100.0 + i : 0.5) * (k == 2 && i == 0 && j == 0 ? 0 : 1);
  }
  recs[2].edge.informationMatrix(0, 0) = 10000;
  {
    std::ofstream ouf((dir + "/vro.log").c_str());
    for (auto& m : recs) g.printVROResult(ouf, m);
  }
  g.readVRORecord(dir + "/vro.log");
  printf("RECORDS %zu\n", g.mv_vro_res.size());
  int added = 0;
  for (auto* mr : g.mv_vro_res) {
    CCameraNode* n = new CCameraNode;
    if (g.addNodeOffline(n, mr, false)) ++added; else delete n;      // caller deletes on failure (test_gt_graph.cpp:92-96)
  }
  printf("ADDED %d NODES %zu FACTORS %zu\n", added, g.camnodeSize(), g.mp_fac_graph->size());
  Pose3 p2 = g.mp_node_values->at<Pose3>(symbol_shorthand::X(2));
  printf("X2 %.17g %.17g %.17g\n", p2.x(), p2.y(), p2.z());
  bool ok = g.writeTrajectory(dir + "/traj.log") && g.trajectoryPLY(dir + "/traj.ply", CG::BLUE);
  g.writeG2O(dir + "/graph.g2o");
  printf("WRITE %d\n", ok ? 1 : 0);
  return ok ? 0 : 1;
}
