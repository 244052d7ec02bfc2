// Multi-frame BA and two-view BA through the reference's OWN wrapper (gtsam/gtsam_graph.cpp built unchanged over compat/ +
// the gtsam facade, compat/build_ref.py):
//   CGraphGT::firstNode / addNodeOffline (VRO edges)      gtsam_graph.cpp:320-368, 1593-1623
//   CGraphGT::addToGTSAM(CCameraNodeBA*, CCameraNodeBA*, map<int,int>&, CamModel*)   :370-448
//   CGraphGT::optimizeGraphBatch                          :1784-1788
//   CGraphGT::bundleAdjust                                :500-610  (LM + Marginals -> edge information)
// Input file: "P n fx fy cx cy k1 k2", then per pose n lines "x y z u v" (camera-frame point, pixel).
// usage: ba_driver features.txt vro.log out.txt
#include <cstdio>
#include <fstream>
#include <ros/ros.h>
#include <gtsam/nonlinear/NonlinearFactorGraph.h>
#include <gtsam/nonlinear/Values.h>
#include <gtsam/inference/Symbol.h>
#include "gtsam_graph.h"
#include "camera_node_ba.h"
#include "matching_result.h"
#include "cam_model.h"
using namespace gtsam;
using symbol_shorthand::X;

struct NodeAll : public CCameraNodeBA {     // "front end": feature k of every frame is the same physical point
  std::map<int, int> matchNodePairBA(CCameraNodeBA* older, Eigen::Matrix4f&, CamModel*) override {
    std::map<int, int> m;
    for (size_t k = 0; k < m_feature_loc_3d.size() && k < older->m_feature_loc_3d.size(); ++k) m[(int)k] = (int)k;
    return m;
  }
};

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  std::ifstream in(argv[1]);
  int P, n; double fx, fy, cx, cy, k1, k2;
  in >> P >> n >> fx >> fy >> cx >> cy >> k1 >> k2;
  CamModel cam(fx, fy, cx, cy, k1, k2);
  std::vector<NodeAll*> nodes(P);
  for (int p = 0; p < P; ++p) {
    nodes[p] = new NodeAll;
    nodes[p]->m_feature_loc_3d.resize(n); nodes[p]->m_feature_loc_2d.resize(n); nodes[p]->mv_feature_qid.assign(n, -1);
    for (int k = 0; k < n; ++k) {
      double x, y, z, u, v; in >> x >> y >> z >> u >> v;
      nodes[p]->m_feature_loc_3d[k] = Eigen::Vector4f((float)x, (float)y, (float)z, 1.f);
      nodes[p]->m_feature_loc_2d[k].pt.x = (float)u; nodes[p]->m_feature_loc_2d[k].pt.y = (float)v;
    }
  }
  CGraphGT g;
  g.readVRORecord(argv[2]);
  // two-view BA of the first edge before the graph takes the nodes (bundleAdjust looks node id1 up in m_graph_map)
  g.firstNode(nodes[0], false);
  nodes[0]->m_seq_id = 0;
  MatchingResult ba = *g.mv_vro_res[0];
  bool ok_ba = g.bundleAdjust(&ba, nodes[1], &cam);
  // the graph: VRO edges for the poses, BA factors between consecutive frames
  for (int p = 1; p < P; ++p) g.addNodeOffline(nodes[p], g.mv_vro_res[p - 1], true);
  std::map<int, int> all;
  for (int k = 0; k < n; ++k) all[k] = k;
  for (int p = 0; p + 1 < P; ++p) g.addToGTSAM(nodes[p], nodes[p + 1], all, &cam);
  double e0 = g.error();
  g.optimizeGraphBatch();
  double e1 = g.error();
  printf("RESULT nodes %zu factors %zu landmarks %d e0 %.17g e1 %.17g ba %d\n", g.camnodeSize(), g.mp_fac_graph->size(), g.m_sift_landmark_id, e0, e1, ok_ba ? 1 : 0);
  FILE* f = fopen(argv[3], "w");
  for (int p = 0; p < P; ++p) {
    Pose3 T = g.mp_node_values->at<Pose3>(X(p));
    double a[12]; T.toArray12(a);
    for (int i = 0; i < 12; ++i) fprintf(f, "%.17g ", a[i]);
    fprintf(f, "\n");
  }
  { double a[12]; Pose3(Eigen::Matrix4d(ba.final_trafo.cast<double>())).toArray12(a); for (int i = 0; i < 12; ++i) fprintf(f, "%.17g ", a[i]); fprintf(f, "\n"); }
  for (int r = 0; r < 6; ++r) { for (int c = 0; c < 6; ++c) fprintf(f, "%.17g ", ba.edge.informationMatrix(r, c)); fprintf(f, "0 0 0 0 0 0\n"); }
  fclose(f);
  return 0;
}
