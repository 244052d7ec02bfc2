// g2o_driver.cpp -- BASELINE config 1 through the reference's OWN CGraphG2O (g2o/g2o_graph.cpp built unchanged over
// compat/ + the g2o facade, compat/build_ref.py): firstNode (vertex 0 fixed), addToGraph per edge (the first edge that
// reaches a vertex sets its estimate, like CGraphG2O::addNode does), error(), optimizeGraph() (10 x optimize(2)), error(),
// writeTrajectory / writeG2O.  Edge file: "id1 id2 R(9, row-major) t(3) information(36, row-major, [trans, rot])" per line.
//   usage: g2o_driver edges.txt out_trajectory.log out_graph.g2o
#include <cstdio>
#include <fstream>
#include <ros/ros.h>
#include "g2o_graph.h"
#include "camera_node.h"
#include "matching_result.h"

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s edges.txt traj.log graph.g2o\n", argv[0]); return 2; }
  try {
    CGraphG2O g;
    CCameraNode* first = new CCameraNode();
    g.firstNode(first);
    std::ifstream in(argv[1]);
    int id1, id2, n_edges = 0;
    while (in >> id1 >> id2) {
      MatchingResult mr;
      mr.edge.id1 = id1; mr.edge.id2 = id2;
      Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) in >> T(i, j);
      for (int i = 0; i < 3; ++i) in >> T(i, 3);
      mr.edge.transform = T;
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) in >> mr.edge.informationMatrix(i, j);
      mr.succeed_match = true;
      const bool is_new = g.m_graph_map.find(id2) == g.m_graph_map.end();
      if (!g.addToGraph(mr, false)) { fprintf(stderr, "addToGraph failed for %d -> %d\n", id1, id2); return 1; }
      if (is_new) { CCameraNode* n = new CCameraNode(); n->m_id = id2; n->m_seq_id = ++g.m_sequence_id; g.m_graph_map[id2] = n; }
      ++n_edges;
    }
    const double e0 = g.error();
    g.optimizeGraph();
    const double e1 = g.error();
    printf("RESULT nodes %zu edges %d chi2_before %.17g chi2_after %.17g\n", g.camnodeSize(), n_edges, e0, e1);
    if (!g.writeTrajectory(argv[2])) return 1;
    g.writeG2O(argv[3]);
  } catch (const std::exception& e) {
    fprintf(stderr, "g2o_driver failed: %s\n", e.what());
    return 1;
  }
  return 0;
}
