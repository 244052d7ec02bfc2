// gtsam_graph.cpp -- CGraphGT / CImuBase / CImuVn100 over the C ABI (see gtsam_graph.h for the reference lines).
// Written from the behaviour described in SURVEY.md sections 3, 8a and Appendix B/D; the numerics live in
// libfg_b200.so.
#include "gtsam_graph.h"
#include <cmath>
#include <sstream>

using namespace gtsam;
using symbol_shorthand::B;
using symbol_shorthand::L;
using symbol_shorthand::Q;
using symbol_shorthand::V;
using symbol_shorthand::X;

// ------------------------------------------------------------------ CGraphGT
CGraphGT::CGraphGT() {
  mp_fac_graph = new NonlinearFactorGraph;
  mp_new_fac = new NonlinearFactorGraph;
  mp_node_values = new Values;
  mp_new_node = new Values;
  initISAM2Params();
  mp_w2o = new Pose3;
  mp_u2c = new Pose3;
  mp_prev_bias = new imuBias::ConstantBias;
  mp_prev_state = new NavState;
}

void CGraphGT::initISAM2Params() {
  mp_isam2_param = new ISAM2Params;
  mp_isam2_param->relinearizeThreshold = 0.1;
  mp_isam2_param->relinearizeSkip = 1;
  mp_isam2 = new ISAM2(*mp_isam2_param);
}

CGraphGT::~CGraphGT() {
  delete mp_prev_bias; delete mp_prev_state; delete mp_fac_graph; delete mp_new_fac; delete mp_new_node;
  delete mp_node_values; delete mp_w2o; delete mp_u2c; delete mp_isam2; delete mp_isam2_param;
  for (auto& kv : m_graph_map) delete kv.second;            // the graph owns its camera nodes
  for (auto* p : mv_vro_res) delete p;
}

double CGraphGT::error() { return mp_fac_graph->error(*mp_node_values); }

void CGraphGT::setWorld2Original(double p) {
  Rot3 R_g2b = Rot3::RzRyRx(-M_PI / 2., 0, -M_PI / 2.);
  Rot3 R_b2o = Rot3::RzRyRx(p, 0, 0);
  (*mp_w2o) = Pose3::Create(R_g2b * R_b2o, Point3());
}
void CGraphGT::setCamera2IMUTranslation(double px, double py, double pz) { (*mp_u2c) = Pose3::Create(Rot3(), vec3(px, py, pz)); }
void CGraphGT::setCamera2IMU(double p) {
  Rot3 R_g2b = Rot3::RzRyRx(M_PI / 2., 0., M_PI / 2.);
  Rot3 R_b2o = Rot3::RzRyRx(p, 0, 0);
  (*mp_u2c) = Pose3::Create(R_g2b * R_b2o, Point3());
}

void CGraphGT::firstNode(CCameraNode* n, bool online) {
  n->m_id = (int)m_graph_map.size();
  m_sequence_id = 0;
  if (online) n->m_seq_id = ++m_sequence_id;
  Pose3 origin_priorMean;                                   // identity
  mp_node_values->insert(X(n->m_id), origin_priorMean);
  mp_new_node->insert(X(n->m_id), origin_priorMean);
  Vector6 s; for (int i = 0; i < 6; ++i) s[i] = 1e-7;
  noiseModel::Diagonal::shared_ptr priorNoise = noiseModel::Diagonal::Sigmas(s);
  mp_fac_graph->add(PriorFactor<Pose3>(X(n->m_id), origin_priorMean, priorNoise));
  mp_new_fac->add(PriorFactor<Pose3>(X(n->m_id), origin_priorMean, priorNoise));
  m_graph_map[n->m_id] = n;
  Vector3 priorVelocity;
  mp_node_values->insert(V(n->m_id), priorVelocity);
  mp_new_node->insert(V(n->m_id), priorVelocity);
  imuBias::ConstantBias priorBias;
  mp_node_values->insert(B(n->m_id), priorBias);
  mp_new_node->insert(B(n->m_id), priorBias);
  noiseModel::Diagonal::shared_ptr velocity_noise_model = noiseModel::Isotropic::Sigma(3, 1e-3);
  noiseModel::Diagonal::shared_ptr bias_noise_model = noiseModel::Isotropic::Sigma(6, 1e-3);
  mp_fac_graph->add(PriorFactor<Vector3>(V(n->m_id), priorVelocity, velocity_noise_model));
  mp_fac_graph->add(PriorFactor<imuBias::ConstantBias>(B(n->m_id), priorBias, bias_noise_model));
  mp_new_fac->add(PriorFactor<Vector3>(V(n->m_id), priorVelocity, velocity_noise_model));
  mp_new_fac->add(PriorFactor<imuBias::ConstantBias>(B(n->m_id), priorBias, bias_noise_model));
}

bool CGraphGT::addToGTSAM(NavState& new_state, int vid, bool add_pose) {
  if (add_pose) {
    mp_node_values->insert(X(vid), new_state.pose());
    mp_new_node->insert(X(vid), new_state.pose());
  }
  mp_node_values->insert(V(vid), new_state.v());
  mp_node_values->insert(B(vid), *mp_prev_bias);
  mp_new_node->insert(V(vid), new_state.v());
  mp_new_node->insert(B(vid), *mp_prev_bias);
  return true;
}

bool CGraphGT::addToGTSAM(MatchingResult& mr, bool set_estimate) {
  bool pre_exist = mp_node_values->exists(X(mr.edge.id1));
  bool cur_exist = mp_node_values->exists(X(mr.edge.id2));
  Pose3 inc_pose = mr.edge.transform;
  inc_pose = (*mp_u2c) * inc_pose * (*mp_u2c).inverse();     // camera frame -> IMU frame
  if (!pre_exist && !cur_exist) {
    ROS_ERROR("%s two nodes %i and %i both not exist ", __FILE__, mr.edge.id1, mr.edge.id2);
    return false;
  } else if (!pre_exist) {
    ROS_WARN("this case is weired, has not solved it!");
    Pose3 cur_pose = mp_node_values->at<Pose3>(X(mr.edge.id2));
    Pose3 pre_pose = cur_pose * inc_pose.inverse();
    mp_node_values->insert(X(mr.edge.id1), pre_pose);
    mp_new_node->insert(X(mr.edge.id1), pre_pose);
  } else if (!cur_exist) {
    Pose3 pre_pose = mp_node_values->at<Pose3>(X(mr.edge.id1));
    Pose3 cur_pose = pre_pose * inc_pose;
    mp_node_values->insert(X(mr.edge.id2), cur_pose);
    mp_new_node->insert(X(mr.edge.id2), cur_pose);
  } else if (set_estimate) {
    Pose3 pre_pose = mp_node_values->at<Pose3>(X(mr.edge.id1));
    Pose3 cur_pose = pre_pose * inc_pose;
    mp_node_values->update(X(mr.edge.id2), cur_pose);
    if (mp_new_node->exists(X(mr.edge.id2))) mp_new_node->update(X(mr.edge.id2), cur_pose);
    else mp_new_node->insert(X(mr.edge.id2), cur_pose);
  }
  // information is frame-changed with the covariance rule Ad * Omega * Ad^T (preserved quirk)
  Matrix6 Adj_Tuc = (*mp_u2c).AdjointMap();
  Matrix6 tmp = Adj_Tuc * mr.edge.informationMatrix * Adj_Tuc.transpose();
  noiseModel::Gaussian::shared_ptr visual_odometry_noise = noiseModel::Gaussian::Information(tmp);
  mp_fac_graph->add(BetweenFactor<Pose3>(X(mr.edge.id1), X(mr.edge.id2), inc_pose, visual_odometry_noise));
  mp_new_fac->add(BetweenFactor<Pose3>(X(mr.edge.id1), X(mr.edge.id2), inc_pose, visual_odometry_noise));
  return true;
}

// Multi-frame BA builder (gtsam_graph.cpp:370-448): a match between two features without a landmark creates one
// (camera point -> IMU frame -> world through X_i) with PriorFactor<Point3>(sigma 0.014) and two projection factors
// (sigma 1 px, Cal3DS2 from the camera model, body_P_sensor = u2c); a match to an existing landmark adds one factor.
bool CGraphGT::addToGTSAM(CCameraNodeBA* ni, CCameraNodeBA* nj, std::map<int, int>& matches, CamModel* pcam) {
  std::shared_ptr<Cal3DS2> K(new Cal3DS2(pcam->fx, pcam->fy, 0, pcam->cx, pcam->cy, pcam->k1, pcam->k2));
  Pose3 Pi = mp_node_values->at<Pose3>(X(ni->m_id));
  noiseModel::Isotropic::shared_ptr pointNoise = noiseModel::Isotropic::Sigma(3, 0.014);
  noiseModel::Isotropic::shared_ptr measNoise = noiseModel::Isotropic::Sigma(2, 1.);
  auto meas = [](CCameraNodeBA* n, int k) { Point2 m; m[0] = n->m_feature_loc_2d[k].pt.x; m[1] = n->m_feature_loc_2d[k].pt.y; return m; };
  auto proj = [&](CCameraNodeBA* n, int k, int qid) {
    GenericProjectionFactor f(meas(n, k), measNoise, X(n->m_id), Q(qid), K, 0, 0, *mp_u2c);
    mp_fac_graph->push_back(f); mp_new_fac->push_back(f);
  };
  for (auto it = matches.begin(); it != matches.end(); ++it) {
    int& qi = ni->mv_feature_qid[it->first];
    int& qj = nj->mv_feature_qid[it->second];
    if (qi == -1 && qj == -1) {
      const std::array<float, 4>& pt = ni->m_feature_loc_3d[it->first];
      Point3 q = Pi.transform_from(mp_u2c->transform_from(vec3(pt[0], pt[1], pt[2])));
      mp_node_values->insertPoint(Q(m_sift_landmark_id), q);
      mp_new_node->insertPoint(Q(m_sift_landmark_id), q);
      mp_fac_graph->push_back(PriorFactor<Point3>(Q(m_sift_landmark_id), q, pointNoise, true));
      mp_new_fac->push_back(PriorFactor<Point3>(Q(m_sift_landmark_id), q, pointNoise, true));
      proj(ni, it->first, m_sift_landmark_id);
      proj(nj, it->second, m_sift_landmark_id);
      qi = qj = m_sift_landmark_id++;
    } else if (qi == -1) {
      proj(ni, it->first, qj); qi = qj;
    } else if (qj == -1) {
      proj(nj, it->second, qi); qj = qi;
    } else if (qi != qj) {
      ROS_ERROR("%s what? Line %d different landmark at the same position", __FILE__, __LINE__);
    }
  }
  return true;
}

// Two-view BA (gtsam_graph.cpp:500-610): refines the transform of a VRO edge and replaces its information matrix by the
// inverse of pose 1's marginal covariance.  Private two-pose graph: prior on s0 (1e-7), PriorFactor<Point3> (0.014) and two
// projection factors per match (1 px, SR4000 Cal3DS2, no body_P_sensor), LM, Marginals.
bool CGraphGT::bundleAdjust(MatchingResult* pm, CCameraNode* pNewNode, CamModel* pcam) {
  int pre_id1 = pm->edge.id1, pre_id2 = pm->edge.id2;
  correctMatchingID(pm);
  CCameraNodeBA* ni = dynamic_cast<CCameraNodeBA*>(m_graph_map[pm->edge.id1]);
  CCameraNodeBA* nj = dynamic_cast<CCameraNodeBA*>(pNewNode);
  Matrix4 Tji = Pose3(pm->final_trafo).inverse().matrix();
  std::map<int, int> matches = (ni && nj) ? nj->matchNodePairBA(ni, Tji, pcam) : std::map<int, int>();
  if (matches.size() <= 4) {
    if (pm->edge.informationMatrix(0, 0) == 10000) ROS_ERROR("Nothing changed for edge from %d to %d", pre_id1, pre_id2);
    pm->edge.id1 = pre_id1; pm->edge.id2 = pre_id2;
    return false;
  }
  NonlinearFactorGraph g;
  Vector6 s; for (int i = 0; i < 6; ++i) s[i] = 1e-7;
  noiseModel::Diagonal::shared_ptr priorNoise = noiseModel::Diagonal::Sigmas(s);
  noiseModel::Isotropic::shared_ptr measNoise = noiseModel::Isotropic::Sigma(2, 1.);
  noiseModel::Isotropic::shared_ptr pointNoise = noiseModel::Isotropic::Sigma(3, 0.014);
  std::shared_ptr<Cal3DS2> K(new Cal3DS2(250.5773, 250.5773, 0, 90, 70, -0.8466, 0.5370));
  Pose3 Pi, Pj;
  g.push_back(PriorFactor<Pose3>(Symbol('s', 0), Pi, priorNoise));
  Values initialEstimate;
  initialEstimate.insert<Pose3>(Symbol('s', 0), Pi);
  initialEstimate.insert<Pose3>(Symbol('s', 1), Pj);
  int j = 0;
  for (auto it = matches.begin(); it != matches.end(); ++it, ++j) {
    const std::array<float, 4>& pt = ni->m_feature_loc_3d[it->first];
    Point3 q = vec3(pt[0], pt[1], pt[2]);
    initialEstimate.insertPoint(Symbol('u', j), q);
    Point2 mi, mj;
    mi[0] = ni->m_feature_loc_2d[it->first].pt.x; mi[1] = ni->m_feature_loc_2d[it->first].pt.y;
    mj[0] = nj->m_feature_loc_2d[it->second].pt.x; mj[1] = nj->m_feature_loc_2d[it->second].pt.y;
    g.push_back(PriorFactor<Point3>(Symbol('u', j), q, pointNoise, true));
    g.push_back(GenericProjectionFactor(mi, measNoise, Symbol('s', 0), Symbol('u', j), K, 0, 0, Pose3()));
    g.push_back(GenericProjectionFactor(mj, measNoise, Symbol('s', 1), Symbol('u', j), K, 0, 0, Pose3()));
  }
  LevenbergMarquardtOptimizer optimizer(g, initialEstimate);
  Values curEstimate = optimizer.optimize();
  Pj = curEstimate.at<Pose3>(Symbol('s', 1));
  pm->final_trafo = Pj.matrix();
  Marginals marginals(g, curEstimate, Marginals::CHOLESKY);
  int dim = 0;
  std::vector<double> S_pose = marginals.marginalCovariance(Symbol('s', 1), &dim);
  // information = inverse of the 6 x 6 marginal covariance (symmetric positive definite): Gauss-Jordan on the host
  double A[6][12];
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { A[r][c] = S_pose[r * 6 + c]; A[r][6 + c] = r == c ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int piv = c; for (int r = c + 1; r < 6; ++r) if (std::fabs(A[r][c]) > std::fabs(A[piv][c])) piv = r;
    if (piv != c) for (int k = 0; k < 12; ++k) std::swap(A[c][k], A[piv][k]);
    double d = A[c][c]; for (int k = 0; k < 12; ++k) A[c][k] /= d;
    for (int r = 0; r < 6; ++r) if (r != c) { double f = A[r][c]; if (f != 0.0) for (int k = 0; k < 12; ++k) A[r][k] -= f * A[c][k]; }
  }
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) pm->edge.informationMatrix(r, c) = A[r][6 + c];
  pm->edge.id1 = pre_id1; pm->edge.id2 = pre_id2;
  return true;
}

void CGraphGT::fakeOdoNode(CCameraNode* new_node) {
  if (new_node->m_id != (int)m_graph_map.size()) {
    std::cerr << __FILE__ << " " << __LINE__ << " Here this should not happen!" << std::endl;
    new_node->m_id = (int)m_graph_map.size();
    new_node->m_seq_id = ++m_sequence_id;
  }
  CCameraNode* pre_node = m_graph_map[new_node->m_id - 1];
  MatchingResult mr;
  mr.edge.id1 = pre_node->m_id;
  mr.edge.id2 = new_node->m_id;
  mr.edge.informationMatrix = Matrix6::Identity() * 1e4;
  addToGTSAM(mr, false);
  m_graph_map[new_node->m_id] = new_node;
}

bool CGraphGT::addPlaneFactor(const Vector4& plane_imu, const Matrix3& S_in, int node_id, int landmark) {
  if (landmark < 0) { ROS_ERROR("%s landmark should not be -1,", __FILE__); return false; }
  bool landmark_exist = mp_node_values->exists(L(landmark));
  if (!mp_node_values->exists(X(node_id))) { ROS_ERROR("%s pose node %d not exist!", __FILE__, node_id); return false; }
  OrientedPlane3 ONJ(plane_imu);
  Matrix3 S_upj = S_in;
  if (!landmark_exist) {
    Pose3 Twu = mp_node_values->at<Pose3>(X(node_id));
    OrientedPlane3 ONW = ONJ.transform(Twu.inverse());
    mp_node_values->insert(L(landmark), ONW);
    mp_new_node->insert(L(landmark), ONW);
    mv_plane_num[landmark] = 0;
  }
  // covariance conditioning of the reference: drop the off-diagonal coupling, quantise the diagonal
  S_upj(0, 1) = S_upj(1, 0) = 0;
  for (int i = 0; i < 3; i++) S_upj(i, i) = (float)((int)(S_upj(i, i) * 1e8)) * 1e-8 + 1e-8;
  OrientedPlane3Factor plane_factor(ONJ.planeCoefficients(), noiseModel::Gaussian::Covariance(S_upj), X(node_id), L(landmark));
  mp_fac_graph->add(plane_factor);
  mp_new_fac->add(plane_factor);
  mv_plane_last_seen[landmark] = node_id;
  return true;
}

void CGraphGT::optimizeGraphIncremental() {
  mp_isam2->update(*mp_new_fac, *mp_new_node);
  (*mp_node_values) = mp_isam2->calculateEstimate();
  mp_new_fac->resize(0);
  mp_new_node->clear();
}
void CGraphGT::optimizeGraph() { return CGraphGT::optimizeGraphBatch(); }
void CGraphGT::optimizeGraphBatch() {
  LevenbergMarquardtOptimizer optimizer(*mp_fac_graph, *mp_node_values);
  (*mp_node_values) = optimizer.optimize();
}

void CGraphGT::readVRORecord(std::string fname) { return readVRORecord(fname, mv_vro_res); }
void CGraphGT::readVRORecord(std::string fname, std::vector<MatchingResult*>& mv) {
  std::ifstream inf(fname.c_str());
  if (!inf.is_open()) { std::cerr << " failed to open file " << fname << std::endl; return; }
  while (!inf.eof()) {
    int id_to, id_from;
    Vector6 r;
    MatchingResult* pm = new MatchingResult;
    inf >> id_to >> id_from;
    for (int i = 0; i < 6; i++) inf >> r(i);
    Pose3 p = Pose3::ChartAtOrigin::Retract(r);
    pm->final_trafo = p.matrix();
    pm->edge.transform = p;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) { inf >> pm->edge.informationMatrix(i, j); pm->edge.informationMatrix(j, i) = pm->edge.informationMatrix(i, j); }
    pm->edge.id2 = id_to;
    pm->edge.id1 = id_from;
    if (inf.eof() || inf.fail()) { delete pm; break; }       // trailing whitespace after the last record
    mv.push_back(pm);
  }
  std::cout << __LINE__ << " read vro records " << mv.size() << std::endl;
}

void CGraphGT::printVROResult(std::ostream& ouf, MatchingResult& m) {
  Vector6 p = Pose3::ChartAtOrigin::Local(Pose3(m.final_trafo));
  ouf.precision(17);
  ouf << m.edge.id2 << " " << m.edge.id1 << " " << p(0) << " " << p(1) << " " << p(2) << " " << p(3) << " " << p(4) << " " << p(5) << " ";
  for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++) ouf << m.edge.informationMatrix(i, j) << " ";
  ouf << std::endl;
}

bool CGraphGT::addNodeOffline(CCameraNode* new_node, MatchingResult* mr, bool only_vo) {
  bool ret = true;
  new_node->m_id = (int)m_graph_map.size();
  new_node->m_seq_id = mr->edge.id2;
  if (only_vo || mr->edge.informationMatrix(0, 0) != 10000) {   // 10000 in slot (0,0) marks a failed match
    m_graph_map[new_node->m_id] = new_node;
    int pre_id1 = mr->edge.id1, pre_id2 = mr->edge.id2;
    correctMatchingID(mr);
    addToGTSAM(*mr, true);
    mr->edge.id1 = pre_id1;
    mr->edge.id2 = pre_id2;
  } else {
    ret = false;
  }
  return ret;
}

void CGraphGT::correctMatchingID(MatchingResult* mr) {
  int from_id = mr->edge.id1, to_id = mr->edge.id2;
  bool from_good = false, to_good = false;
  for (auto it = m_graph_map.begin(); it != m_graph_map.end(); ++it) {
    if (it->second->m_seq_id == from_id) { mr->edge.id1 = it->second->m_id; from_good = true; }
    if (it->second->m_seq_id == to_id) { mr->edge.id2 = it->second->m_id; to_good = true; }
    if (from_good && to_good) break;
  }
}

void CGraphGT::addEdgeOffline(MatchingResult* mr) {
  if (mr->edge.informationMatrix(0, 0) != 10000) {
    int pre_id1 = mr->edge.id1, pre_id2 = mr->edge.id2;
    correctMatchingID(mr);
    addToGTSAM(*mr, false);
    mr->edge.id1 = pre_id1;
    mr->edge.id2 = pre_id2;
  }
}

bool CGraphGT::writeTrajectory(std::string f) {
  std::ofstream ouf(f.c_str());
  if (!ouf.is_open()) { printf("%s failed to open f: %s to write trajectory!\n", __FILE__, f.c_str()); return false; }
  ouf.precision(17);
  for (auto it = m_graph_map.begin(); it != m_graph_map.end(); ++it) {
    Pose3 p = (*mp_w2o) * mp_node_values->at<Pose3>(X(it->first));
    const Matrix3& R = p.rotation().matrix();
    double qw = 0.5 * std::sqrt(std::max(0.0, 1 + R(0, 0) + R(1, 1) + R(2, 2)));
    double qx = 0.5 * std::sqrt(std::max(0.0, 1 + R(0, 0) - R(1, 1) - R(2, 2))); if (R(2, 1) - R(1, 2) < 0) qx = -qx;
    double qy = 0.5 * std::sqrt(std::max(0.0, 1 - R(0, 0) + R(1, 1) - R(2, 2))); if (R(0, 2) - R(2, 0) < 0) qy = -qy;
    double qz = 0.5 * std::sqrt(std::max(0.0, 1 - R(0, 0) - R(1, 1) + R(2, 2))); if (R(1, 0) - R(0, 1) < 0) qz = -qz;
    ouf << it->first << " " << p.x() << " " << p.y() << " " << p.z() << " " << qx << " " << qy << " " << qz << " " << qw << " " << it->second->m_seq_id << std::endl;
  }
  return true;
}

namespace CG {
unsigned char g_color[][3] = {{255, 0, 0}, {0, 255, 0}, {0, 0, 255}, {255, 0, 255}, {255, 255, 255}, {255, 255, 0}, {0, 0, 0}};
}

static void quat_of(const Matrix3& R, double q[4]) {         // (x, y, z, w), w >= 0
  q[3] = 0.5 * std::sqrt(std::max(0.0, 1 + R(0, 0) + R(1, 1) + R(2, 2)));
  q[0] = 0.5 * std::sqrt(std::max(0.0, 1 + R(0, 0) - R(1, 1) - R(2, 2))); if (R(2, 1) - R(1, 2) < 0) q[0] = -q[0];
  q[1] = 0.5 * std::sqrt(std::max(0.0, 1 - R(0, 0) + R(1, 1) - R(2, 2))); if (R(0, 2) - R(2, 0) < 0) q[1] = -q[1];
  q[2] = 0.5 * std::sqrt(std::max(0.0, 1 - R(0, 0) - R(1, 1) + R(2, 2))); if (R(1, 0) - R(0, 1) < 0) q[2] = -q[2];
}

void CGraphGT::headerPLY(std::ofstream& ouf, int vertex_number) {
  ouf << "ply" << std::endl << "format ascii 1.0" << std::endl << "element vertex " << vertex_number << std::endl
      << "property float x" << std::endl << "property float y" << std::endl << "property float z" << std::endl
      << "property uchar red" << std::endl << "property uchar green" << std::endl << "property uchar blue" << std::endl
      << "end_header" << std::endl;
}

bool CGraphGT::trajectoryPLY(std::string f, CG::COLOR c) {
  std::ofstream ouf(f.c_str());
  if (!ouf.is_open()) { printf("%s %d failed to open f: %s to write trajectory!\n", __FILE__, __LINE__, f.c_str()); return false; }
  headerPLY(ouf, (int)m_graph_map.size());
  for (auto it = m_graph_map.begin(); it != m_graph_map.end(); ++it) {
    Pose3 p = (*mp_w2o) * mp_node_values->at<Pose3>(X(it->first));
    ouf << p.x() << " " << p.y() << " " << p.z() << " " << (int)CG::g_color[c][0] << " " << (int)CG::g_color[c][1] << " " << (int)CG::g_color[c][2] << std::endl;
  }
  return true;
}

// gtsam::writeG2o semantics [ext]: one VERTEX_SE3:QUAT per Pose3 value (key index, t, quaternion xyzw), one
// EDGE_SE3:QUAT per BetweenFactor<Pose3> with the information matrix re-ordered from GTSAM's [rot, trans] tangent to
// g2o's [trans, rot] and written as its 21 upper-triangular entries.
void CGraphGT::writeG2O(std::string f) {
  std::ofstream ouf(f.c_str());
  if (!ouf.is_open()) { printf("%s failed to open f: %s to write g2o!\n", __FILE__, f.c_str()); return; }
  ouf.precision(17);
  const unsigned long long mask = (1ull << 56) - 1;
  for (auto& kv : mp_node_values->m) {
    if (kv.second.type != FG_T_POSE) continue;
    Pose3 p = mp_node_values->at<Pose3>(kv.first);
    double q[4]; quat_of(p.rotation().matrix(), q);
    ouf << "VERTEX_SE3:QUAT " << (kv.first & mask) << " " << p.x() << " " << p.y() << " " << p.z() << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
  }
  for (auto& fac : mp_fac_graph->f) {
    auto* b = dynamic_cast<BetweenFactor<Pose3>*>(fac.get());
    if (!b) continue;
    double q[4]; quat_of(b->z.rotation().matrix(), q);
    ouf << "EDGE_SE3:QUAT " << (b->k1 & mask) << " " << (b->k2 & mask) << " " << b->z.x() << " " << b->z.y() << " " << b->z.z() << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3];
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) ouf << " " << b->nm->info[((i + 3) % 6) * 6 + (j + 3) % 6];     // swap the rot / trans blocks
    ouf << std::endl;
  }
}

// ------------------------------------------------------------------ CImuBase
CImuBase::CImuBase(double delta_t, imuBias::ConstantBias prior_bias)
    : m_curr_i(0), m_syn_start_id(0), m_prior_imu_bias(prior_bias), m_dt((float)delta_t), mp_combined_pre_imu(0) {
  m_prev_state = NavState();
  m_prev_imu_bias = m_prior_imu_bias;
}
CImuBase::~CImuBase() { if (mp_combined_pre_imu) { delete mp_combined_pre_imu; mp_combined_pre_imu = 0; } }

double CImuBase::getLastTimeStamp() {
  if (mv_timestamps.size() > 0) return mv_timestamps[mv_timestamps.size() - 1];
  std::cerr << __FILE__ << " at " << __LINE__ << " no imu timestamps available!" << std::endl;
  return 0;
}
bool CImuBase::predictNextFlag(double t, NavState& s) {
  int index = findIndexAt(t);
  if (index < 0) { std::cerr << __FILE__ << " failed to predictNext given t = " << t << std::endl; return false; }
  return predictNextFlag(index, s);
}
NavState CImuBase::predictNext(double t) {
  NavState ret;
  int index = findIndexAt(t);
  if (index < 0) { std::cerr << __FILE__ << " failed to predictNext given t = " << t << std::endl; return ret; }
  return predictNext(index);
}
bool CImuBase::predictNextFlag(int next_t, NavState& s) {
  if (next_t < 0) return false;
  s = predictNext(next_t);
  return true;
}
NavState CImuBase::predictNext(int next_i) {
  PreintegratedCombinedMeasurements* pim = dynamic_cast<PreintegratedCombinedMeasurements*>(mp_combined_pre_imu);
  for (int i = m_syn_start_id + m_curr_i; i < m_syn_start_id + next_i; i++) {
    if (i >= (int)mv_measurements.size()) { printf("%s i >= mv_measurements.size()\n", __FILE__); break; }
    const std::array<double, 6>& imu = mv_measurements[i];
    pim->integrateMeasurement(vec3(imu[3], imu[4], imu[5]), vec3(imu[0], imu[1], imu[2]), m_dt);   // acc = tail<3>, gyro = head<3>
  }
  m_curr_i = next_i;
  return pim->predict(m_prev_state, m_prev_imu_bias);
}
void CImuBase::resetPreintegrationAndBias(imuBias::ConstantBias bias) {
  m_prev_imu_bias = bias;
  dynamic_cast<PreintegratedCombinedMeasurements*>(mp_combined_pre_imu)->resetIntegrationAndSetBias(bias);
}
void CImuBase::resetPreintegrationAndBias() {
  dynamic_cast<PreintegratedCombinedMeasurements*>(mp_combined_pre_imu)->resetIntegrationAndSetBias(m_prev_imu_bias);
}
bool CImuBase::readImuData(std::string) { printf("%s readImuData not implemented\n", __FILE__); return false; }
void CImuBase::setStartPoint(double t) {
  m_syn_start_id = 0;
  int index = findIndexAt(t);
  if (index < 0) { std::cerr << __FILE__ << " failed to synchronize with timestamp t = " << t << std::endl; return; }
  m_syn_start_id = index;
}
int CImuBase::findIndexAt(double t) {
  if (mv_timestamps.size() != mv_measurements.size()) {
    std::cerr << __FILE__ << " something is wrong: mv_timestamps.size() != mv_measurements.size()" << std::endl;
    return -1;
  }
  int e = (int)mv_timestamps.size() - 1;
  int i;
  for (i = 0; i + m_syn_start_id <= e; i++) {
    if (mv_timestamps[i + m_syn_start_id] > t) {
      if (i >= 1) {
        // nearest of the two neighbours; ties go to the later sample (reference tie-break, imu_base.cpp:141-144)
        if (mv_timestamps[i + m_syn_start_id] - t > t - mv_timestamps[i + m_syn_start_id - 1]) return i - 1;
        else return i;
      }
      std::cout << __FILE__ << " timestamp[0] > t " << std::endl;
      return i;
    }
  }
  std::cout << __FILE__ << " timestamp[-1] < t = " << std::fixed << t << " i=" << i << std::endl;
  return -1;
}
NavState CImuBase::predictBetween(int i, int j, NavState& state_i, imuBias::ConstantBias bias_i) {
  resetPreintegrationAndBias(bias_i);
  PreintegratedCombinedMeasurements* pim = dynamic_cast<PreintegratedCombinedMeasurements*>(mp_combined_pre_imu);
  for (int m = i; m < j; m++) {
    if (m >= (int)mv_measurements.size()) { printf("%s m >= mv_measurements.size()\n", __FILE__); break; }
    const std::array<double, 6>& imu = mv_measurements[m];
    pim->integrateMeasurement(vec3(imu[3], imu[4], imu[5]), vec3(imu[0], imu[1], imu[2]), m_dt);
  }
  return pim->predict(state_i, m_prev_imu_bias);
}
void CImuBase::setState(NavState& ns) { m_prev_state = ns; }
void CImuBase::resetGravity(double gx, double gy, double gz) { getParam()->n_gravity = vec3(gx, gy, gz); }
std::shared_ptr<PreintegratedCombinedMeasurements::Params> CImuBase::getParam() {
  // one function-static Params shared by every CImuBase, gravity MakeSharedD(9.71)
  static std::shared_ptr<PreintegratedCombinedMeasurements::Params> p = PreintegratedCombinedMeasurements::Params::MakeSharedD(9.71);
  return p;
}

// ------------------------------------------------------------------ CImuVn100
CImuVn100::CImuVn100(double dt, imuBias::ConstantBias prior_bias) : CImuBase(dt, prior_bias) {
  std::shared_ptr<PreintegratedCombinedMeasurements::Params> p = getIMUParams();
  mp_combined_pre_imu = new PreintegratedCombinedMeasurements(p, m_prior_imu_bias);
}
CImuVn100::~CImuVn100() {}

std::shared_ptr<PreintegratedCombinedMeasurements::Params> CImuVn100::getIMUParams() {
  std::shared_ptr<PreintegratedCombinedMeasurements::Params> p = CImuBase::getParam();
  static bool b_once = true;
  if (b_once) {
    float fps = 200;
    int hour = 3600;
    double g = 9.81;
    double gyro_noise_density = 0.0035, accel_noise_density = 0.14, gyro_bias_stability = 10, accel_bias_stability = 0.04;
    double accel_noise_sigma = accel_noise_density * 1e-3 * g;
    double gyro_noise_sigma = D2R(gyro_noise_density);
    double accel_bias_rw_sigma = (accel_bias_stability * 1e-3 * g) * sqrt(fps);
    double gyro_bias_rw_sigma = (D2R(gyro_bias_stability) / hour) * sqrt(fps);
    p->accelerometerCovariance = Matrix33::Identity() * pow(accel_noise_sigma, 2);
    p->gyroscopeCovariance = Matrix33::Identity() * pow(gyro_noise_sigma, 2);
    p->integrationCovariance = Matrix33::Identity() * 1e-4;
    p->biasAccCovariance = Matrix33::Identity() * pow(accel_bias_rw_sigma, 2);
    p->biasOmegaCovariance = Matrix33::Identity() * pow(gyro_bias_rw_sigma, 2);
    p->biasAccOmegaInt = Matrix66::Identity() * 1e-3;
    b_once = false;
  }
  return p;
}

bool CImuVn100::readImuData(std::string fname) {
  std::ifstream inf(fname.c_str());
  if (!inf.is_open()) { printf("%s failed to open imu file %s\n", __FILE__, fname.c_str()); return false; }
  double t;
  float ax, ay, az, gx, gy, gz, yaw, pitch, roll;          // the reference reads the samples as float (imu_vn100.cpp:86)
  // `t ax ay az gx gy gz yaw pitch roll` per line, stored [gx gy gz ax ay az].  The reference's while(!eof) loop
  // re-appends the last record; that duplicate is reproduced so indices line up.
  bool any = false;
  while (true) {
    bool ok = static_cast<bool>(inf >> t >> ax >> ay >> az >> gx >> gy >> gz >> yaw >> pitch >> roll);
    if (!ok && !any) break;
    any = true;
    mv_measurements.push_back({gx, gy, gz, ax, ay, az});
    mv_rpy.push_back({roll, pitch, yaw});
    mv_timestamps.push_back(t);
    if (!ok) break;
  }
  return mv_measurements.size() > 0;
}
