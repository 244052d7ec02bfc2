// plane_check.cpp -- the front-end numerics next to the path through the reference's OWN code, built unchanged
// over compat/ + the gtsam facade: CGraphGT::computeSdj (gtsam/gtsam_graph.cpp:725-748: plane propagated through a relative
// pose, variance of the point-to-plane distance from the plane and pose covariances) and CGraphGT::inThisPlane (:750-764:
// the Mahalanobis-style inlier test predictPlaneNode applies per pixel).  Host only: no optimisation, no device.
//   input  (stdin), one case per line:  ni(4)  Sni(9, row-major)  Sdi  T(12: R row-major, t)  St(9)  point(3)
//   output (stdout), one line per case: S_dj  nj(4)  inlier(0/1)
#include <cstdio>
#include <iostream>
#include <gtsam/geometry/OrientedPlane3.h>
#include <gtsam/geometry/Pose3.h>
#include "gtsam_graph.h"
#include "plane.h"

using namespace gtsam;

int main() {
  CGraphGT g;
  double v[38];
  while (true) {
    for (int i = 0; i < 38; ++i) if (!(std::cin >> v[i])) return 0;
    Vector4 ni, nj;
    ni << v[0], v[1], v[2], v[3];
    Matrix3 Sni, St, R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Sni(i, j) = v[4 + 3 * i + j]; R(i, j) = v[14 + 3 * i + j]; St(i, j) = v[26 + 3 * i + j]; }
    const double Sdi = v[13];
    Pose3 T(Rot3(R), Point3(v[23], v[24], v[25]));
    const double S_dj = g.computeSdj(ni, Sni, Sdi, &T, St, nj);
    CPlane pj;
    pj.nx_ = nj(0); pj.ny_ = nj(1); pj.nz_ = nj(2); pj.d1_ = nj(3);
    const bool in = g.inThisPlane(&pj, S_dj, v[35], v[36], v[37]);
    printf("%.17g %.17g %.17g %.17g %.17g %d\n", S_dj, nj(0), nj(1), nj(2), nj(3), in ? 1 : 0);
  }
}
