// fg_math.cuh -- SO(3)/SE(3) primitives in fp64 for device code (row-major 3x3 as double[9]).
// Conventions: SURVEY.md A.1 (GTSAM 4.0 semantics, full EXPMAP charts, tangent [rot, trans]).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FG_HD __host__ __device__ __forceinline__
#else
#define FG_HD inline
#endif

namespace fg {

FG_HD void m3_mul(const double* A, const double* B, double* C) {   // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
FG_HD void m3_tmul(const double* A, const double* B, double* C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
FG_HD void m3_mult(const double* A, const double* B, double* C) {  // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
FG_HD void m3_vec(const double* A, const double* v, double* o) {   // o = A v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
FG_HD void m3_tvec(const double* A, const double* v, double* o) {  // o = A^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}
FG_HD void skew3(const double* w, double* S) {
  S[0] = 0; S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2]; S[4] = 0; S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}
FG_HD void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
FG_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// coefficients a = sin t / t, b = (1-cos t)/t^2, c = (t - sin t)/t^3
FG_HD void so3_abc(double th2, double& a, double& b, double& c) {
  if (th2 < 1e-10) {
    a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; c = 1.0 / 6.0 - th2 / 120.0;
  } else {
    double t = sqrt(th2), s, co;
    sincos(t, &s, &co);
    double sh = sin(0.5 * t);
    a = s / t; b = 2.0 * sh * sh / th2; c = (t - s) / (th2 * t);
  }
}

// R = I + a W + b W^2
FG_HD void so3_exp(const double* w, double* R) {
  double th2 = dot3(w, w), a, b, c;
  so3_abc(th2, a, b, c);
  double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  R[0] = 1.0 - b * (yy + zz); R[1] = -a * w[2] + b * xy;  R[2] = a * w[1] + b * xz;
  R[3] = a * w[2] + b * xy;   R[4] = 1.0 - b * (xx + zz); R[5] = -a * w[0] + b * yz;
  R[6] = -a * w[1] + b * xz;  R[7] = a * w[0] + b * yz;   R[8] = 1.0 - b * (xx + yy);
}

// Rot3::Logmap (trace based, GTSAM SO3::Logmap branches)
FG_HD void so3_log(const double* R, double* w) {
  double tr = R[0] + R[4] + R[8];
  if (fabs(tr + 1.0) < 1e-10) {
    const double PI = 3.14159265358979323846;
    if (fabs(R[8] + 1.0) > 1e-10) {
      double k = PI / sqrt(2.0 + 2.0 * R[8]);
      w[0] = k * R[2]; w[1] = k * R[5]; w[2] = k * (1.0 + R[8]);
    } else if (fabs(R[4] + 1.0) > 1e-10) {
      double k = PI / sqrt(2.0 + 2.0 * R[4]);
      w[0] = k * R[1]; w[1] = k * (1.0 + R[4]); w[2] = k * R[7];
    } else {
      double k = PI / sqrt(2.0 + 2.0 * R[0]);
      w[0] = k * (1.0 + R[0]); w[1] = k * R[3]; w[2] = k * R[6];
    }
    return;
  }
  double tr3 = tr - 3.0, mag;
  if (tr3 < -1e-7) {
    double c = 0.5 * (tr - 1.0);
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    double th = acos(c);
    mag = th / (2.0 * sin(th));
  } else {
    mag = 0.5 - tr3 * tr3 / 12.0;
  }
  w[0] = mag * (R[7] - R[5]); w[1] = mag * (R[2] - R[6]); w[2] = mag * (R[3] - R[1]);
}

// right Jacobian Jr = I - b W + c W^2  (ExpmapDerivative)
FG_HD void so3_jr(const double* w, double* J) {
  double th2 = dot3(w, w), a, b, c;
  so3_abc(th2, a, b, c);
  double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  J[0] = 1.0 - c * (yy + zz); J[1] = b * w[2] + c * xy;   J[2] = -b * w[1] + c * xz;
  J[3] = -b * w[2] + c * xy;  J[4] = 1.0 - c * (xx + zz); J[5] = b * w[0] + c * yz;
  J[6] = b * w[1] + c * xz;   J[7] = -b * w[0] + c * yz;  J[8] = 1.0 - c * (xx + yy);
}

// inverse right Jacobian = I + W/2 + k W^2 (LogmapDerivative)
FG_HD void so3_jr_inv(const double* w, double* J) {
  double th2 = dot3(w, w), k;
  if (th2 < 1e-10) {
    k = 1.0 / 12.0 + th2 / 720.0;
  } else {
    double t = sqrt(th2);
    k = 1.0 / th2 - (1.0 + cos(t)) / (2.0 * t * sin(t));
  }
  double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  J[0] = 1.0 - k * (yy + zz);     J[1] = -0.5 * w[2] + k * xy;    J[2] = 0.5 * w[1] + k * xz;
  J[3] = 0.5 * w[2] + k * xy;     J[4] = 1.0 - k * (xx + zz);     J[5] = -0.5 * w[0] + k * yz;
  J[6] = -0.5 * w[1] + k * xz;    J[7] = 0.5 * w[0] + k * yz;     J[8] = 1.0 - k * (xx + yy);
}

// Pose3::Expmap([w, v]) -> R, t = V(w) v with V = I + b W + c W^2
FG_HD void se3_exp(const double* xi, double* R, double* t) {
  const double* w = xi;
  const double* v = xi + 3;
  so3_exp(w, R);
  double th2 = dot3(w, w), a, b, c;
  so3_abc(th2, a, b, c);
  double wv[3], wwv[3];
  cross3(w, v, wv);
  cross3(w, wv, wwv);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = v[i] + b * wv[i] + c * wwv[i];
}

// Pose3::Logmap(R,t) -> [w, u]
FG_HD void se3_log(const double* R, const double* t, double* xi) {
  double w[3];
  so3_log(R, w);
  double th = sqrt(dot3(w, w));
  xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
  if (th < 1e-10) {
    xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2];
  } else {
    double n[3] = {w[0] / th, w[1] / th, w[2] / th};
    double WT[3], WWT[3];
    cross3(n, t, WT);
    cross3(n, WT, WWT);
    double coef = 1.0 - th / (2.0 * tan(0.5 * th));
#pragma unroll
    for (int i = 0; i < 3; ++i) xi[3 + i] = t[i] - 0.5 * th * WT[i] + coef * WWT[i];
  }
}

// pose helpers: pose = R[9] then t[3]
FG_HD void pose_between(const double* Ra, const double* ta, const double* Rb, const double* tb, double* R, double* t) {
  m3_tmul(Ra, Rb, R);
  double d[3] = {tb[0] - ta[0], tb[1] - ta[1], tb[2] - ta[2]};
  m3_tvec(Ra, d, t);
}
FG_HD void pose_compose(const double* Ra, const double* ta, const double* Rb, const double* tb, double* R, double* t) {
  m3_mul(Ra, Rb, R);
  double o[3];
  m3_vec(Ra, tb, o);
  t[0] = o[0] + ta[0]; t[1] = o[1] + ta[1]; t[2] = o[2] + ta[2];
}
// X (+) xi = X * Expmap(xi)
FG_HD void pose_retract(const double* R, const double* t, const double* xi, double* Ro, double* to) {
  double dR[9], dt[3];
  se3_exp(xi, dR, dt);
  pose_compose(R, t, dR, dt, Ro, to);
}
// ---- g2o's SE3 chart (g2o/types/slam3d/isometry3d_mappings: toVectorMQT / fromVectorMQT)
// unit quaternion (w, x, y, z) with w >= 0 of a rotation matrix (Shepperd's method; g2o normalises the sign the same way)
FG_HD void quat_from_rot(const double* M, double* q) {
  const double tr = M[0] + M[4] + M[8];
  if (tr > 0) {
    const double s = sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (M[7] - M[5]) / s; q[2] = (M[2] - M[6]) / s; q[3] = (M[3] - M[1]) / s;
  } else if (M[0] > M[4] && M[0] > M[8]) {
    const double s = sqrt(1.0 + M[0] - M[4] - M[8]) * 2;
    q[0] = (M[7] - M[5]) / s; q[1] = 0.25 * s; q[2] = (M[1] + M[3]) / s; q[3] = (M[2] + M[6]) / s;
  } else if (M[4] > M[8]) {
    const double s = sqrt(1.0 + M[4] - M[0] - M[8]) * 2;
    q[0] = (M[2] - M[6]) / s; q[1] = (M[1] + M[3]) / s; q[2] = 0.25 * s; q[3] = (M[5] + M[7]) / s;
  } else {
    const double s = sqrt(1.0 + M[8] - M[0] - M[4]) * 2;
    q[0] = (M[3] - M[1]) / s; q[1] = (M[2] + M[6]) / s; q[2] = (M[5] + M[7]) / s; q[3] = 0.25 * s;
  }
  if (q[0] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
}
FG_HD void rot_from_quat(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
// VertexSE3::oplus: X <- X * fromVectorMQT(d), d = [t, q_xyz]
FG_HD void g2o_oplus(const double* R, const double* t, const double* d, double* Ro, double* to) {
  const double w2 = 1.0 - (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
  double q[4] = {w2 > 0.0 ? sqrt(w2) : 0.0, d[3], d[4], d[5]};
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
  double dR[9];
  rot_from_quat(q, dR);
  pose_compose(R, t, dR, d, Ro, to);
}

// ---- chart options: GTSAM picks the Pose3 / Rot3 retraction at compile time (GTSAM_POSE3_EXPMAP,
// GTSAM_ROT3_EXPMAP) and the reference's build flags are unknown, so the chart is a context option (fg_set_pose_chart):
//   0 full EXPMAP (default)   2 Pose3 FIRST_ORDER over Rot3 EXPMAP   3 Pose3 FIRST_ORDER over Rot3 CAYLEY (GTSAM 4.0's
//   default build)            1 is g2o's [t, q_xyz] chart (set by the g2o back-end, g2o_oplus above)
// Rot3::CayleyChart::Retract: the Cayley transform of [w/2]x
FG_HD void cayley_retract(const double* w, double* R) {
  const double x = w[0], y = w[1], z = w[2];
  const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
  const double f = 1.0 / (4.0 + x2 + y2 + z2), f2 = 2.0 * f;
  R[0] = (4 + x2 - y2 - z2) * f; R[1] = (xy - 2 * z) * f2;       R[2] = (xz + 2 * y) * f2;
  R[3] = (xy + 2 * z) * f2;       R[4] = (4 - x2 + y2 - z2) * f; R[5] = (yz - 2 * x) * f2;
  R[6] = (xz - 2 * y) * f2;       R[7] = (yz + 2 * x) * f2;       R[8] = (4 - x2 - y2 + z2) * f;
}
// Rot3::CayleyChart::Local: its inverse.  With R = (I + A)(I - A)^-1, A = [a]x, a = w / 2:  a = vee(R - R^T) / (1 + tr R)
FG_HD void cayley_local(const double* R, double* w) {
  const double k = 2.0 / (1.0 + R[0] + R[4] + R[8]);
  w[0] = k * (R[7] - R[5]); w[1] = k * (R[2] - R[6]); w[2] = k * (R[3] - R[1]);
}
// Pose3::ChartAtOrigin::Retract / Local under the chosen chart, tangent [rot, trans]
FG_HD void pose_chart_retract0(const double* xi, int chart, double* R, double* t) {
  if (chart == 2 || chart == 3) {
    if (chart == 2) so3_exp(xi, R); else cayley_retract(xi, R);
    t[0] = xi[3]; t[1] = xi[4]; t[2] = xi[5];
  } else {
    se3_exp(xi, R, t);
  }
}
FG_HD void pose_chart_local0(const double* R, const double* t, int chart, double* xi) {
  if (chart == 2 || chart == 3) {
    if (chart == 2) so3_log(R, xi); else cayley_local(R, xi);
    xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2];
  } else {
    se3_log(R, t, xi);
  }
}
// X (+) xi under the chosen chart (Values::retract of a Pose3)
FG_HD void pose_chart_retract(const double* R, const double* t, const double* xi, int chart, double* Ro, double* to) {
  double dR[9], dt[3];
  pose_chart_retract0(xi, chart, dR, dt);
  pose_compose(R, t, dR, dt, Ro, to);
}

// Ad(T) = [[R,0],[[t]x R, R]]  (6x6 row-major)
FG_HD void adjoint(const double* R, const double* t, double* Ad) {
  double S[9], SR[9];
  skew3(t, S);
  m3_mul(S, R, SR);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Ad[6 * i + j] = R[3 * i + j];
      Ad[6 * i + 3 + j] = 0.0;
      Ad[6 * (i + 3) + j] = SR[3 * i + j];
      Ad[6 * (i + 3) + 3 + j] = R[3 * i + j];
    }
}

// Unit3::basis(): b1 = normalize(n x axis_of_min_abs), b2 = n x b1.  B is 3x2 row-major.
FG_HD void unit3_basis(const double* n, double* B) {
  double mx = fabs(n[0]), my = fabs(n[1]), mz = fabs(n[2]);
  double ax[3] = {0, 0, 0};
  if (mx <= my && mx <= mz) ax[0] = 1.0;
  else if (my <= mx && my <= mz) ax[1] = 1.0;
  else ax[2] = 1.0;
  double b1[3], b2[3];
  cross3(n, ax, b1);
  double inv = 1.0 / sqrt(dot3(b1, b1));
  b1[0] *= inv; b1[1] *= inv; b1[2] *= inv;
  cross3(n, b1, b2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { B[2 * i] = b1[i]; B[2 * i + 1] = b2[i]; }
}

// Unit3::localCoordinates(q) at n -> 2-vector
FG_HD void unit3_local(const double* n, const double* q, double* out) {
  double x = dot3(n, q);
  double z = 1.0 - x * x, y;
  if (z < 2.220446049250313e-16) {
    if (x > 0) y = 1.0 - (x - 1.0) / 3.0;
    else { out[0] = 3.14159265358979323846; out[1] = 0.0; return; }
  } else {
    double xc = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);
    y = acos(xc) / sqrt(z);
  }
  double B[6], d[3];
  unit3_basis(n, B);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = y * (q[i] - x * n[i]);
  out[0] = B[0] * d[0] + B[2] * d[1] + B[4] * d[2];
  out[1] = B[1] * d[0] + B[3] * d[1] + B[5] * d[2];
}

// OrientedPlane3::retract(v3): n <- exp_n(B v01), d += v2
FG_HD void plane_retract(const double* pl, const double* v, double* out) {
  double B[6], xi[3];
  unit3_basis(pl, B);
#pragma unroll
  for (int i = 0; i < 3; ++i) xi[i] = B[2 * i] * v[0] + B[2 * i + 1] * v[1];
  double th = sqrt(dot3(xi, xi));
  double sc = th < 2.220446049250313e-16 ? 1.0 : sin(th) / th;
  double c = cos(th), p[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = c * pl[i] + sc * xi[i];
  double inv = 1.0 / sqrt(dot3(p, p));
  out[0] = p[0] * inv; out[1] = p[1] * inv; out[2] = p[2] * inv;
  out[3] = pl[3] + v[2];
}

}  // namespace fg
