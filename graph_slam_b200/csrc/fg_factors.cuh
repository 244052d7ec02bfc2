// fg_factors.cuh -- per-factor residual + Jacobian evaluation in fp64 (device code).
// Each function restates the GTSAM 4.0 semantics the reference relies on;
// the reference call site that instantiates the factor is cited per function.
#pragma once
#include "fg_math.cuh"

namespace fg {

// BetweenFactor<Pose3>  (gtsam/gtsam_graph.cpp:691-692; A.3)
//   r = Local(Z, h) = ChartAtOrigin::Local(Z^-1 X1^-1 X2) under the context's chart (fg_math.cuh), H1 = -Ad(h^-1), H2 = I
//   (GTSAM's fast path: the derivative of Local is not chained in).  J1 is 6x6 row-major (only if JAC).
template <bool JAC>
FG_HD void between_eval(const double* X1, const double* X2, const double* Z, double* r, double* J1, int chart = 0) {
  double Rh[9], th[3], Re[9], te[3];
  pose_between(X1, X1 + 9, X2, X2 + 9, Rh, th);
  pose_between(Z, Z + 9, Rh, th, Re, te);
  pose_chart_local0(Re, te, chart, r);
  if (JAC) {
    double Rhi[9], thi[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rhi[3 * i + j] = Rh[3 * j + i];
    m3_tvec(Rh, th, thi);
    thi[0] = -thi[0]; thi[1] = -thi[1]; thi[2] = -thi[2];
    adjoint(Rhi, thi, J1);
#pragma unroll
    for (int i = 0; i < 36; ++i) J1[i] = -J1[i];
  }
}

// g2o::EdgeSE3  (g2o/g2o_graph.cpp:125-132): e = toVectorMQT(Z^-1 X1^-1 X2) = [t, q_xyz], tangent order
// [trans, rot]; Jacobians w.r.t. VertexSE3::oplus (X <- X * fromVectorMQT(d)) of both ends, 6x6 row-major, at d = 0:
//   E = A B, A = Z^-1, B = X1^-1 X2;  q_e = (w, v) the quaternion of R_E, Q = w I + [v]x
//   J2 = [[R_E, 0], [0, Q]]        J1 = [[-R_A, 2 R_A [t_B]x], [0, -Q R_B^T]]
// (the exact derivatives of the error map, which is what g2o's analytic EdgeSE3 Jacobians are).
template <bool JAC>
FG_HD void g2o_edge_eval(const double* X1, const double* X2, const double* Z, double* e, double* J1, double* J2) {
  double Rb[9], tb[3], Re[9], te[3], q[4];
  pose_between(X1, X1 + 9, X2, X2 + 9, Rb, tb);
  pose_between(Z, Z + 9, Rb, tb, Re, te);
  quat_from_rot(Re, q);
  e[0] = te[0]; e[1] = te[1]; e[2] = te[2]; e[3] = q[1]; e[4] = q[2]; e[5] = q[3];
  if (JAC) {
    double Q[9], S[9], RaS[9], QRbt[9];
    skew3(q + 1, Q);
    Q[0] += q[0]; Q[4] += q[0]; Q[8] += q[0];
    skew3(tb, S);
    m3_tmul(Z, S, RaS);                 // R_A = R_Z^T
    m3_mult(Q, Rb, QRbt);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        J2[6 * i + j] = Re[3 * i + j];  J2[6 * i + 3 + j] = 0.0;
        J2[6 * (i + 3) + j] = 0.0;      J2[6 * (i + 3) + 3 + j] = Q[3 * i + j];
        J1[6 * i + j] = -Z[3 * j + i];  J1[6 * i + 3 + j] = 2.0 * RaS[3 * i + j];
        J1[6 * (i + 3) + j] = 0.0;      J1[6 * (i + 3) + 3 + j] = -QRbt[3 * i + j];
      }
  }
}

// PriorFactor<Pose3>  (gtsam/gtsam_graph.cpp:341; A.3): r = Local(prior, x) = ChartAtOrigin::Local(prior^-1 x), H = I.
FG_HD void prior_pose_eval(const double* X, const double* Pm, double* r, int chart = 0) {
  double R[9], t[3];
  pose_between(Pm, Pm + 9, X, X + 9, R, t);
  pose_chart_local0(R, t, chart, r);
}

// Camera model: Cal3DS2 K = (fx,fy,s,u0,v0,k1,k2,p1,p2); sensor = body_P_sensor pose (12).
// GenericProjectionFactor<Pose3,Point3,Cal3DS2>  (gtsam/gtsam_graph.cpp:405-406; A.2)
//   r (2), Jp (2x6 row-major, w.r.t. body pose), Jl (2x3).  Cheirality (z<=0): r = 2fx, J = 0.
template <bool JAC>
FG_HD void projection_eval(const double* X, const double* p, const double* uv, const double* K,
                           const double* S, double* r, double* Jp, double* Jl) {
  double Rc[9], tc[3];
  pose_compose(X, X + 9, S, S + 9, Rc, tc);
  double d3[3] = {p[0] - tc[0], p[1] - tc[1], p[2] - tc[2]}, q[3];
  m3_tvec(Rc, d3, q);
  if (q[2] <= 0.0) {
    r[0] = 2.0 * K[0]; r[1] = 2.0 * K[0];
    if (JAC) {
#pragma unroll
      for (int i = 0; i < 12; ++i) Jp[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) Jl[i] = 0.0;
    }
    return;
  }
  const double fx = K[0], fy = K[1], s = K[2], u0 = K[3], v0 = K[4], k1 = K[5], k2 = K[6], p1 = K[7], p2 = K[8];
  double d = 1.0 / q[2];
  double x = q[0] * d, y = q[1] * d;
  double r2 = x * x + y * y;
  double g = 1.0 + k1 * r2 + k2 * r2 * r2;
  double xd = g * x + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
  double yd = g * y + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
  r[0] = fx * xd + s * yd + u0 - uv[0];
  r[1] = fy * yd + v0 - uv[1];
  if (JAC) {
    double gk = 2.0 * (k1 + 2.0 * k2 * r2);
    double gx = x * gk, gy = y * gk;
    double D00 = g + x * gx + 2.0 * p1 * y + 6.0 * p2 * x;
    double D01 = x * gy + 2.0 * p1 * x + 2.0 * p2 * y;
    double D10 = y * gx + 2.0 * p1 * x + 2.0 * p2 * y;
    double D11 = g + y * gy + 6.0 * p1 * y + 2.0 * p2 * x;
    // Dpi = [[fx, s],[0, fy]] * D
    double P00 = fx * D00 + s * D10, P01 = fx * D01 + s * D11;
    double P10 = fy * D10, P11 = fy * D11;
    // d(x,y)/d camera-pose (right tangent)
    double A[12] = {x * y, -(1.0 + x * x), y, -d, 0.0, d * x,
                    1.0 + y * y, -x * y, -x, 0.0, -d, d * y};
    double C[12];   // Dpi * A  (2x6) w.r.t. camera pose
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      C[j] = P00 * A[j] + P01 * A[6 + j];
      C[6 + j] = P10 * A[j] + P11 * A[6 + j];
    }
    // body pose: Jp = C * Ad(S^-1);  S^-1 = (Rs^T, -Rs^T ts);  Ad = [[Ri,0],[[ti]x Ri, Ri]]
    const double* Rs = S;
    const double* ts = S + 9;
    double Ri[9], ti[3], SR[9], Sk[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Ri[3 * i + j] = Rs[3 * j + i];
    m3_tvec(Rs, ts, ti);
    ti[0] = -ti[0]; ti[1] = -ti[1]; ti[2] = -ti[2];
    skew3(ti, Sk);
    m3_mul(Sk, Ri, SR);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        // rotation columns: C[:,0:3] Ri + C[:,3:6] SR ; translation columns: C[:,3:6] Ri
        Jp[6 * a + j] = C[6 * a] * Ri[j] + C[6 * a + 1] * Ri[3 + j] + C[6 * a + 2] * Ri[6 + j]
                      + C[6 * a + 3] * SR[j] + C[6 * a + 4] * SR[3 + j] + C[6 * a + 5] * SR[6 + j];
        Jp[6 * a + 3 + j] = C[6 * a + 3] * Ri[j] + C[6 * a + 4] * Ri[3 + j] + C[6 * a + 5] * Ri[6 + j];
      }
    // point: Dpi * d*[[1,0,-x],[0,1,-y]] * Rc^T
    double E[6] = {d * P00, d * P01, -d * (P00 * x + P01 * y),
                   d * P10, d * P11, -d * (P10 * x + P11 * y)};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Jl[3 * a + j] = E[3 * a] * Rc[3 * j] + E[3 * a + 1] * Rc[3 * j + 1] + E[3 * a + 2] * Rc[3 * j + 2];
  }
}

// OrientedPlane3Factor  (gtsam/gtsam_graph.cpp:1265; A.4)
//   r = transform(plane, pose).error(z);  Hr 3x6, Hp 3x3 are the Jacobians of transform.
template <bool JAC>
FG_HD void plane_eval(const double* X, const double* pl, const double* z, double* r, double* Hr, double* Hp) {
  const double* R = X;
  const double* t = X + 9;
  double q[3];
  m3_tvec(R, pl, q);
  double dp = dot3(pl, t) + pl[3];
  double e2[2];
  unit3_local(q, z, e2);
  r[0] = -e2[0]; r[1] = -e2[1]; r[2] = dp - z[3];
  if (JAC) {
    double Bq[6], Bn[6], Sq[9];
    unit3_basis(q, Bq);
    unit3_basis(pl, Bn);
    skew3(q, Sq);
#pragma unroll
    for (int i = 0; i < 18; ++i) Hr[i] = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Hr[6 * a + j] = Bq[a] * Sq[j] + Bq[2 + a] * Sq[3 + j] + Bq[4 + a] * Sq[6 + j];
    Hr[12 + 3] = q[0]; Hr[12 + 4] = q[1]; Hr[12 + 5] = q[2];
    // Hp[0:2,0:2] = Bq^T R^T Bn
    double RtBn[6];  // 3x2
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int b = 0; b < 2; ++b) RtBn[2 * i + b] = R[i] * Bn[b] + R[3 + i] * Bn[2 + b] + R[6 + i] * Bn[4 + b];
#pragma unroll
    for (int i = 0; i < 9; ++i) Hp[i] = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) Hp[3 * a + b] = Bq[a] * RtBn[b] + Bq[2 + a] * RtBn[2 + b] + Bq[4 + a] * RtBn[4 + b];
    Hp[6] = Bn[0] * t[0] + Bn[2] * t[1] + Bn[4] * t[2];
    Hp[7] = Bn[1] * t[0] + Bn[3] * t[1] + Bn[5] * t[2];
    Hp[8] = 1.0;
  }
}

// Device-side record of a CombinedImuFactor (pim + information = preintMeasCov^-1).
struct ImuRec {
  double dt;
  double preint[9];
  double Hba[27];
  double Hbg[27];
  double bias_hat[6];
  double gravity[3];
  double info[225];
};

// small dense helpers on row-major blocks with leading dimensions
FG_HD void mm33(const double* A, int lda, const double* B, int ldb, double* C, int ldc, double alpha) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * ldc + j] = alpha * (A[i * lda] * B[j] + A[i * lda + 1] * B[ldb + j] + A[i * lda + 2] * B[2 * ldb + j]);
}

// CombinedImuFactor::evaluateError  (gtsam/test_vro_imu_graph.cpp:191-196; A.5/A.6)
//   r[15]; J (15x30 row-major, ld 30) columns: pose_i 0-5, vel_i 6-8, pose_j 9-14, vel_j 15-17, bias_i 18-23, bias_j 24-29.
template <bool JAC>
FG_HD void imu_eval(const double* Xi, const double* vi, const double* Xj, const double* vj,
                    const double* bi, const double* bj, const ImuRec* f, double* r, double* J) {
  const double* Ri = Xi; const double* ti = Xi + 9;
  const double* Rj = Xj; const double* tj = Xj + 9;
  double inc[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) inc[i] = bi[i] - f->bias_hat[i];
  double bc[9];
  for (int i = 0; i < 9; ++i)
    bc[i] = f->preint[i] + f->Hba[3 * i] * inc[0] + f->Hba[3 * i + 1] * inc[1] + f->Hba[3 * i + 2] * inc[2]
          + f->Hbg[3 * i] * inc[3] + f->Hbg[3 * i + 1] * inc[4] + f->Hbg[3 * i + 2] * inc[5];
  double dt = f->dt, dt22 = 0.5 * dt * dt;
  double Rtv[3], Rtg[3];
  m3_tvec(Ri, vi, Rtv);
  m3_tvec(Ri, f->gravity, Rtg);
  double xth[3] = {bc[0], bc[1], bc[2]}, xp[3], xv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    xp[i] = bc[3 + i] + dt * Rtv[i] + dt22 * Rtg[i];
    xv[i] = bc[6 + i] + dt * Rtg[i];
  }
  double bRc[9], Rp[9], tp[3], vp[3], tmp[3];
  so3_exp(xth, bRc);
  m3_mul(Ri, bRc, Rp);
  m3_vec(Ri, xp, tmp);
  tp[0] = ti[0] + tmp[0]; tp[1] = ti[1] + tmp[1]; tp[2] = ti[2] + tmp[2];
  m3_vec(Ri, xv, tmp);
  vp[0] = vi[0] + tmp[0]; vp[1] = vi[1] + tmp[1]; vp[2] = vi[2] + tmp[2];
  double dR[9], dtr[3], dvr[3], eth[3];
  m3_tmul(Rj, Rp, dR);
  double dd[3] = {tp[0] - tj[0], tp[1] - tj[1], tp[2] - tj[2]};
  m3_tvec(Rj, dd, dtr);
  double dv[3] = {vp[0] - vj[0], vp[1] - vj[1], vp[2] - vj[2]};
  m3_tvec(Rj, dv, dvr);
  so3_log(dR, eth);
#pragma unroll
  for (int i = 0; i < 3; ++i) { r[i] = eth[i]; r[3 + i] = dtr[i]; r[6 + i] = dvr[i]; }
#pragma unroll
  for (int i = 0; i < 6; ++i) r[9 + i] = bi[i] - bj[i];
  if (!JAC) return;

  for (int i = 0; i < 450; ++i) J[i] = 0.0;
  // H1p = D_predict_state + D_predict_delta * D_delta_state  (9x9, NavState tangent [theta,p,v])
  double bRcT[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) bRcT[3 * i + j] = bRc[3 * j + i];
  double Sxp[9], Sxv[9], Stv[9], Stg[9];
  skew3(xp, Sxp); skew3(xv, Sxv); skew3(Rtv, Stv); skew3(Rtg, Stg);
  double H1p[81];
  for (int i = 0; i < 81; ++i) H1p[i] = 0.0;
  // row block theta: [bRcT, 0, 0]
  // row block p: [-bRcT Sxp + bRcT (dt Stv + dt22 Stg), bRcT, bRcT*dt]
  // row block v: [-bRcT Sxv + bRcT (dt Stg),            0,    bRcT]
  double Mp[9], Mv[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    Mp[i] = -Sxp[i] + dt * Stv[i] + dt22 * Stg[i];
    Mv[i] = -Sxv[i] + dt * Stg[i];
  }
  double BMp[9], BMv[9];
  m3_mul(bRcT, Mp, BMp);
  m3_mul(bRcT, Mv, BMv);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      H1p[9 * i + j] = bRcT[3 * i + j];
      H1p[9 * (3 + i) + j] = BMp[3 * i + j];
      H1p[9 * (3 + i) + 3 + j] = bRcT[3 * i + j];
      H1p[9 * (3 + i) + 6 + j] = dt * bRcT[3 * i + j];
      H1p[9 * (6 + i) + j] = BMv[3 * i + j];
      H1p[9 * (6 + i) + 6 + j] = bRcT[3 * i + j];
    }
  // localCoordinates Jacobians
  double Jri[9], dRT[9];
  so3_jr_inv(eth, Jri);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) dRT[3 * i + j] = dR[3 * j + i];
  // Dep = blkdiag(Jri, dR, dR);  DH = Dep * H1p (9x9)
  double DH[81];
  for (int c = 0; c < 9; ++c) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      DH[9 * i + c] = Jri[3 * i] * H1p[c] + Jri[3 * i + 1] * H1p[9 + c] + Jri[3 * i + 2] * H1p[18 + c];
      DH[9 * (3 + i) + c] = dR[3 * i] * H1p[27 + c] + dR[3 * i + 1] * H1p[36 + c] + dR[3 * i + 2] * H1p[45 + c];
      DH[9 * (6 + i) + c] = dR[3 * i] * H1p[54 + c] + dR[3 * i + 1] * H1p[63 + c] + dR[3 * i + 2] * H1p[72 + c];
    }
  }
  // pose_i: DH[:,0:6]; vel_i: DH[:,6:9] * Ri^T
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) J[30 * i + j] = DH[9 * i + j];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      J[30 * i + 6 + j] = DH[9 * i + 6] * Ri[3 * j] + DH[9 * i + 7] * Ri[3 * j + 1] + DH[9 * i + 8] * Ri[3 * j + 2];
  }
  // pose_j: Dej[:,0:6] = [[-Jri dR^T, 0],[skew(dtr), -I],[skew(dvr), 0]]; vel_j: Dej[:,6:9] Rj^T = [0;0;-Rj^T]
  double JdT[9], Sdt[9], Sdv[9];
  m3_mul(Jri, dRT, JdT);
  skew3(dtr, Sdt); skew3(dvr, Sdv);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[30 * i + 9 + j] = -JdT[3 * i + j];
      J[30 * (3 + i) + 9 + j] = Sdt[3 * i + j];
      J[30 * (3 + i) + 12 + j] = (i == j) ? -1.0 : 0.0;
      J[30 * (6 + i) + 9 + j] = Sdv[3 * i + j];
      J[30 * (6 + i) + 15 + j] = -Rj[3 * j + i];
    }
  // bias_i: Dep * Dpd * [Hba, Hbg] with Dpd = blkdiag(Jr(xth), bRcT, bRcT)
  double Jr[9], T0[9], T1[9];
  so3_jr(xth, Jr);
  m3_mul(Jri, Jr, T0);      // theta rows
  m3_mul(dR, bRcT, T1);     // p and v rows
  for (int c = 0; c < 6; ++c) {
    const double* Hsrc = (c < 3) ? f->Hba : f->Hbg;
    int cc = c % 3;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      J[30 * i + 18 + c] = T0[3 * i] * Hsrc[cc] + T0[3 * i + 1] * Hsrc[3 + cc] + T0[3 * i + 2] * Hsrc[6 + cc];
      J[30 * (3 + i) + 18 + c] = T1[3 * i] * Hsrc[9 + cc] + T1[3 * i + 1] * Hsrc[12 + cc] + T1[3 * i + 2] * Hsrc[15 + cc];
      J[30 * (6 + i) + 18 + c] = T1[3 * i] * Hsrc[18 + cc] + T1[3 * i + 1] * Hsrc[21 + cc] + T1[3 * i + 2] * Hsrc[24 + cc];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    J[30 * (9 + i) + 18 + i] = 1.0;
    J[30 * (9 + i) + 24 + i] = -1.0;
  }
}

// ---------------------------------------------------------------- preintegration (A.5)
struct ImuParamsDev {
  double acc_cov[9], gyro_cov[9], int_cov[9], bias_acc_cov[9], bias_gyro_cov[9], bint[36], gravity[3];
};

// d/dtheta [Jr(theta) c] at fixed c (exact derivative; see oracle/imu.py::_d_jr_c)
FG_HD void d_jr_c(const double* th, const double* c, double* D) {
  double th2 = dot3(th, th), b, c3, db, dc;
  if (th2 < 1e-10) {
    b = 0.5 - th2 / 24.0; c3 = 1.0 / 6.0 - th2 / 120.0; db = -1.0 / 12.0; dc = -1.0 / 60.0;
  } else {
    double t = sqrt(th2), s, co;
    sincos(t, &s, &co);
    b = (1.0 - co) / th2; c3 = (t - s) / (th2 * t);
    db = (t * s - 2.0 * (1.0 - co)) / (th2 * t) / t;
    dc = ((1.0 - co) / (th2 * t) - 3.0 * (t - s) / (th2 * th2)) / t;
  }
  double txc[3], ttc[3], Sc[9];
  cross3(th, c, txc);
  cross3(th, txc, ttc);
  skew3(c, Sc);
  double tc = dot3(th, c);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      D[3 * i + j] = b * Sc[3 * i + j] - txc[i] * db * th[j]
                   + c3 * ((i == j ? tc : 0.0) + th[i] * c[j] - 2.0 * c[i] * th[j]) + ttc[i] * dc * th[j];
}

FG_HD void inv3(const double* A, double* Ai) {
  double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  double det = A[0] * c0 + A[1] * c1 + A[2] * c2, id = 1.0 / det;
  Ai[0] = c0 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  Ai[3] = c1 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  Ai[6] = c2 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

}  // namespace fg
