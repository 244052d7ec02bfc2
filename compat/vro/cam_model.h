// vro stand-in (see matching_result.h): CamModel, the pinhole + radial distortion model of the sibling package cam_model.
#pragma once
class CamModel {
 public:
  double fx, fy, cx, cy, k1, k2, k3, p1, p2;
  double z_offset = 0;
  double m_z_scale = 1.0;
  int m_cols = 0, m_rows = 0;
  CamModel(double fx_ = 1, double fy_ = 1, double cx_ = 0, double cy_ = 0, double k1_ = 0, double k2_ = 0, double k3_ = 0, double p1_ = 0, double p2_ = 0)
      : fx(fx_), fy(fy_), cx(cx_), cy(cy_), k1(k1_), k2(k2_), k3(k3_), p1(p1_), p2(p2_) {}
  void setDepthScale(double s) { m_z_scale = s; }
  static CamModel& gCamModel_ref() { static CamModel m; return m; }
  static CamModel* gCamModel() { return &gCamModel_ref(); }
  static void updategCamModel(const CamModel& m) { gCamModel_ref() = m; }
  // pixel (u, v) + depth z -> camera frame point (undistorted pinhole; the front end that needs more is out of scope)
  void convertUVZ2XYZ(float u, float v, double z, double& ox, double& oy, double& oz) const {
    oz = z + z_offset; ox = (u - cx) / fx * oz; oy = (v - cy) / fy * oz;
  }
  void convertXYZ2UV(float x, float y, float z, float& u, float& v) const { u = (float)(fx * x / z + cx); v = (float)(fy * y / z + cy); }
};
