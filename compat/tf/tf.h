// tf/tf.h stand-in (ROS tf is not installed): tf::Vector3 / Quaternion / Transform as g2o/g2o_graph.cpp:299-330 and
// g2o/misc.h use them (setOrigin / setRotation, composition, getOrigin().x(), getRotation().w()).
#pragma once
#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
namespace tf {
class Vector3 {
  double v_[3] = {0, 0, 0};
 public:
  Vector3() {}
  Vector3(double x, double y, double z) { v_[0] = x; v_[1] = y; v_[2] = z; }
  double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
  double getX() const { return v_[0]; } double getY() const { return v_[1]; } double getZ() const { return v_[2]; }
  void setX(double x) { v_[0] = x; } void setY(double y) { v_[1] = y; } void setZ(double z) { v_[2] = z; }
  double operator[](int i) const { return v_[i]; }
  Vector3 operator+(const Vector3& o) const { return Vector3(v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]); }
};
class Quaternion {
  double x_ = 0, y_ = 0, z_ = 0, w_ = 1;
 public:
  Quaternion() {}
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
  double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } double w() const { return w_; }
  double getX() const { return x_; } double getY() const { return y_; } double getZ() const { return z_; } double getW() const { return w_; }
  void setX(double v) { x_ = v; } void setY(double v) { y_ = v; } void setZ(double v) { z_ = v; } void setW(double v) { w_ = v; }
  Quaternion operator*(const Quaternion& o) const {
    return Quaternion(w_ * o.x_ + x_ * o.w_ + y_ * o.z_ - z_ * o.y_, w_ * o.y_ - x_ * o.z_ + y_ * o.w_ + z_ * o.x_,
                      w_ * o.z_ + x_ * o.y_ - y_ * o.x_ + z_ * o.w_, w_ * o.w_ - x_ * o.x_ - y_ * o.y_ - z_ * o.z_);
  }
  Vector3 rotate(const Vector3& v) const {
    const double tx = 2 * (y_ * v.z() - z_ * v.y()), ty = 2 * (z_ * v.x() - x_ * v.z()), tz = 2 * (x_ * v.y() - y_ * v.x());
    return Vector3(v.x() + w_ * tx + (y_ * tz - z_ * ty), v.y() + w_ * ty + (z_ * tx - x_ * tz), v.z() + w_ * tz + (x_ * ty - y_ * tx));
  }
};
class Transform {
  Quaternion q_; Vector3 t_;
 public:
  Transform() {}
  Transform(const Quaternion& q, const Vector3& t = Vector3()) : q_(q), t_(t) {}
  void setIdentity() { q_ = Quaternion(); t_ = Vector3(); }
  void setOrigin(const Vector3& t) { t_ = t; }
  void setRotation(const Quaternion& q) { q_ = q; }
  const Vector3& getOrigin() const { return t_; }
  Quaternion getRotation() const { return q_; }
  Transform operator*(const Transform& o) const { return Transform(q_ * o.q_, q_.rotate(o.t_) + t_); }
};
}  // namespace tf
