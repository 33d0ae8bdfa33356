#include "transforms.h"
#include <cmath>

namespace {
struct Basis { double m[3][3]; };

// what tf2::Matrix3x3::setRotation does with a (not necessarily unit) quaternion
Basis basis_of(const geometry_msgs::Quaternion& q) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  const double s = 2.0 / n2;
  const double xs = q.x * s, ys = q.y * s, zs = q.z * s;
  const double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
  const double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
  const double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
  Basis b;
  b.m[0][0] = 1.0 - (yy + zz); b.m[0][1] = xy - wz;         b.m[0][2] = xz + wy;
  b.m[1][0] = xy + wz;         b.m[1][1] = 1.0 - (xx + zz); b.m[1][2] = yz - wx;
  b.m[2][0] = xz - wy;         b.m[2][1] = yz + wx;         b.m[2][2] = 1.0 - (xx + yy);
  return b;
}

// what tf2::Matrix3x3::getRotation does (trace branch, else largest diagonal element)
geometry_msgs::Quaternion quaternion_of(const Basis& b) {
  double t[4];
  const double trace = b.m[0][0] + b.m[1][1] + b.m[2][2];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    t[3] = s * 0.5;
    s = 0.5 / s;
    t[0] = (b.m[2][1] - b.m[1][2]) * s;
    t[1] = (b.m[0][2] - b.m[2][0]) * s;
    t[2] = (b.m[1][0] - b.m[0][1]) * s;
  } else {
    const int i = b.m[0][0] < b.m[1][1] ? (b.m[1][1] < b.m[2][2] ? 2 : 1) : (b.m[0][0] < b.m[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(b.m[i][i] - b.m[j][j] - b.m[k][k] + 1.0);
    t[i] = s * 0.5;
    s = 0.5 / s;
    t[3] = (b.m[k][j] - b.m[j][k]) * s;
    t[j] = (b.m[j][i] + b.m[i][j]) * s;
    t[k] = (b.m[k][i] + b.m[i][k]) * s;
  }
  geometry_msgs::Quaternion q;
  q.x = t[0]; q.y = t[1]; q.z = t[2]; q.w = t[3];
  return q;
}

// tf2::doTransform on a Vector3: rotation only
void rotate(const geometry_msgs::Quaternion& q, double x, double y, double z, double out[3]) {
  const Basis b = basis_of(q);
  for (int r = 0; r < 3; ++r) out[r] = b.m[r][0] * x + b.m[r][1] * y + b.m[r][2] * z;
}
}  // namespace

void Transforms::CarToWorldRotation(const geometry_msgs::Pose& pose, double R[4]) {
  const Basis b = basis_of(quaternion_of(basis_of(pose.orientation)));
  R[0] = b.m[0][0]; R[1] = b.m[0][1]; R[2] = b.m[1][0]; R[3] = b.m[1][1];
}

std::pair<float, float> Transforms::CarPointToWorldPoint(float x, float y, geometry_msgs::Pose& current_pose) {
  const geometry_msgs::Quaternion q = quaternion_of(basis_of(current_pose.orientation));  // fromMsg -> toMsg
  double w[3];
  rotate(q, x, y, 0, w);                                                                   // doTransform
  const float car_x = current_pose.position.x, car_y = current_pose.position.y;            // narrowed (transforms.cpp:17-18)
  return std::pair<float, float>(w[0] + car_x, w[1] + car_y);
}

geometry_msgs::TransformStamped Transforms::WorldToCarTransform(const geometry_msgs::Pose& pose) {
  const Basis b = basis_of(pose.orientation);
  Basis inv;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) inv.m[r][c] = b.m[c][r];
  const double o[3] = {-pose.position.x, -pose.position.y, -pose.position.z};
  geometry_msgs::TransformStamped out;
  out.transform.translation.x = inv.m[0][0] * o[0] + inv.m[0][1] * o[1] + inv.m[0][2] * o[2];
  out.transform.translation.y = inv.m[1][0] * o[0] + inv.m[1][1] * o[1] + inv.m[1][2] * o[2];
  out.transform.translation.z = inv.m[2][0] * o[0] + inv.m[2][1] * o[1] + inv.m[2][2] * o[2];
  out.transform.rotation = quaternion_of(inv);
  return out;
}

std::pair<float, float> Transforms::TransformPoint(std::pair<float, float> point, geometry_msgs::TransformStamped& tm) {
  double w[3];
  rotate(tm.transform.rotation, point.first, point.second, 0, w);
  w[0] += tm.transform.translation.x;  // transforms.cpp:41-42
  w[1] += tm.transform.translation.y;
  return std::pair<float, float>(w[0], w[1]);
}

float Transforms::GetCarOrientation(geometry_msgs::Pose pose) {
  return std::atan2(2 * pose.orientation.w * pose.orientation.z, 1 - 2 * pose.orientation.z * pose.orientation.z);
}

float Transforms::CalcDist(std::pair<float, float> p1, std::pair<float, float> p2) {
  const double dx = p1.first - p2.first, dy = p1.second - p2.second;  // float differences promoted by pow(float, int)
  return std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
}
