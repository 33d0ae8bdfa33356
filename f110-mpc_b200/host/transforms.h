// Transforms — same static interface as the reference (include/f110-mpc/transforms.h:10-18) with the tf2 calls
// replaced by the arithmetic they perform (quaternion -> basis, basis -> quaternion round trip included,
// because the reference goes Pose -> tf2::Transform -> geometry_msgs::Transform -> tf2::Transform).
#pragma once
#include <utility>
#include "msgs.h"

namespace geometry_msgs {
struct Transform { Point translation; Quaternion rotation; };
struct TransformStamped { Transform transform; };
}  // namespace geometry_msgs

class Transforms {
 public:
  static std::pair<float, float> CarPointToWorldPoint(float x, float y, geometry_msgs::Pose& current_pose);
  static geometry_msgs::TransformStamped WorldToCarTransform(const geometry_msgs::Pose& pose);
  static std::pair<float, float> TransformPoint(std::pair<float, float> point, geometry_msgs::TransformStamped& transform_msg);
  static float GetCarOrientation(geometry_msgs::Pose pose);
  static float CalcDist(std::pair<float, float> p1, std::pair<float, float> p2);
  // rows 0,1 x cols 0,1 of the basis CarPointToWorldPoint rotates with (input of the device collision check)
  static void CarToWorldRotation(const geometry_msgs::Pose& pose, double R[4]);
};
