// Trajectory — CSV raceline and the look-ahead point (reference include/f110-mpc/trajectory.h, src/trajectory.cpp).
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "msgs.h"
#include "state.h"
#include "transforms.h"

class Trajectory {
 public:
  explicit Trajectory(const f110::Params& params) : lookahead(params.lookahead) {}
  virtual ~Trajectory() = default;
  bool ReadCSV(const std::string& path);                           // trajectory.cpp:18-55 (full path instead of a package name)
  void SetWaypointsXY(const std::vector<std::pair<float, float>>& xy);  // the post-parse half of ReadCSV
  int get_best_global_idx(geometry_msgs::Pose current_pose);       // trajectory.cpp:81-108; -1 when nothing is ahead
  std::vector<State> waypoints_;
 private:
  float lookahead;
};
