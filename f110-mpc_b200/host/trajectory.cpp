#include "trajectory.h"
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>

bool Trajectory::ReadCSV(const std::string& path) {
  std::ifstream in(path);
  if (!in.is_open()) return false;
  std::vector<std::pair<float, float>> xy;
  std::string sx, rest;
  while (std::getline(in, sx, ',')) {  // first column, then the remainder of the line (stof stops at the next comma)
    std::getline(in, rest);
    xy.emplace_back(std::stof(sx), std::stof(rest));
  }
  SetWaypointsXY(xy);
  return true;
}

void Trajectory::SetWaypointsXY(const std::vector<std::pair<float, float>>& xy) {
  waypoints_.clear();
  const std::size_t n = xy.size();
  for (unsigned int i = 0; i < n; i++) {
    // heading from the previous point; the 32-bit unsigned (i - 1) wraps at i = 0, exactly as in the reference
    const std::pair<float, float>& prev = xy[(i - 1) % n];
    const float heading = std::atan2(xy[i].second - prev.second, xy[i].first - prev.first);
    waypoints_.emplace_back(xy[i].first, xy[i].second, heading);
  }
}

int Trajectory::get_best_global_idx(geometry_msgs::Pose current_pose) {
  float best = std::numeric_limits<float>::max();
  int best_idx = -1;
  geometry_msgs::TransformStamped to_car = Transforms::WorldToCarTransform(current_pose);
  for (int i = 0; i < static_cast<int>(waypoints_.size()); ++i) {
    const std::pair<float, float> wp(waypoints_[i].x(), waypoints_[i].y());
    const std::pair<float, float> in_car = Transforms::TransformPoint(wp, to_car);
    if (in_car.first < 0) continue;  // behind the car
    const double dist = std::pow(std::pow(static_cast<double>(in_car.first), 2) + std::pow(static_cast<double>(in_car.second), 2), 0.5);
    const double off = std::abs(dist - lookahead);
    if (off < best) { best = off; best_idx = i; }
  }
  return best_idx;
}
