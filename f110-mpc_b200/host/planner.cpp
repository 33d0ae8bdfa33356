#include "planner.h"
#include <cmath>
#include <cstdint>
#include <limits>
#include "../../include/f110_mpc_b200.h"
#include "transforms.h"

bool MiniPathPlanner::Plan(geometry_msgs::Pose& pose, const OccGrid& grid, std::vector<State>* mini_path) {
  const auto& tab = table_.table();
  valid_traj_idx_.clear();
  best_trajectory_idx_ = best_global_idx_ = -1;
  if (tab.empty()) return false;
  const int paths = static_cast<int>(tab.size()), samples = static_cast<int>(tab[0].size());
  const std::vector<double> xy = table_.table_xy();
  double R[4];
  Transforms::CarToWorldRotation(pose, R);
  const double pose_xy[2] = {pose.position.x, pose.position.y};
  const float off[2] = {grid.offset().first, grid.offset().second};
  std::vector<uint8_t> valid(paths);
  std::vector<int32_t> free_count(paths);
  std::vector<float> end_world(2 * static_cast<std::size_t>(paths));
  // project.cpp:76-113 on the device
  if (f110_collision_check_host(1, paths, samples, grid.blocks(), grid.discrete(), grid.data(), off, R, pose_xy, xy.data(),
                                valid.data(), free_count.data(), end_world.data(), device_) != F110_OK)
    return false;
  for (int i = 0; i < paths; ++i)
    if (valid[i]) valid_traj_idx_.push_back(i);
  if (valid_traj_idx_.empty()) return false;
  best_global_idx_ = raceline_.get_best_global_idx(pose);             // project.cpp:122
  if (best_global_idx_ < 0) return false;
  const State& target = raceline_.waypoints_[best_global_idx_];
  double best = std::numeric_limits<double>::max();                  // project.cpp:125-136: strict <, first wins
  for (int i : valid_traj_idx_) {
    const double ex = end_world[2 * i], ey = end_world[2 * i + 1];
    const double d = std::pow(std::pow(ex - target.x(), 2) + std::pow(ey - target.y(), 2), 0.5);
    if (d < best) { best = d; best_trajectory_idx_ = i; }
  }
  mini_path->clear();
  for (const State& s : tab[best_trajectory_idx_]) {                 // project.cpp:145-149: world frame, ori = 0.0
    const std::pair<float, float> w = Transforms::CarPointToWorldPoint(s.x(), s.y(), pose);
    mini_path->emplace_back(w.first, w.second, 0.0);
  }
  return true;
}
