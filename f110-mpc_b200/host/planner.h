// MiniPathPlanner — the planning half of the reference's OdomCallback (src/project.cpp:73-157): collision-check
// every mini-path of the table against the occupancy grid (on the GPU, bit-exact), pick the surviving path
// whose end point is nearest the look-ahead point of the raceline, and return it in the world frame.
#pragma once
#include <vector>
#include "occupancy_grid.h"
#include "state.h"
#include "trajectory.h"
#include "trajectory_planner.h"

class MiniPathPlanner {
 public:
  MiniPathPlanner(Traj_Plan& table, Trajectory& raceline, int device = 0) : table_(table), raceline_(raceline), device_(device) {}
  // Returns false when no path is valid ("NO VALID TRAJS", project.cpp:115-119) or nothing lies ahead.
  bool Plan(geometry_msgs::Pose& current_pose, const OccGrid& grid, std::vector<State>* mini_path);
  const std::vector<int>& valid_traj_idx() const { return valid_traj_idx_; }
  int best_trajectory_idx() const { return best_trajectory_idx_; }
  int best_global_idx() const { return best_global_idx_; }
 private:
  Traj_Plan& table_;
  Trajectory& raceline_;
  int device_;
  std::vector<int> valid_traj_idx_;
  int best_trajectory_idx_ = -1, best_global_idx_ = -1;
};
