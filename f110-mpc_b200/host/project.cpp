#include "project.h"
#include <cstdio>

project::project(const f110::Params& prm, int device)
    : occ_grid_(prm), constraints_(prm), traj_read_(prm), mpc_(prm, device), traj_plan_(prm),
      planner_(traj_plan_, traj_read_, device) {
  traj_plan_.generate_traj_table();  // project.cpp:37
}

bool project::LoadRaceline(const std::string& csv_path) { return traj_read_.ReadCSV(csv_path); }
void project::SetRaceline(const std::vector<std::pair<float, float>>& xy) { traj_read_.SetWaypointsXY(xy); }

void project::ScanCallback(const sensor_msgs::LaserScan& scan_msg) {
  if (!first_pose_estimate_) return;
  if (!first_scan_estimate_) {
    first_scan_estimate_ = true;
    mpc_.UpdateScan(scan_msg);  // the MPC keeps the FIRST scan for good (project.cpp:45-49)
  }
  State origin(0.0, 0.0, 0.0);
  sensor_msgs::LaserScan copy = scan_msg;
  constraints_.FindHalfSpaces(origin, copy);  // project.cpp:51-54 (result only drawn in the reference)
  occ_grid_.FillOccGrid(current_pose_, scan_msg);
}

void project::OdomCallback(const geometry_msgs::Pose& pose) {
  current_pose_ = pose;
  first_pose_estimate_ = true;
  if (!get_mini_path_) {
    // planning cycle (project.cpp:73-157): no control update here
    std::vector<State> path;
    if (!planner_.Plan(current_pose_, occ_grid_, &path)) {
      std::fprintf(stderr, "NO VALID TRAJS\n");
      return;
    }
    miniPath_ = path;
    get_mini_path_ = true;
    ++n_plans_;
    return;
  }
  if (!first_scan_estimate_) return;  // project.cpp:167
  const float yaw = Transforms::GetCarOrientation(current_pose_);
  State current_state(current_pose_.position.x, current_pose_.position.y, yaw);
  Input input_to_pass = GetNextInput();
  input_to_pass.set_v(4.5);  // project.cpp:170
  const std::pair<float, float> end_point(miniPath_.back().x(), miniPath_.back().y());
  const std::pair<float, float> car_point(current_pose_.position.x, current_pose_.position.y);
  if (Transforms::CalcDist(car_point, end_point) < 1.98) {  // project.cpp:182
    get_mini_path_ = false;
    miniPath_.clear();  // the Update below then sees an empty trajectory (skipped by MPC::Update here)
  }
  mpc_.Update(current_state, input_to_pass, miniPath_);
  if (!miniPath_.empty()) ++n_solves_;
  std::lock_guard<std::mutex> lock(inputs_mutex_);
  current_inputs_ = mpc_.solved_trajectory();  // project.cpp:190-191
  inputs_idx_ = 0;
}

Input project::GetNextInput() {
  std::lock_guard<std::mutex> lock(inputs_mutex_);
  if (inputs_idx_ >= current_inputs_.size()) return Input(0.5, 0.0);  // "ran out of QP soln" (project.cpp:212-216)
  return current_inputs_[inputs_idx_];
}

bool project::DriveStep(Input* out) {
  if (!(first_pose_estimate_ && first_scan_estimate_)) return false;
  *out = GetNextInput();
  std::lock_guard<std::mutex> lock(inputs_mutex_);
  inputs_idx_++;  // project.cpp:234
  return true;
}
