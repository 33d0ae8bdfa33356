// project — the reference's orchestrator (include/f110-mpc/project.h:25-77, src/project.cpp) without ROS: the
// callbacks are plain methods, the detached DriveLoop thread becomes DriveStep() that the caller invokes at the
// drive rate, and the planning block runs its collision check on the GPU (MiniPathPlanner).  The state machine
// is the reference's: plan a mini-path when none is held (no control update in that cycle), otherwise run one
// MPC cycle against it, dropping the path once the car is within 1.98 m of its end (project.cpp:180-186).
#pragma once
#include <mutex>
#include <string>
#include <vector>
#include "constraints.h"
#include "mpc.h"
#include "occupancy_grid.h"
#include "planner.h"
#include "trajectory.h"
#include "trajectory_planner.h"
#include "transforms.h"

class project {
 public:
  project(const f110::Params& params, int device = 0);
  virtual ~project() = default;

  bool LoadRaceline(const std::string& csv_path);                          // traj_read_.ReadCSV (project.cpp:34-35)
  void SetRaceline(const std::vector<std::pair<float, float>>& xy);
  void ScanCallback(const sensor_msgs::LaserScan& scan_msg);               // project.cpp:41-59
  void OdomCallback(const geometry_msgs::Pose& pose);                      // project.cpp:62-208
  Input GetNextInput();                                                    // project.cpp:210-218
  // One pass of the DriveLoop body (project.cpp:224-236): the input to publish now; advances the input index.
  // Returns false until a pose and a scan have been seen.
  bool DriveStep(Input* out);

  bool has_mini_path() const { return get_mini_path_; }
  const std::vector<State>& mini_path() const { return miniPath_; }
  int cycles_planned() const { return n_plans_; }
  int cycles_solved() const { return n_solves_; }
  MPC& mpc() { return mpc_; }

 private:
  bool first_pose_estimate_ = false, first_scan_estimate_ = false;
  geometry_msgs::Pose current_pose_;
  OccGrid occ_grid_;
  Constraints constraints_;
  Trajectory traj_read_;
  MPC mpc_;
  Traj_Plan traj_plan_;
  MiniPathPlanner planner_;
  std::vector<Input> current_inputs_;
  unsigned int inputs_idx_ = 0;
  std::mutex inputs_mutex_;  // the reference shares current_inputs_/inputs_idx_ between threads without one
  bool get_mini_path_ = false;
  std::vector<State> miniPath_;
  int n_plans_ = 0, n_solves_ = 0;
};
