// Traj_Plan — table of constant-steer roll-outs ("mini-paths") in the car frame
// (reference include/f110-mpc/trajectory_planner.h, src/trajectory_planner.cpp:26-72).
#pragma once
#include <vector>
#include "model.h"
#include "msgs.h"

class Traj_Plan {
 public:
  explicit Traj_Plan(const f110::Params& params)
      : speed_max(params.speed_max), steer_max(params.steer_max), speed_discrete(params.speed_discrete),
        steer_discrete(params.steer_discrete), traj_discrete(params.traj_discrete), dt(params.dt_planner) {}
  virtual ~Traj_Plan() = default;
  std::vector<std::vector<State>> generate_traj_table();
  const std::vector<std::vector<State>>& table() const { return dwa_traj_table_; }
  // (paths x samples x 2) doubles for the device collision check
  std::vector<double> table_xy() const;
 private:
  double speed_max, steer_max;
  int speed_discrete, steer_discrete, traj_discrete;
  double dt;
  Model model_;
  std::vector<std::vector<State>> dwa_traj_table_;
};
