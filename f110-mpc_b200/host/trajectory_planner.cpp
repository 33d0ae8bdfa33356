#include "trajectory_planner.h"

std::vector<std::vector<State>> Traj_Plan::generate_traj_table() {
  dwa_traj_table_.clear();
  const double step = 2 * steer_max / steer_discrete;
  for (int i = 0; i <= steer_discrete; ++i) {  // steer_discrete + 1 paths, -steer_max .. +steer_max
    Input in(speed_max, -steer_max + i * step);
    std::vector<State> path;
    State cur(0.0, 0.0, 0.0), nxt;
    path.push_back(cur);
    for (int k = 0; k < traj_discrete - 1; ++k) {  // traj_discrete points including the origin
      model_.simulate_dynamics(cur, in, dt, nxt);
      path.push_back(nxt);
      cur = nxt;
    }
    dwa_traj_table_.push_back(path);
  }
  return dwa_traj_table_;
}

std::vector<double> Traj_Plan::table_xy() const {
  std::vector<double> out;
  for (const auto& path : dwa_traj_table_)
    for (const State& s : path) { out.push_back(s.x()); out.push_back(s.y()); }
  return out;
}
