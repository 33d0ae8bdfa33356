// C++ unit checks of the host classes that mirror the reference's interfaces (f110-mpc_b200/host/), run without a GPU by
// tests/test_host_cpu.py.  The reference ships no tests; every expectation below is derived from its source (file:line given).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "constraints.h"
#include "cost.h"
#include "model.h"
#include "msgs.h"
#include "occupancy_grid.h"
#include "state.h"
#include "trajectory.h"
#include "trajectory_planner.h"
#include "transforms.h"

static int failures = 0;
#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

int main(int argc, char** argv) {
  const std::string tmp = argc > 1 ? argv[1] : "/tmp";
  // ---- State / Input (state.cpp:3-67, input.cpp:3-43)
  State s(1.5, -2.25, 0.3);
  CHECK(s.x() == 1.5 && s.y() == -2.25 && s.ori() == 0.3 && s.size() == 3);
  const f110::Vector sv = s.StateToVector();
  CHECK(sv.size() == 3 && sv(0) == 1.5 && sv(1) == -2.25 && sv(2) == 0.3);
  CHECK(s.GetPair().first == 1.5f && s.GetPair().second == -2.25f);
  s.set_x(7); s.set_y(8); s.set_ori(9);
  CHECK(s.x() == 7 && s.y() == 8 && s.ori() == 9);
  Input in(4.5, 0.1);
  CHECK(in.v() == 4.5 && in.steer_ang() == 0.1 && in.InputToVector()(1) == 0.1);

  // ---- Params defaults = params.yaml with the reference's destination types; FromYaml overrides
  f110::Params prm;
  CHECK(prm.horizon == 30 && prm.dt == 0.01f && prm.q0 == 10.0 && prm.q2 == 0.0 && prm.r1 == 5.0 && prm.umax == 4.5f && prm.umin == 3.0f);
  {
    std::ofstream y(tmp + "/params_unit.yaml");
    y << "# comment\nhorizon: 12\nq0: 3.5\ndt: 0.02\numax: 5.0\nlookahead: 1.75\nsteer_rate_max: 3.2\nunknown_key: 9\nsteer_discrete: 19\n";
  }
  const f110::Params py = f110::Params::FromYaml(tmp + "/params_unit.yaml");
  CHECK(py.horizon == 12 && py.q0 == 3.5 && py.dt == 0.02f && py.umax == 5.0f && py.speed_max == 5.0 && py.lookahead == 1.75f);
  CHECK(py.steer_rate_max == 3.2 && py.steer_discrete == 19 && py.q1 == 10.0);

  // ---- Cost (cost.cpp:6-21; mpc.cpp:20-24): Q = diag(q0,q1,q2), R = diag(r0,r1)
  Cost cost(f110::Matrix::Diagonal({prm.q0, prm.q1, prm.q2}), f110::Matrix::Diagonal({prm.r0, prm.r1}));
  CHECK(cost.q()(0, 0) == 10.0 && cost.q()(1, 1) == 10.0 && cost.q()(2, 2) == 0.0 && cost.q()(0, 1) == 0.0);
  CHECK(cost.r()(0, 0) == 0.10 && cost.r()(1, 1) == 5.0 && cost.r().rows() == 2);

  // ---- Constraints (constraints.cpp:14-21): state box +-INFTY (never stacked), input box with float-narrowed steering limits
  Constraints con(prm);
  CHECK(con.x_max()(0) == 1e30 && con.x_min()(2) == -1e30);
  CHECK(con.u_max()(0) == (double)4.5f && con.u_max()(1) == (double)0.43f && con.u_min()(0) == (double)3.0f && con.u_min()(1) == (double)-0.43f);

  // ---- Model (model.cpp:30-76)
  Model m;
  State s0(0.5, -1.0, 0.7);
  Input u0(4.5, 0.2);
  const double dt = (double)0.01f, L = (double)0.3302f;
  m.Linearize(s0, u0, dt);
  CHECK(m.A()(0, 0) == 1.0 && m.A()(0, 2) == -1.0 * 4.5 * std::sin(0.7) * dt && m.A()(1, 2) == 4.5 * std::cos(0.7) * dt && m.A()(2, 2) == 1.0);
  CHECK(m.B()(0, 0) == std::cos(0.7) * dt && m.B()(1, 0) == std::sin(0.7) * dt && m.B()(0, 1) == 0.0);
  CHECK(std::fabs(m.B()(2, 0) - std::tan(0.2) * dt / L) < 1e-15 && std::fabs(m.B()(2, 1) - 4.5 * dt / (L * std::cos(0.2) * std::cos(0.2))) < 1e-15);
  State s1;
  m.simulate_dynamics(s0, u0, 0.01, s1);   // x+ = x + dt [v cos, v sin, v tan(delta) / 0.35]  (model.cpp:68-72: its own 0.35 wheelbase)
  CHECK(std::fabs(s1.x() - (0.5 + 0.01 * 4.5 * std::cos(0.7))) < 1e-15 && std::fabs(s1.y() - (-1.0 + 0.01 * 4.5 * std::sin(0.7))) < 1e-15);
  CHECK(std::fabs(s1.ori() - (0.7 + 0.01 * 4.5 * std::tan(0.2) / 0.35)) < 1e-15);

  // ---- Traj_Plan (trajectory_planner.cpp:26-72): steer_discrete + 1 roll-outs of traj_discrete points from the origin
  Traj_Plan tp(prm);
  const auto table = tp.generate_traj_table();
  CHECK((int)table.size() == prm.steer_discrete + 1 && (int)table[0].size() == prm.traj_discrete);
  CHECK(table[0][0].x() == 0.0 && table[0][0].y() == 0.0 && table[0][0].ori() == 0.0);
  const int mid = prm.steer_discrete / 2;   // delta = -0.4 + i * 0.8 / steer_discrete: the middle path is (nearly) straight
  CHECK(std::fabs(table[mid][49].x() - 49 * 0.045) < 1e-6 && std::fabs(table[mid][49].y()) < 1e-6);
  CHECK(table[0][49].y() < 0 && table[prm.steer_discrete][49].y() > 0);   // right turn first, left turn last
  CHECK(tp.table_xy().size() == table.size() * 50 * 2);

  // ---- OccGrid (occupancy_grid.cpp:3-11, 27-33, 55-101, 165-168)
  OccGrid grid(prm);
  CHECK(grid.blocks() == 100 && grid.size() == 10 && grid.discrete() == 0.1f);
  CHECK(grid.InGrid(0, 0) && grid.InGrid(99, 99) && !grid.InGrid(100, 0) && !grid.InGrid(0, -1));
  geometry_msgs::Pose pose;   // identity pose at the origin: offset = (0.275, 0)
  sensor_msgs::LaserScan scan;
  scan.angle_min = -2.35f; scan.angle_max = 2.35f; scan.angle_increment = 4.7f / 1079;
  scan.ranges.assign(1080, 30.0f);          // far outside the 10 m grid
  scan.ranges[540] = 2.0f;                  // one return (almost) straight ahead
  grid.FillOccGrid(pose, scan);
  CHECK(grid.offset().first == 0.275f && grid.offset().second == 0.0f);
  const auto origin = grid.WorldToOccupancy(0.275f, 0.0f);
  CHECK(origin.first == 50 && origin.second == 50);
  CHECK(grid.WorldToOccupancy(0.275f - 5.05f, 0.0f).first == 0);   // (int) truncates toward zero: -0.5 -> 0, still "in grid"
  int occupied = 0;
  for (int c = 0; c < 100; ++c) for (int r = 0; r < 100; ++r) occupied += grid.IsOccupied((float)c, (float)r) ? 1 : 0;
  CHECK(occupied >= 9 && occupied <= 25);   // one beam stamps a 4 x 4 dilation block (occupancy_grid.cpp:76-84)
  const auto hit = grid.WorldToOccupancy(0.275f + 2.0f, 0.0f);
  CHECK(grid.IsOccupied((float)hit.second, (float)hit.first));   // (row, col): the "SWAP" of project.cpp:89-92
  CHECK(!grid.IsOccupied(50.f, 50.f) && grid.CartesianInGrid(1.0f, 1.0f) && !grid.CartesianInGrid(20.0f, 0.0f));

  // ---- Transforms (transforms.cpp): yaw of a pure-z quaternion, car point -> world, world -> car round trip
  geometry_msgs::Pose p2;
  p2.position.x = 1.0; p2.position.y = 2.0;
  p2.orientation.z = std::sin(0.25); p2.orientation.w = std::cos(0.25);   // yaw 0.5
  CHECK(std::fabs(Transforms::GetCarOrientation(p2) - 0.5f) < 1e-6f);
  const auto w = Transforms::CarPointToWorldPoint(1.0f, 0.0f, p2);
  CHECK(std::fabs(w.first - (1.0f + std::cos(0.5f))) < 1e-6f && std::fabs(w.second - (2.0f + std::sin(0.5f))) < 1e-6f);
  auto tf = Transforms::WorldToCarTransform(p2);
  const auto back = Transforms::TransformPoint(w, tf);
  CHECK(std::fabs(back.first - 1.0f) < 1e-5f && std::fabs(back.second) < 1e-5f);
  CHECK(std::fabs(Transforms::CalcDist({0.f, 0.f}, {3.f, 4.f}) - 5.f) < 1e-6f);

  // ---- Trajectory::ReadCSV + look-ahead (trajectory.cpp:18-55, 81-108): columns 0-1 only, parsed as float
  {
    std::ofstream c(tmp + "/line_unit.csv");
    for (int i = 0; i < 40; ++i) c << (0.25 * i) << "," << 0.0 << ",4.0,0.0,0.0,0.0\n";
  }
  Trajectory traj(prm);
  CHECK(traj.ReadCSV(tmp + "/line_unit.csv") && traj.waypoints_.size() == 40);
  CHECK(traj.waypoints_[8].x() == (double)2.0f && traj.waypoints_[8].y() == 0.0);
  geometry_msgs::Pose p3;   // at x = 1, facing +x: waypoint nearest to the 2.5 m look-ahead is x = 3.5 (index 14)
  p3.position.x = 1.0;
  CHECK(traj.get_best_global_idx(p3) == 14);
  p3.position.x = 100.0;    // everything is behind the car
  CHECK(traj.get_best_global_idx(p3) == -1);
  CHECK(!traj.ReadCSV(tmp + "/does_not_exist.csv"));

  std::printf(failures ? "%d check(s) failed\n" : "host unit checks passed\n", failures);
  return failures ? 1 : 0;
}
