// C shims over the C++ host classes so the Python harness (tests/, bench.py) can drive them through ctypes.
// Poses travel as 7 doubles (px, py, pz, qx, qy, qz, qw); scans as (angle_min, angle_max, angle_increment,
// ranges, n).  Not part of the drop-in boundary (that is include/f110_mpc_b200.h); these are conveniences.
#include <cmath>
#include <cstring>
#include <memory>
#include "constraints.h"
#include "model.h"
#include "mpc.h"
#include "occupancy_grid.h"
#include "planner.h"
#include "project.h"
#include "trajectory.h"
#include "trajectory_planner.h"
#include "transforms.h"

namespace {
geometry_msgs::Pose pose_of(const double* p) {
  geometry_msgs::Pose q;
  q.position.x = p[0]; q.position.y = p[1]; q.position.z = p[2];
  q.orientation.x = p[3]; q.orientation.y = p[4]; q.orientation.z = p[5]; q.orientation.w = p[6];
  return q;
}
sensor_msgs::LaserScan scan_of(float amin, float amax, float inc, const float* r, int n) {
  sensor_msgs::LaserScan s;
  s.angle_min = amin; s.angle_max = amax; s.angle_increment = inc;
  s.ranges.assign(r, r + n);
  return s;
}
struct MpcBox {
  f110::Params prm;
  std::unique_ptr<MPC> mpc;
};
}  // namespace

extern "C" {

void f110h_linearize(double ori, double v, double steer, double dt, double* A, double* B, double* C) {
  Model m;
  State s(0, 0, ori);
  Input in(v, steer);
  m.Linearize(s, in, dt);
  std::memcpy(A, m.A().data(), 9 * sizeof(double));
  std::memcpy(B, m.B().data(), 6 * sizeof(double));
  std::memcpy(C, m.C().data(), 3 * sizeof(double));
}

int f110h_traj_table(int steer_discrete, int traj_discrete, double* out_xyo) {
  f110::Params prm;
  prm.steer_discrete = steer_discrete; prm.traj_discrete = traj_discrete;
  Traj_Plan tp(prm);
  const auto tab = tp.generate_traj_table();
  std::size_t k = 0;
  for (const auto& path : tab)
    for (const State& s : path) { out_xyo[k++] = s.x(); out_xyo[k++] = s.y(); out_xyo[k++] = s.ori(); }
  return static_cast<int>(tab.size());
}

void f110h_car_to_world_R(const double* pose7, double* R4) { Transforms::CarToWorldRotation(pose_of(pose7), R4); }

int f110h_fill_grid(const double* pose7, float amin, float amax, float inc, const float* ranges, int n, float* grid, float* offset) {
  f110::Params prm;
  OccGrid g(prm);
  g.FillOccGrid(pose_of(pose7), scan_of(amin, amax, inc, ranges, n));
  std::memcpy(grid, g.data(), sizeof(float) * g.blocks() * g.blocks());
  offset[0] = g.offset().first; offset[1] = g.offset().second;
  return g.blocks();
}

int f110h_find_half_spaces(const double* state3, float amin, float amax, float inc, const float* ranges, int n, double* l1, double* l2, int* lohi) {
  f110::Params prm;
  Constraints c(prm);
  State s(state3[0], state3[1], state3[2]);
  sensor_msgs::LaserScan scan = scan_of(amin, amax, inc, ranges, n);
  const bool ok = c.FindHalfSpaces(s, scan);
  lohi[0] = c.best_gap().first; lohi[1] = c.best_gap().second;
  if (!ok) return 0;
  for (int j = 0; j < 3; ++j) { l1[j] = c.l1()(j); l2[j] = c.l2()(j); }
  return 1;
}

int f110h_best_global_idx(const float* wp_xy, int W, const double* pose7, double* headings_out) {
  f110::Params prm;
  Trajectory t(prm);
  std::vector<std::pair<float, float>> xy(W);
  for (int i = 0; i < W; ++i) xy[i] = {wp_xy[2 * i], wp_xy[2 * i + 1]};
  t.SetWaypointsXY(xy);
  if (headings_out) for (int i = 0; i < W; ++i) headings_out[i] = t.waypoints_[i].ori();
  return t.get_best_global_idx(pose_of(pose7));
}

// ---- MPC object -------------------------------------------------------------------------------------------
void* f110h_mpc_create_rate(int horizon, int gap_mode, double steer_rate_max, int device);
void* f110h_mpc_create(int horizon, int gap_mode, int device) { return f110h_mpc_create_rate(horizon, gap_mode, 0.0, device); }
void* f110h_mpc_create_rate(int horizon, int gap_mode, double steer_rate_max, int device) {
  try {
    std::unique_ptr<MpcBox> b(new MpcBox());
    b->prm.horizon = horizon; b->prm.gap_mode = gap_mode; b->prm.steer_rate_max = steer_rate_max;
    b->mpc.reset(new MPC(b->prm, device));
    return b.release();
  } catch (const std::exception&) {
    return nullptr;
  }
}
void f110h_mpc_destroy(void* h) { delete static_cast<MpcBox*>(h); }
void f110h_mpc_update_scan(void* h, float amin, float amax, float inc, const float* ranges, int n) {
  static_cast<MpcBox*>(h)->mpc->UpdateScan(scan_of(amin, amax, inc, ranges, n));
}
// desired: K x 3.  Outputs: inputs (N x 2, the solved trajectory), primal x, dual y, l1l2 (6).  Returns the number of inputs.
int f110h_mpc_update(void* h, const double* state3, const double* input2, const double* desired, int K, double* inputs_out,
                     double* x_out, double* y_out, int* status, int* iters, double* l1l2) {
  MPC& m = *static_cast<MpcBox*>(h)->mpc;
  std::vector<State> des;
  for (int k = 0; k < K; ++k) des.emplace_back(desired[3 * k], desired[3 * k + 1], desired[3 * k + 2]);
  m.Update(State(state3[0], state3[1], state3[2]), Input(input2[0], input2[1]), des);
  const std::vector<Input> sol = m.solved_trajectory();
  for (std::size_t k = 0; k < sol.size(); ++k) { inputs_out[2 * k] = sol[k].v(); inputs_out[2 * k + 1] = sol[k].steer_ang(); }
  if (x_out) std::memcpy(x_out, m.last_primal().data(), sizeof(double) * m.num_variables());
  if (y_out) std::memcpy(y_out, m.last_dual().data(), sizeof(double) * m.num_constraints());
  *status = m.last_status(); *iters = m.last_iterations();
  if (l1l2) for (int j = 0; j < 3; ++j) { l1l2[j] = m.constraints().l1()(j); l1l2[3 + j] = m.constraints().l2()(j); }
  return static_cast<int>(sol.size());
}

// ---- planning cycle: grid fill (host) -> collision check (device) -> selection (host) --------------------------
// Returns the chosen table index (or -1); mini_path_out: traj_discrete x 3; valid_out: steer_discrete + 1 flags.
int f110h_plan(int steer_discrete, int traj_discrete, const double* pose7, float amin, float amax, float inc, const float* ranges,
               int n, const float* wp_xy, int W, int device, double* mini_path_out, unsigned char* valid_out, int* best_global) {
  f110::Params prm;
  prm.steer_discrete = steer_discrete; prm.traj_discrete = traj_discrete;
  OccGrid grid(prm);
  geometry_msgs::Pose pose = pose_of(pose7);
  grid.FillOccGrid(pose, scan_of(amin, amax, inc, ranges, n));
  Traj_Plan tp(prm);
  tp.generate_traj_table();
  Trajectory race(prm);
  std::vector<std::pair<float, float>> xy(W);
  for (int i = 0; i < W; ++i) xy[i] = {wp_xy[2 * i], wp_xy[2 * i + 1]};
  race.SetWaypointsXY(xy);
  MiniPathPlanner planner(tp, race, device);
  std::vector<State> path;
  const bool ok = planner.Plan(pose, grid, &path);
  for (int i = 0; i <= steer_discrete; ++i) valid_out[i] = 0;
  for (int i : planner.valid_traj_idx()) valid_out[i] = 1;
  *best_global = planner.best_global_idx();
  if (!ok) return -1;
  for (std::size_t k = 0; k < path.size(); ++k) { mini_path_out[3 * k] = path[k].x(); mini_path_out[3 * k + 1] = path[k].y(); mini_path_out[3 * k + 2] = path[k].ori(); }
  return planner.best_trajectory_idx();
}

// ---- closed loop: the `project` orchestrator driving a simulated car (Model::simulate_dynamics) ------------------------
// ticks of `dt_tick` seconds; OdomCallback every tick, DriveStep every `drive_every` ticks, ScanCallback with a
// constant scan every `scan_every` ticks.  Outputs per tick: pose (x, y, yaw) and the applied input (v, steer).
// Returns the number of MPC cycles solved.
int f110h_closed_loop(int ticks, double dt_tick, int drive_every, int scan_every, const float* wp_xy, int W, const double* start_xyyaw,
                      float amin, float amax, float inc, const float* ranges, int n, int device, double* traj_out, int* plans_out) {
  try {
    f110::Params prm;
    project node(prm, device);
    std::vector<std::pair<float, float>> xy(W);
    for (int i = 0; i < W; ++i) xy[i] = {wp_xy[2 * i], wp_xy[2 * i + 1]};
    node.SetRaceline(xy);
    const sensor_msgs::LaserScan scan = scan_of(amin, amax, inc, ranges, n);
    Model plant;
    State car(start_xyyaw[0], start_xyyaw[1], start_xyyaw[2]);
    Input applied(0.5, 0.0);
    for (int t = 0; t < ticks; ++t) {
      geometry_msgs::Pose pose;
      pose.position.x = car.x(); pose.position.y = car.y();
      pose.orientation.z = std::sin(car.ori() / 2.0); pose.orientation.w = std::cos(car.ori() / 2.0);
      node.OdomCallback(pose);
      if (t % scan_every == 0) node.ScanCallback(scan);
      if (t % drive_every == 0) {
        Input in;
        if (node.DriveStep(&in)) applied = in;
      }
      traj_out[5 * t] = car.x(); traj_out[5 * t + 1] = car.y(); traj_out[5 * t + 2] = car.ori();
      traj_out[5 * t + 3] = applied.v(); traj_out[5 * t + 4] = applied.steer_ang();
      State next;
      plant.simulate_dynamics(car, applied, dt_tick, next);
      car = next;
    }
    if (plans_out) *plans_out = node.cycles_planned();
    return node.cycles_solved();
  } catch (const std::exception&) {
    return -1;
  }
}

}  // extern "C"
