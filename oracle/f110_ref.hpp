// ORACLE — TEST INFRASTRUCTURE ONLY (see osqp_restated.hpp header). PARITY UNPINNED by the
// reference (it has no tests / golden vectors).
//
// CPU restatement of the f110-mpc per-cycle path, ROS-free, with the reference's C++
// destination types (float vs double) kept literally.  Every function cites the
// reference file:line it follows (paths relative to /root/reference).
//
// Choices where the reference is ambiguous or undefined (SURVEY.md §8c, a13'):
//  * unqualified cos/sin/atan2/sqrt on float arguments bind to the float overloads
//    (libstdc++ <math.h> wrapper is in the include closure) -> cosf/sinf here;
//  * FindHalfSpaces with no gap (best_lo == -1, ranges[-1] UB) -> returns false, rows loose;
//  * desired trajectory shorter than the horizon -> index clamped to the last element;
//  * get_best_global_idx with nothing ahead (-1, .at(-1) throws) -> returns -1.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "osqp_restated.hpp"

namespace f110_ref {

// ---- params.yaml with the C++ destination types ---------------------------
struct Params {
  double q0 = 10.0, q1 = 10.0, q2 = 0.0;       // params.yaml:1-3  -> mpc.cpp:12-14 (double)
  double r0 = 0.10, r1 = 5.0;                  // params.yaml:5-6  -> mpc.cpp:15-16 (double)
  int horizon = 30;                            // params.yaml:12   -> mpc.cpp:5
  float dt_f = 0.01f;                          // params.yaml:13   -> mpc.cpp:6 (float dt_)
  double dt_plan = 0.01;                       // params.yaml:13   -> trajectory_planner.cpp:10 (double)
  int occ_size = 10;                           // params.yaml:16
  float occ_discrete = 0.1f;                   // params.yaml:17
  float occ_dilation = 0.15f;                  // params.yaml:18
  double des_vel = 4.5, des_steer = 0.0;       // params.yaml:42-43
  float umax = 4.5f, umin = 3.0f;              // params.yaml:46-47 -> constraints.cpp:7-8 (float)
  float follow_gap_thresh = 3.0f;              // params.yaml:49
  float fov_divider = 1.5f;                    // params.yaml:51
  float buffer = 3.0f;                         // params.yaml:52
  int steer_discrete = 30;                     // params.yaml:55
  double steer_max = 0.4;                      // params.yaml:56
  int traj_discrete = 50;                      // params.yaml:57
  float lookahead = 2.5f;                      // params.yaml:59
  double speed_max = 4.5;                      // params.yaml:46 -> trajectory_planner.cpp:5 (double)
};

struct State3 { double x = 0, y = 0, ori = 0; };
struct Input2 { double v = 0, steer = 0; };
struct Pose {  // geometry_msgs::Pose: all doubles
  double px = 0, py = 0, pz = 0, qx = 0, qy = 0, qz = 0, qw = 1;
};
struct Scan {  // sensor_msgs::LaserScan: float32 fields
  float angle_min = 0, angle_max = 0, angle_increment = 0;
  const float* ranges = nullptr;
  int n_ranges = 0;
};

// ---- Model (src/model.cpp) ---------------------------------------------------------------
// Model::Linearize, model.cpp:30-59.  A row-major 3x3, B row-major 3x2, C 3.
inline void linearize(double ori, double v, double steer, double dt, double* A, double* B, double* C) {
  float L = 0.3302f;  // model.cpp:32
  for (int k = 0; k < 9; ++k) A[k] = 0.0;
  for (int k = 0; k < 6; ++k) B[k] = 0.0;
  for (int k = 0; k < 3; ++k) C[k] = 0.0;
  A[0 * 3 + 2] = -1 * v * std::sin(ori) * dt;                       // :42
  A[1 * 3 + 2] = v * std::cos(ori) * dt;                            // :43
  A[0] = 1; A[4] = 1; A[8] = 1;                                     // :44-46
  B[0 * 2 + 0] = std::cos(ori) * dt;                                // :48
  B[1 * 2 + 0] = std::sin(ori) * dt;                                // :49
  B[2 * 2 + 0] = std::tan(steer) * dt / L;                          // :50
  B[2 * 2 + 1] = v * std::pow(std::cos(steer), -2) * dt / L;        // :51
  C[0] = v * ori * std::sin(ori) * dt;                              // :53
  C[1] = -1 * v * ori * std::cos(ori) * dt;                         // :54
  C[2] = -1 * steer * v * std::pow(std::cos(steer), -2) * dt / L;   // :55
}

// Model::simulate_dynamics, model.cpp:61-76 (CAR_LENGTH = 0.35, model.cpp:2)
inline State3 simulate_dynamics(const State3& s, const Input2& in, double dt) {
  const double CAR_LENGTH = 0.35;
  double d0 = in.v * std::cos(s.ori);
  double d1 = in.v * std::sin(s.ori);
  double d2 = std::tan(in.steer) * in.v / CAR_LENGTH;
  State3 o;
  o.x = s.x + d0 * dt;
  o.y = s.y + d1 * dt;
  o.ori = s.ori + d2 * dt;
  return o;
}

// Traj_Plan::generate_traj_table, trajectory_planner.cpp:26-72.
// Output: (steer_discrete+1) paths x traj_discrete states, base_link frame.
inline std::vector<std::vector<State3>> generate_traj_table(const Params& p) {
  std::vector<std::vector<State3>> table;
  double ds = 2 * +p.steer_max / p.steer_discrete;  // :30
  for (int i = 0; i < p.steer_discrete + 1; ++i) {  // :39
    std::vector<State3> traj;
    double steer = -p.steer_max + i * ds;           // :43
    Input2 in{p.speed_max, steer};
    State3 s, ns;
    for (int k = 0; k < p.traj_discrete - 1; ++k) { // :52
      if (k == 0) traj.push_back(s);
      ns = simulate_dynamics(s, in, p.dt_plan);
      traj.push_back(ns);
      s = ns;
    }
    table.push_back(traj);
  }
  return table;
}

// ---- tf2 restated [memory] (SURVEY.md §8c "tf2 math to restate") ----------------------------
struct Mat3 { double m[3][3]; };
// tf2::Matrix3x3::setRotation(q)
inline Mat3 tf2_set_rotation(double x, double y, double z, double w) {
  double d = x * x + y * y + z * z + w * w;
  double s = 2.0 / d;
  double xs = x * s, ys = y * s, zs = z * s;
  double wx = w * xs, wy = w * ys, wz = w * zs;
  double xx = x * xs, xy = x * ys, xz = x * zs;
  double yy = y * ys, yz = y * zs, zz = z * zs;
  Mat3 r;
  r.m[0][0] = 1.0 - (yy + zz); r.m[0][1] = xy - wz;         r.m[0][2] = xz + wy;
  r.m[1][0] = xy + wz;         r.m[1][1] = 1.0 - (xx + zz); r.m[1][2] = yz - wx;
  r.m[2][0] = xz - wy;         r.m[2][1] = yz + wx;         r.m[2][2] = 1.0 - (xx + yy);
  return r;
}
// tf2::Matrix3x3::getRotation(q) (trace / largest-diagonal branches)
inline void tf2_get_rotation(const Mat3& a, double* q /*x,y,z,w*/) {
  double trace = a.m[0][0] + a.m[1][1] + a.m[2][2];
  double temp[4];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    temp[3] = s * 0.5;
    s = 0.5 / s;
    temp[0] = (a.m[2][1] - a.m[1][2]) * s;
    temp[1] = (a.m[0][2] - a.m[2][0]) * s;
    temp[2] = (a.m[1][0] - a.m[0][1]) * s;
  } else {
    int i = a.m[0][0] < a.m[1][1] ? (a.m[1][1] < a.m[2][2] ? 2 : 1) : (a.m[0][0] < a.m[2][2] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0);
    temp[i] = s * 0.5;
    s = 0.5 / s;
    temp[3] = (a.m[k][j] - a.m[j][k]) * s;
    temp[j] = (a.m[j][i] + a.m[i][j]) * s;
    temp[k] = (a.m[k][i] + a.m[i][k]) * s;
  }
  q[0] = temp[0]; q[1] = temp[1]; q[2] = temp[2]; q[3] = temp[3];
}
// Rotation actually applied by Transforms::CarPointToWorldPoint (transforms.cpp:3-20):
// fromMsg(pose) -> toMsg (matrix->quaternion) -> doTransform (quaternion->matrix).
inline Mat3 car_to_world_rotation(const Pose& p) {
  Mat3 r0 = tf2_set_rotation(p.qx, p.qy, p.qz, p.qw);
  double q[4];
  tf2_get_rotation(r0, q);
  return tf2_set_rotation(q[0], q[1], q[2], q[3]);
}
// The bit-exact part of CarPointToWorldPoint given R (row 0 and row 1 of the basis):
// worldPoint = basis * (x, y, 0); return float(worldPoint.x + float(pose.x)), ... (transforms.cpp:13-19)
inline std::pair<float, float> car_point_to_world(float x, float y, const double R[4], double pose_x, double pose_y) {
  double cx = x, cy = y, cz = 0;
  // tf2::Vector3::dot: m[0]*v[0] + m[1]*v[1] + m[2]*v[2]; the z term is R?2 * 0 = (+/-)0
  double wx = R[0] * cx + R[1] * cy + 0.0 * cz;
  double wy = R[2] * cx + R[3] * cy + 0.0 * cz;
  float carPoseX = (float)pose_x;  // :17
  float carPoseY = (float)pose_y;  // :18
  return std::pair<float, float>((float)(wx + carPoseX), (float)(wy + carPoseY));  // :19
}
// Transforms::GetCarOrientation, transforms.cpp:46-49 (returns float)
inline float car_orientation(const Pose& p) {
  return (float)std::atan2(2 * p.qw * p.qz, 1 - 2 * p.qz * p.qz);
}
// Transforms::CalcDist, transforms.cpp:51-55: pow(float,int) promotes to double; sqrt(double); narrowed.
inline float calc_dist(std::pair<float, float> p1, std::pair<float, float> p2) {
  float dist = (float)std::sqrt(std::pow((double)(p1.first - p2.first), 2) + std::pow((double)(p1.second - p2.second), 2));
  return dist;
}
// Transforms::WorldToCarTransform (transforms.cpp:22-31) + TransformPoint (:33-44)
struct WorldToCar { Mat3 R; double tx, ty; };
inline WorldToCar world_to_car(const Pose& p) {
  Mat3 b = tf2_set_rotation(p.qx, p.qy, p.qz, p.qw);
  Mat3 inv;  // transpose
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) inv.m[r][c] = b.m[c][r];
  double ox = -p.px, oy = -p.py, oz = -p.pz;
  double t[3];
  for (int r = 0; r < 3; ++r) t[r] = inv.m[r][0] * ox + inv.m[r][1] * oy + inv.m[r][2] * oz;
  double q[4];
  tf2_get_rotation(inv, q);                       // toMsg
  WorldToCar w;
  w.R = tf2_set_rotation(q[0], q[1], q[2], q[3]);  // doTransform's fromMsg
  w.tx = t[0]; w.ty = t[1];
  return w;
}
inline std::pair<float, float> transform_point(std::pair<float, float> pt, const WorldToCar& w) {
  double x = pt.first, y = pt.second, z = 0;
  double ox = w.R.m[0][0] * x + w.R.m[0][1] * y + w.R.m[0][2] * z;
  double oy = w.R.m[1][0] * x + w.R.m[1][1] * y + w.R.m[1][2] * z;
  ox += w.tx;  // :41
  oy += w.ty;  // :42
  return std::pair<float, float>((float)ox, (float)oy);
}

// ---- Trajectory (src/trajectory.cpp) ---------------------------------------------------------
// ReadCSV heading rule, trajectory.cpp:40-53: inputs are the float-parsed first two columns.
inline std::vector<State3> waypoints_from_xy(const std::vector<std::pair<float, float>>& temp) {
  std::vector<State3> wp;
  for (unsigned int i = 0; i < temp.size(); i++) {
    float prev_x = temp[(i - 1) % temp.size()].first;   // unsigned wrap: i=0 -> (2^32-1) % size
    float prev_y = temp[(i - 1) % temp.size()].second;
    float x = temp[i].first, y = temp[i].second;
    float ori = std::atan2(y - prev_y, x - prev_x);     // atan2f (float overload)
    State3 s; s.x = x; s.y = y; s.ori = ori;
    wp.push_back(s);
  }
  return wp;
}
// Trajectory::get_best_global_idx, trajectory.cpp:81-108
inline int get_best_global_idx(const std::vector<State3>& waypoints, const Pose& pose, float lookahead) {
  float minDistance = std::numeric_limits<float>::max();
  int closest_idx = -1;
  WorldToCar w2c = world_to_car(pose);
  for (int i = 0; i < (int)waypoints.size(); i++) {
    std::pair<float, float> point((float)waypoints[i].x, (float)waypoints[i].y);
    std::pair<float, float> tp = transform_point(point, w2c);
    if (tp.first < 0) continue;
    double distance = std::pow(std::pow((double)tp.first, 2) + std::pow((double)tp.second, 2), 0.5);
    double lookahead_diff = std::abs(distance - lookahead);
    if (lookahead_diff < minDistance) {
      minDistance = (float)lookahead_diff;
      closest_idx = i;
    }
  }
  return closest_idx;
}

// ---- OccGrid (src/occupancy_grid.cpp) ---------------------------------------------------------
struct OccGrid {
  int size_ = 10;
  float discrete_ = 0.1f;
  float dilation_ = 0.15f;
  int grid_blocks_ = 100;
  std::pair<float, float> occ_offset_{0.f, 0.f};
  std::vector<float> grid_;  // Eigen::MatrixXf, column-major: (row, col) -> row + col*grid_blocks_

  explicit OccGrid(const Params& p) : size_(p.occ_size), discrete_(p.occ_discrete), dilation_(p.occ_dilation) {
    grid_blocks_ = size_ / discrete_;  // occupancy_grid.cpp:9  int = int/float (truncation)
    grid_.assign((size_t)grid_blocks_ * grid_blocks_, 0.f);
  }
  // occupancy_grid.cpp:27-33 — returns (col, row)
  std::pair<int, int> WorldToOccupancy(float x, float y) const {
    int occ_col = (x - occ_offset_.first) / discrete_ + grid_blocks_ / 2;
    int occ_row = (y - occ_offset_.second) / discrete_ + grid_blocks_ / 2;
    return std::pair<int, int>(occ_col, occ_row);
  }
  // occupancy_grid.cpp:90-101
  bool InGrid(int col, int row) const {
    return !(col >= grid_blocks_ || col < 0 || row >= grid_blocks_ || row < 0);
  }
  // occupancy_grid.cpp:165-168 — grid_(x_ind, y_ind) with float indices converted to Eigen::Index
  bool IsOccupied(float x_ind, float y_ind) const {
    return grid_[(size_t)(long)x_ind + (size_t)(long)y_ind * grid_blocks_] != 0.f;
  }
  // occupancy_grid.cpp:55-88
  void FillOccGrid(const Pose& pose, const Scan& scan) {
    std::fill(grid_.begin(), grid_.end(), 0.f);
    float current_angle = std::atan2(2 * pose.qw * pose.qz, 1 - 2 * pose.qz * pose.qz);  // :60
    occ_offset_.first = pose.px + 0.275 * std::cos(current_angle);   // :63 (cosf; double arithmetic; narrowed)
    occ_offset_.second = pose.py + 0.275 * std::sin(current_angle);  // :64
    int num_scans = (scan.angle_max - scan.angle_min) / scan.angle_increment + 1;  // :66 float expr
    if (num_scans > scan.n_ranges) num_scans = scan.n_ranges;  // oracle guard (reference would read OOB)
    for (int ii = 0; ii < num_scans; ++ii) {
      float angle = scan.angle_min + ii * scan.angle_increment + current_angle;  // :71
      float cx = scan.ranges[ii] * std::cos(angle);  // :50
      float cy = scan.ranges[ii] * std::sin(angle);  // :51
      cx += occ_offset_.first;   // :73
      cy += occ_offset_.second;  // :74
      for (float x_off = -dilation_; x_off <= dilation_; x_off += discrete_) {    // :76
        for (float y_off = -dilation_; y_off <= dilation_; y_off += discrete_) {  // :78
          std::pair<int, int> gp = WorldToOccupancy(cx + x_off, cy + y_off);
          if (InGrid(gp.first, gp.second)) grid_[(size_t)gp.second + (size_t)gp.first * grid_blocks_] = 1.f;  // :83 grid_(row, col)
        }
      }
    }
  }
};

// ---- collision check + selection (src/project.cpp:73-157) ------------------------------------
// Bit-exact contract: inputs = grid, occ_offset, discrete, grid_blocks, R (4 doubles), pose xy,
// table points narrowed to float at the call (project.cpp:86).
struct CheckResult {
  std::vector<uint8_t> valid;       // per path
  std::vector<int> free_count;      // per path
  std::vector<float> end_world;     // per path (x, y) — only meaningful where valid
};
inline CheckResult collision_check(const OccGrid& g, const double R[4], double pose_x, double pose_y,
                                   const double* table_xy, int P, int S) {
  CheckResult r;
  r.valid.assign(P, 0); r.free_count.assign(P, 0); r.end_world.assign(2 * (size_t)P, 0.f);
  for (int i = 0; i < P; ++i) {                                     // project.cpp:76
    int free_points_count = 0;
    for (int j = 0; j < S; ++j) {                                   // :79
      double bx = table_xy[((size_t)i * S + j) * 2 + 0];
      double by = table_xy[((size_t)i * S + j) * 2 + 1];
      std::pair<float, float> wp = car_point_to_world((float)bx, (float)by, R, pose_x, pose_y);  // :86
      std::pair<int, int> occ = g.WorldToOccupancy(wp.first, wp.second);                         // :87
      if (g.InGrid(occ.second, occ.first)) {                        // :89 (swapped args, square grid)
        if (g.IsOccupied((float)occ.second, (float)occ.first)) {    // :92 -> grid_(row, col)
        } else {
          free_points_count++;                                      // :98
        }
      }
    }
    r.free_count[i] = free_points_count;
    if (free_points_count == S) {                                   // :103
      r.valid[i] = 1;
      double ex = table_xy[((size_t)i * S + (S - 1)) * 2 + 0], ey = table_xy[((size_t)i * S + (S - 1)) * 2 + 1];
      std::pair<float, float> wp = car_point_to_world((float)ex, (float)ey, R, pose_x, pose_y);  // :108
      r.end_world[2 * i] = wp.first; r.end_world[2 * i + 1] = wp.second;
    }
  }
  return r;
}
// best-path argmin, project.cpp:125-136: over valid paths in index order, strict <, first wins.
// Returns the table index of the chosen path, or -1 when none is valid (:115-119).
inline int select_best_path(const CheckResult& r, double gx, double gy) {
  double min_dist = std::numeric_limits<double>::max();
  int best = -1;
  for (int i = 0; i < (int)r.valid.size(); ++i) {
    if (!r.valid[i]) continue;
    double px = r.end_world[2 * i], py = r.end_world[2 * i + 1];  // geometry_msgs::Point: float -> double
    double dist = std::pow(std::pow(px - gx, 2) + std::pow(py - gy, 2), 0.5);
    if (dist < min_dist) { min_dist = dist; best = i; }
  }
  return best;
}

// ---- Constraints::FindHalfSpaces (src/constraints.cpp:116-265) -------------------------------
struct HalfSpaces { double l1[3], l2[3]; int best_lo, best_hi; float p1[2], p2[2], p[2]; };
inline bool find_half_spaces(const Params& prm, const State3& state, const Scan& scan, HalfSpaces* out) {
  int num_scans = (scan.angle_max - scan.angle_min) / scan.angle_increment + 1;  // :118
  if (num_scans > scan.n_ranges) num_scans = scan.n_ranges;  // oracle guard
  int max_gap = -1, best_lo = 0, best_hi = 0, lo = -1, hi = -1;
  double poseX = state.x, poseY = state.y;
  float current_angle = state.ori;  // :127
  bool in_gap = 0;
  const float divider_ = prm.fov_divider, ftg_thresh_ = prm.follow_gap_thresh, buffer_ = prm.buffer;
  for (int ii = 0; ii < num_scans; ii++) {
    float angle = scan.angle_min + ii * scan.angle_increment;  // :133
    if (angle > -1.571f / divider_ && angle < 1.571f / divider_) {  // :135
      if (scan.ranges[ii] > ftg_thresh_) {  // :138
        if (in_gap) { hi = ii; } else { lo = ii; in_gap = 1; }   // hi is NOT reset (a13')
      } else {
        in_gap = 0;
        if (hi - lo > max_gap) { max_gap = hi - lo; best_hi = hi; best_lo = lo; }
      }
      if (hi - lo > max_gap) { max_gap = hi - lo; best_hi = hi; best_lo = lo; }
    }
  }
  if (best_hi - best_lo > 2 * buffer_) {  // :173
    best_hi = best_hi - buffer_;
    best_lo = best_lo + buffer_;
  }
  out->best_lo = best_lo; out->best_hi = best_hi;
  if (best_lo < 0 || best_hi < 0 || best_lo >= scan.n_ranges || best_hi >= scan.n_ranges) return false;  // UB in the reference
  float angle1 = scan.angle_min + best_lo * scan.angle_increment + current_angle;  // :179
  float angle2 = scan.angle_min + best_hi * scan.angle_increment + current_angle;  // :180
  std::pair<float, float> p1_, p2_, p_;
  p1_.first = scan.ranges[best_lo] * std::cos(angle1) + poseX;   // :182 float*cosf -> float, + double, narrowed
  p1_.second = scan.ranges[best_lo] * std::sin(angle1) + poseY;  // :183
  p2_.first = scan.ranges[best_hi] * std::cos(angle2) + poseX;   // :185
  p2_.second = scan.ranges[best_hi] * std::sin(angle2) + poseY;  // :186
  p_.first = poseX;   // :188
  p_.second = poseY;  // :189
  float a1, b1, c1, a2, b2, c2;
  a1 = p_.second - p1_.second;                            // :233
  b1 = p1_.first - p_.first;
  c1 = p_.first * p1_.second - p_.second * p1_.first;
  if (a1 * p2_.first + b1 * p2_.second + c1 < 0) { a1 = -a1; b1 = -b1; c1 = -c1; }  // :237
  a2 = p_.second - p2_.second;                            // :244
  b2 = p2_.first - p_.first;
  c2 = p_.first * p2_.second - p_.second * p2_.first;
  if (a2 * p1_.first + b2 * p1_.second + c2 < 0) { a2 = -a2; b2 = -b2; c2 = -c2; }  // :248
  out->l1[0] = a1; out->l1[1] = b1; out->l1[2] = c1 + 0.5;  // :258-260
  out->l2[0] = a2; out->l2[1] = b2; out->l2[2] = c2 + 0.5;  // :262-264
  out->p1[0] = p1_.first; out->p1[1] = p1_.second; out->p2[0] = p2_.first; out->p2[1] = p2_.second;
  out->p[0] = p_.first; out->p[1] = p_.second;
  return true;
}

// ---- MPC QP assembly (src/mpc.cpp:26-35, 208-306) -----------------------------------------------
// Per-QP parameter record (doubles): x0[3] | u_lin[2] = (v, steer) | l1[3] | l2[3] | ref[3*N]
inline int qp_record_doubles(int N) { return 11 + 3 * N; }

struct MpcConfig {
  int N = 30;
  double dt = (double)0.01f;        // float dt_ widened at the Linearize call (mpc.cpp:73, mpc.h:48)
  double Q[3] = {10.0, 10.0, 0.0};  // diag, mpc.cpp:20-24
  double R[2] = {0.10, 5.0};
  double u_des[2] = {4.5, 0.0};     // mpc.cpp:18-19
  double u_min[2] = {(double)3.0f, (double)-0.43f};  // constraints.cpp:18-21 (floats into VectorXd)
  double u_max[2] = {(double)4.5f, (double)0.43f};
  int gap_mode = 0;  // 0: as shipped, gap bounds (-INFTY, +INFTY) (mpc.cpp:297-298); 1: lower = -l(2) (the commented code);
                     // 2: like 1 but the stage-0 pair (all-ones rows, not half-planes — SURVEY fact 3) stays loose
  // Steering-rate rows.  NOT in the reference (its only relic of an extra input constraint is the
  // commented slip block, constraints.cpp:23-39, mpc.cpp:250): N rows appended after the input box,
  //   row k:  delta_k - delta_{k-1}  in [-rate_delta, +rate_delta]   (k >= 1)
  //   row 0:  delta_0                in [steer_prev - rate_delta, steer_prev + rate_delta],  steer_prev = u_lin[1]
  int rate_rows = 0;
  double rate_delta = 0.0;   // max steering change per step (rad)
  // State-box rows.  The reference STORES x_min_ / x_max_ (constraints.cpp:14-17) and has
  // Constraints::SetXLims (constraints.cpp:108-114: x and y within +-d of the current state, d = /state_lims; the orientation
  // stays at +-INFTY) but never stacks them into the QP.  Stacked here the way the input box is: 3(N+1) identity rows on
  // x_0..x_N appended after the input box (and after the steering-rate rows when both are on),
  //   rows 3k, 3k+1:  x_k, y_k  in [x_cur - d, x_cur + d], [y_cur - d, y_cur + d]      row 3k+2:  ori_k in (-INFTY, +INFTY)
  int state_rows = 0;
  double state_lim = 0.0;    // d
};

struct QpData {
  int N = 0, n = 0, m = 0;
  osqp_restated::Csc P, A;           // P upper-triangular incl. the explicit zeros of the dense Q/R blocks
  std::vector<double> q, l, u;
  std::vector<int> perm;             // stage-interleaved KKT ordering
  // positions of the per-cycle coefficients inside A.x
  std::vector<int> posA, posB, posG; // [N][9], [N][6], [N][6] (k = 1..N)
};

// Structure as built by MPC::MPC / CreateHessianMatrix / CreateLinearConstraintMatrix /
// Create{Lower,Upper}Bound (mpc.cpp:26-47, 208-219, 231-254, 275-291).
inline void qp_build_structure(const MpcConfig& cfg, QpData* d) {
  const int N = cfg.N, ns = 3 * (N + 1), nu = 2 * N;
  d->N = N; d->n = ns + nu; d->m = ns + 2 * (N + 1) + nu + (cfg.rate_rows ? N : 0) + (cfg.state_rows ? ns : 0);
  const int n = d->n, m = d->m;
  struct T { int r, c; double v; int tag; };
  // Hessian: dense diagonal blocks (upper triangle kept, OsqpEigen passes triu to OSQP)
  std::vector<T> tp;
  for (int k = 0; k <= N; ++k)
    for (int r = 0; r < 3; ++r) for (int c = r; c < 3; ++c) tp.push_back({3 * k + r, 3 * k + c, r == c ? cfg.Q[r] : 0.0, 0});
  for (int k = 0; k < N; ++k)
    for (int r = 0; r < 2; ++r) for (int c = r; c < 2; ++c) tp.push_back({ns + 2 * k + r, ns + 2 * k + c, r == c ? cfg.R[r] : 0.0, 0});
  auto to_csc = [](std::vector<T>& t, int nrow, int ncol, osqp_restated::Csc* M, std::vector<int>* where) {
    std::vector<int> order(t.size());
    for (size_t k = 0; k < t.size(); ++k) order[k] = (int)k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      return t[a].c != t[b].c ? t[a].c < t[b].c : t[a].r < t[b].r;
    });
    M->nrow = nrow; M->ncol = ncol; M->p.assign(ncol + 1, 0); M->i.clear(); M->x.clear();
    if (where) where->assign(t.size(), -1);
    for (int idx : order) {
      M->i.push_back(t[idx].r); M->x.push_back(t[idx].v); M->p[t[idx].c + 1]++;
      if (where) (*where)[idx] = (int)M->x.size() - 1;
    }
    for (int c = 0; c < ncol; ++c) M->p[c + 1] += M->p[c];
  };
  to_csc(tp, n, n, &d->P, nullptr);
  // Linear constraint matrix
  std::vector<T> ta;
  for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) ta.push_back({ns + r, c, 1.0, 0});  // :241 gap_con ones at stage 0
  for (int r = 0; r < ns; ++r) ta.push_back({r, r, -1.0, 0});                                 // :244
  for (int k = 1; k <= N; ++k) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) ta.push_back({3 * k + r, 3 * (k - 1) + c, r == c ? 1.0 : 0.0, 1000 + (k - 1) * 9 + r * 3 + c});          // :247
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 2; ++c) ta.push_back({3 * k + r, ns + 2 * (k - 1) + c, r == c ? 1.0 : 0.0, 100000 + (k - 1) * 6 + r * 2 + c});   // :248
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) ta.push_back({ns + 2 * k + r, 3 * k + c, 1.0, 200000 + (k - 1) * 6 + r * 3 + c});                        // :249
  }
  for (int r = 0; r < nu; ++r) ta.push_back({ns + 2 * (N + 1) + r, ns + r, 1.0, 0});           // :253
  const int rate0 = ns + 2 * (N + 1) + nu;   // first steering-rate row
  if (cfg.rate_rows)
    for (int k = 0; k < N; ++k) {
      ta.push_back({rate0 + k, ns + 2 * k + 1, 1.0, 0});
      if (k > 0) ta.push_back({rate0 + k, ns + 2 * (k - 1) + 1, -1.0, 0});
    }
  const int st0 = rate0 + (cfg.rate_rows ? N : 0);   // first state-box row
  if (cfg.state_rows)
    for (int r = 0; r < ns; ++r) ta.push_back({st0 + r, r, 1.0, 0});
  std::vector<int> where;
  to_csc(ta, m, n, &d->A, &where);
  d->posA.assign(9 * N, -1); d->posB.assign(6 * N, -1); d->posG.assign(6 * N, -1);
  for (size_t k = 0; k < ta.size(); ++k) {
    int tag = ta[k].tag;
    if (tag >= 200000) d->posG[tag - 200000] = where[k];
    else if (tag >= 100000) d->posB[tag - 100000] = where[k];
    else if (tag >= 1000) d->posA[tag - 1000] = where[k];
  }
  d->q.assign(n, 0.0);
  d->l.assign(m, 0.0); d->u.assign(m, 0.0);
  for (int k = 0; k <= N; ++k) for (int r = 0; r < 2; ++r) {
    d->l[ns + 2 * k + r] = -osqp_restated::OSQP_INFTY;  // :279-281
    d->u[ns + 2 * k + r] = osqp_restated::OSQP_INFTY;   // :288-290
  }
  for (int k = 0; k < N; ++k) for (int r = 0; r < 2; ++r) {
    d->l[ns + 2 * (N + 1) + 2 * k + r] = cfg.u_min[r];
    d->u[ns + 2 * (N + 1) + 2 * k + r] = cfg.u_max[r];
  }
  if (cfg.rate_rows)
    for (int k = 0; k < N; ++k) { d->l[rate0 + k] = -cfg.rate_delta; d->u[rate0 + k] = cfg.rate_delta; }
  if (cfg.state_rows)
    for (int r = 0; r < ns; ++r) { d->l[st0 + r] = -osqp_restated::OSQP_INFTY; d->u[st0 + r] = osqp_restated::OSQP_INFTY; }   // constraints.cpp:14-17
  // KKT ordering, stage by stage: [dyn rows k | x_k | gap rows k | u_k | box rows k | rate row k]
  d->perm.clear();
  for (int k = 0; k <= N; ++k) {
    for (int r = 0; r < 3; ++r) d->perm.push_back(n + 3 * k + r);
    for (int r = 0; r < 3; ++r) d->perm.push_back(3 * k + r);
    for (int r = 0; r < 2; ++r) d->perm.push_back(n + ns + 2 * k + r);
    if (cfg.state_rows) for (int r = 0; r < 3; ++r) d->perm.push_back(n + st0 + 3 * k + r);
    if (k < N) {
      for (int r = 0; r < 2; ++r) d->perm.push_back(ns + 2 * k + r);
      for (int r = 0; r < 2; ++r) d->perm.push_back(n + ns + 2 * (N + 1) + 2 * k + r);
      if (cfg.rate_rows) d->perm.push_back(n + rate0 + k);
    }
  }
}

// Per-cycle values: CreateGradientVector (mpc.cpp:221-229), UpdateLinearConstraintMatrix
// (:256-273), Update{Lower,Upper}Bound (:293-306), with Model::Linearize (:73).
inline void qp_fill_values(const MpcConfig& cfg, const double* rec, QpData* d) {
  const int N = cfg.N, ns = 3 * (N + 1);
  const double* x0 = rec; const double* ulin = rec + 3; const double* l1 = rec + 5; const double* l2 = rec + 8;
  const double* ref = rec + 11;
  double A[9], B[6], C[3];
  linearize(x0[2], ulin[0], ulin[1], cfg.dt, A, B, C);
  for (int k = 0; k < N; ++k) {
    for (int r = 0; r < 3; ++r) d->q[3 * k + r] = -1 * cfg.Q[r] * ref[3 * k + r];  // :225 (Q diagonal: -1*Q*ref)
    for (int r = 0; r < 2; ++r) d->q[ns + 2 * k + r] = -1 * cfg.R[r] * cfg.u_des[r];  // :226
  }
  for (int r = 0; r < 3; ++r) d->q[3 * N + r] = -1 * cfg.Q[r] * ref[3 * (N - 1) + r];  // :228
  double G[6] = {l1[0], l1[1], 0.0, l2[0], l2[1], 0.0};  // :260-266
  for (int k = 0; k < N; ++k) {
    for (int e = 0; e < 9; ++e) d->A.x[d->posA[9 * k + e]] = A[e];  // :269
    for (int e = 0; e < 6; ++e) d->A.x[d->posB[6 * k + e]] = B[e];  // :270
    for (int e = 0; e < 6; ++e) d->A.x[d->posG[6 * k + e]] = G[e];  // :271
  }
  for (int r = 0; r < 3; ++r) d->l[r] = d->u[r] = -x0[r];           // :299, :305
  for (int k = 1; k <= N; ++k) for (int r = 0; r < 3; ++r) d->l[3 * k + r] = d->u[3 * k + r] = -C[r];
  for (int k = 0; k <= N; ++k) {
    const bool on = cfg.gap_mode == 1 || (cfg.gap_mode == 2 && k > 0);
    d->l[ns + 2 * k + 0] = on ? -l1[2] : -osqp_restated::OSQP_INFTY;  // :297
    d->l[ns + 2 * k + 1] = on ? -l2[2] : -osqp_restated::OSQP_INFTY;  // :298
  }
  if (cfg.rate_rows) {  // row 0 is measured from the steering applied last cycle
    const int rate0 = ns + 2 * (N + 1) + 2 * N;
    d->l[rate0] = ulin[1] - cfg.rate_delta;
    d->u[rate0] = ulin[1] + cfg.rate_delta;
  }
  if (cfg.state_rows) {  // Constraints::SetXLims(current_state), constraints.cpp:108-114
    const int st0 = ns + 2 * (N + 1) + 2 * N + (cfg.rate_rows ? N : 0);
    for (int k = 0; k <= N; ++k)
      for (int r = 0; r < 2; ++r) { d->l[st0 + 3 * k + r] = x0[r] - cfg.state_lim; d->u[st0 + 3 * k + r] = x0[r] + cfg.state_lim; }
  }
}

}  // namespace f110_ref
