// Plain-data stand-ins for the ROS message types the reference classes take (field widths as in ROS:
// LaserScan fields are float32, Pose fields are float64), a tiny dense Vector/Matrix pair for the Eigen
// accessors, and the parameter set of params.yaml with the reference's C++ destination types.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
}  // namespace geometry_msgs

namespace sensor_msgs {
struct LaserScan {
  float angle_min = 0, angle_max = 0, angle_increment = 0;
  float range_min = 0, range_max = 0;
  std::vector<float> ranges;
};
}  // namespace sensor_msgs

namespace f110 {

class Vector {
 public:
  Vector() = default;
  explicit Vector(std::size_t n, double fill = 0.0) : d_(n, fill) {}
  Vector(std::initializer_list<double> il) : d_(il) {}
  std::size_t size() const { return d_.size(); }
  void resize(std::size_t n) { d_.assign(n, 0.0); }
  double& operator()(std::size_t i) { return d_[i]; }
  double operator()(std::size_t i) const { return d_[i]; }
  const double* data() const { return d_.data(); }
 private:
  std::vector<double> d_;
};

class Matrix {  // row-major
 public:
  Matrix() = default;
  Matrix(std::size_t r, std::size_t c, double fill = 0.0) : r_(r), c_(c), d_(r * c, fill) {}
  static Matrix Diagonal(std::initializer_list<double> diag) {
    Matrix m(diag.size(), diag.size());
    std::size_t i = 0;
    for (double v : diag) { m(i, i) = v; ++i; }
    return m;
  }
  std::size_t rows() const { return r_; }
  std::size_t cols() const { return c_; }
  double& operator()(std::size_t r, std::size_t c) { return d_[r * c_ + c]; }
  double operator()(std::size_t r, std::size_t c) const { return d_[r * c_ + c]; }
  const double* data() const { return d_.data(); }
 private:
  std::size_t r_ = 0, c_ = 0;
  std::vector<double> d_;
};

// params.yaml (reference) -> typed fields; the comment names the reader in the reference.
struct Params {
  double q0 = 10.0, q1 = 10.0, q2 = 0.0;   // mpc.cpp:12-14
  double r0 = 0.10, r1 = 5.0;              // mpc.cpp:15-16
  int horizon = 30;                        // mpc.cpp:5
  float dt = 0.01f;                        // mpc.cpp:6 (float member, mpc.h:48)
  double dt_planner = 0.01;                // trajectory_planner.cpp:10 (double member)
  int occ_size = 10;                       // occupancy_grid.cpp:6
  float occ_discrete = 0.1f;               // occupancy_grid.cpp:7
  float occ_dilation = 0.15f;              // occupancy_grid.cpp:8
  double des_vel = 4.5, des_steer = 0.0;   // mpc.cpp:10-11
  float umax = 4.5f, umin = 3.0f;          // constraints.cpp:7-8
  double speed_max = 4.5;                  // trajectory_planner.cpp:5 (reads "umax" as double)
  float follow_gap_thresh = 3.0f;          // constraints.cpp:9
  float state_lims = 1.0f;                 // constraints.cpp:10
  float fov_divider = 1.5f;                // constraints.cpp:11
  float buffer = 3.0f;                     // constraints.cpp:12
  int speed_discrete = 40;                 // trajectory_planner.cpp:7 (unused there as well)
  int steer_discrete = 30;                 // trajectory_planner.cpp:8
  double steer_max = 0.4;                  // trajectory_planner.cpp:6
  int traj_discrete = 50;                  // trajectory_planner.cpp:9
  float lookahead = 2.5f;                  // trajectory.cpp:10
  // gap rows: 0 = as shipped (bounds +-INFTY, mpc.cpp:297-298), 1 = lower bound -l(2) restored
  int gap_mode = 0;
  // steering-rate rows (not in the reference): |delta_k - delta_{k-1}| <= steer_rate_max * dt, 0 = off
  double steer_rate_max = 0.0;             // rad/s
  // state box (stored by the reference, never stacked — constraints.cpp:14-17, 108-114): true stacks x_k, y_k within +-state_lims of
  // the current state (Constraints::SetXLims) as 3(N+1) extra rows; horizon <= 31, not together with the steering-rate rows
  bool state_box = false;

  // "key: value" lines of a params.yaml-style file override the defaults; unknown keys are ignored.
  static Params FromYaml(const std::string& path);
};

}  // namespace f110
