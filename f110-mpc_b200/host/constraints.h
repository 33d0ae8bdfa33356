// Constraints — input box, (unused) state box and the largest-gap half-planes
// (reference include/f110-mpc/constraints.h:17-41, src/constraints.cpp).
#pragma once
#include <utility>
#include "msgs.h"
#include "state.h"

class Constraints {
 public:
  static constexpr double INFTY = 1e30;  // OsqpEigen::INFTY
  explicit Constraints(const f110::Params& params);
  virtual ~Constraints() = default;

  void set_x_max(const f110::Vector& v) { x_max_ = v; }
  void set_u_max(const f110::Vector& v) { u_max_ = v; }
  void set_x_min(const f110::Vector& v) { x_min_ = v; }
  void set_u_min(const f110::Vector& v) { u_min_ = v; }
  void set_state(State& state) { state_ = state; }
  void SetXLims(State x);  // constraints.cpp:108-114

  f110::Vector x_max() const { return x_max_; }
  f110::Vector u_max() const { return u_max_; }
  f110::Vector x_min() const { return x_min_; }
  f110::Vector u_min() const { return u_min_; }
  f110::Vector l1() const { return l1_; }
  f110::Vector l2() const { return l2_; }

  // constraints.cpp:116-265.  Where the reference indexes ranges[-1] (no gap of >= 2 beams in the field of
  // view) this returns false and leaves l1/l2 untouched; otherwise true.
  bool FindHalfSpaces(State& state, sensor_msgs::LaserScan& scan_msg);
  std::pair<int, int> best_gap() const { return {best_lo_, best_hi_}; }

 private:
  f110::Vector x_max_, u_max_, x_min_, u_min_, l1_, l2_;
  State state_;
  float d_, ftg_thresh_, umax_val_, umin_val_, divider_, buffer_;
  std::pair<float, float> p1_, p2_, p_;
  int best_lo_ = 0, best_hi_ = 0;
};
