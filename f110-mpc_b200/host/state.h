// State / Input value types — same public methods as the reference's (include/f110-mpc/state.h:10-45,
// include/f110-mpc/input.h:11-34), Eigen replaced by f110::Vector.
#pragma once
#include <utility>
#include "msgs.h"

class State {
 public:
  State() = default;
  State(double x, double y, double ori) : x_(x), y_(y), ori_(ori) {}
  virtual ~State() = default;
  f110::Vector StateToVector() const { return f110::Vector{x_, y_, ori_}; }
  void set_x(double x) { x_ = x; }
  void set_y(double y) { y_ = y; }
  void set_ori(double ori) { ori_ = ori; }
  std::pair<float, float> GetPair() const { return {static_cast<float>(x_), static_cast<float>(y_)}; }  // state.cpp:43-46 narrows
  double x() const { return x_; }
  double y() const { return y_; }
  double ori() const { return ori_; }
  int size() const { return 3; }
 private:
  double x_ = 0, y_ = 0, ori_ = 0;
};

class Input {
 public:
  Input() = default;
  Input(double v, double steer_ang) : v_(v), steer_ang_(steer_ang) {}
  virtual ~Input() = default;
  f110::Vector InputToVector() const { return f110::Vector{v_, steer_ang_}; }
  void set_v(double v) { v_ = v; }
  void set_steer_ang(double a) { steer_ang_ = a; }
  double v() const { return v_; }
  double steer_ang() const { return steer_ang_; }
 private:
  double v_ = 0, steer_ang_ = 0;
};
