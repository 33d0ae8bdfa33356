// Model — forward-Euler linearised kinematic bicycle, x+ = A x + B u + C (reference include/f110-mpc/model.h:16-26).
// The batched solve re-derives A, B, C on the device from (ori, v, steer); this host copy serves the accessors
// and the mini-path roll-out.
#pragma once
#include "state.h"

class Model {
 public:
  Model();
  virtual ~Model() = default;
  f110::Matrix A() const { return A_; }
  f110::Matrix B() const { return B_; }
  f110::Matrix C() const { return C_; }
  void Linearize(State& S, Input& I, double dt);                                    // model.cpp:30-59
  void simulate_dynamics(State& state, Input& input, double dt, State& new_state);  // model.cpp:61-76
 private:
  f110::Matrix A_, B_, C_;
};
