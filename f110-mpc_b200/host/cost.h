// Cost — holds the Q (3x3) and R (2x2) weight matrices (reference include/f110-mpc/cost.h:11-16).
#pragma once
#include "msgs.h"

class Cost {
 public:
  Cost() = default;
  Cost(const f110::Matrix& q, const f110::Matrix& r) : q_(q), r_(r) {}
  virtual ~Cost() = default;
  f110::Matrix q() const { return q_; }
  f110::Matrix r() const { return r_; }
 private:
  f110::Matrix q_, r_;
};
