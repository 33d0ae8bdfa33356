// C++ check of the multi-GPU batch interface (BatchMPC over several devices -> f110_mpc_create_multi /
// f110_mpc_solve_multi_host), run on the GPU box by tests/test_host_gpu.py.  The same problems are solved by a one-device
// BatchMPC and by a BatchMPC over every visible device, sharded in whole scenarios of 140 QPs (SURVEY.md section 8e, config 4:
// 7 lanes x 20 paths per scenario); every first control, status and iteration count must be identical.
#include <cmath>
#include <cstdio>
#include <vector>

#include "mpc.h"

static int failures = 0;
#define CHECK(cond)                                                                          \
  do {                                                                                       \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

int main() {
  const int ndev = f110_device_count();
  if (ndev < 1) { std::printf("no CUDA device\n"); return 2; }
  f110::Params prm;
  const int unit = 140, scenarios = 5, B = unit * scenarios;
  std::vector<int> devices;
  for (int d = 0; d < ndev; ++d) devices.push_back(d);
  BatchMPC one(prm, B, 0, false);
  BatchMPC many(prm, B, devices, unit);
  CHECK(many.num_devices() == ndev && one.num_devices() == 1);
  for (int b = 0; b < B; ++b) {
    // a car near the origin heading along x, tracking a gently curving path; every QP differs in start offset and curvature
    const double lat = 0.002 * (b % 97) - 0.1, yaw0 = 0.001 * (b % 53) - 0.02, curv = 0.0004 * (b % 41);
    State x0(0.01 * (b % 7), lat, yaw0);
    Input in(4.5, 0.01 * ((b % 9) - 4));
    std::vector<State> ref;
    for (int k = 0; k < prm.horizon; ++k) {
      const double sx = 0.045 * (k + 1);
      ref.emplace_back(sx, curv * sx * sx * 10.0, 0.0);
    }
    f110::Vector l1(3), l2(3);
    l1(0) = 0.3; l1(1) = -0.8; l1(2) = 1.5; l2(0) = -0.4; l2(1) = 0.7; l2(2) = 2.0;
    one.SetProblem(b, x0, in, ref, l1, l2);
    many.SetProblem(b, x0, in, ref, l1, l2);
  }
  CHECK(one.Solve(B) == F110_OK);
  CHECK(many.Solve(B) == F110_OK);
  int solved = 0;
  for (int b = 0; b < B; ++b) {
    CHECK(one.status(b) == many.status(b));
    CHECK(one.iterations(b) == many.iterations(b));
    CHECK(one.first_input(b).v() == many.first_input(b).v());
    CHECK(one.first_input(b).steer_ang() == many.first_input(b).steer_ang());
    solved += one.status(b) == F110_SOLVED;
  }
  CHECK(solved > B / 2);
  // a batch of fewer scenarios than devices, and an argument error
  CHECK(many.Solve(unit) == F110_OK);
  for (int b = 0; b < unit; ++b) CHECK(one.first_input(b).v() == many.first_input(b).v() && one.status(b) == many.status(b));
  CHECK(many.Solve(unit + 1) == F110_ERR_ARG);
  if (failures) { std::printf("%d failures\n", failures); return 1; }
  std::printf("multi-GPU host checks passed on %d device(s), %d of %d QPs solved\n", ndev, solved, B);
  return 0;
}
