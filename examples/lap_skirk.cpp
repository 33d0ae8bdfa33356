// Example: the ROS-free `project` orchestrator lapping a raceline against the roll-out model, one car, GPU solves.
//   ./lap_skirk <params.yaml> <raceline.csv> [ticks]
// Every tick (10 ms): OdomCallback(pose); every 2nd tick the DriveLoop body publishes the next solved input; a constant
// free-space scan arrives every 4th tick.  Prints the pose every 50 ticks and a summary.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "../f110-mpc_b200/host/project.h"

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s params.yaml raceline.csv [ticks]\n", argv[0]); return 2; }
  const f110::Params prm = f110::Params::FromYaml(argv[1]);
  const int ticks = argc > 3 ? std::atoi(argv[3]) : 1000;
  project node(prm, /*device=*/0);
  if (!node.LoadRaceline(argv[2])) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
  sensor_msgs::LaserScan scan;
  scan.angle_min = -2.35f; scan.angle_increment = 4.7f / 1079; scan.angle_max = scan.angle_min + 1079 * scan.angle_increment;
  scan.ranges.assign(1080, 10.0f);
  Model plant;
  State car(0.0, 0.0, 0.0);
  Input applied(0.5, 0.0);
  for (int t = 0; t < ticks; ++t) {
    geometry_msgs::Pose pose;
    pose.position.x = car.x(); pose.position.y = car.y();
    pose.orientation.z = std::sin(car.ori() / 2.0); pose.orientation.w = std::cos(car.ori() / 2.0);
    node.OdomCallback(pose);
    if (t % 4 == 0) node.ScanCallback(scan);
    Input in;
    if (t % 2 == 0 && node.DriveStep(&in)) applied = in;
    if (t % 50 == 0) std::printf("t=%5.2fs  x=%7.3f y=%7.3f yaw=%6.3f  v=%4.2f steer=%6.3f\n", t * 0.01, car.x(), car.y(), car.ori(), applied.v(), applied.steer_ang());
    State next;
    plant.simulate_dynamics(car, applied, 0.01, next);
    car = next;
  }
  std::printf("%d planning cycles, %d MPC solves\n", node.cycles_planned(), node.cycles_solved());
  return 0;
}
