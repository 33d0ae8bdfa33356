#include <cstdlib>
#include <fstream>
#include <sstream>
#include "msgs.h"

namespace f110 {
Params Params::FromYaml(const std::string& path) {
  Params p;
  std::ifstream in(path);
  std::string line;
  while (std::getline(in, line)) {
    const std::size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    const std::size_t colon = line.find(':');
    if (colon == std::string::npos) continue;
    std::string key = line.substr(0, colon), val = line.substr(colon + 1);
    auto trim = [](std::string& s) {
      const char* ws = " \t\r\n\"";
      s.erase(0, s.find_first_not_of(ws));
      s.erase(s.find_last_not_of(ws) + 1);
    };
    trim(key); trim(val);
    if (key.empty() || val.empty()) continue;
    const double d = std::atof(val.c_str());
    if (key == "q0") p.q0 = d; else if (key == "q1") p.q1 = d; else if (key == "q2") p.q2 = d;
    else if (key == "r0") p.r0 = d; else if (key == "r1") p.r1 = d;
    else if (key == "horizon") p.horizon = static_cast<int>(d);
    else if (key == "dt") { p.dt = static_cast<float>(d); p.dt_planner = d; }
    else if (key == "occ_size") p.occ_size = static_cast<int>(d);
    else if (key == "occ_discrete") p.occ_discrete = static_cast<float>(d);
    else if (key == "occ_dilation") p.occ_dilation = static_cast<float>(d);
    else if (key == "des_vel") p.des_vel = d; else if (key == "des_steer") p.des_steer = d;
    else if (key == "umax") { p.umax = static_cast<float>(d); p.speed_max = d; }
    else if (key == "umin") p.umin = static_cast<float>(d);
    else if (key == "follow_gap_thresh") p.follow_gap_thresh = static_cast<float>(d);
    else if (key == "state_lims") p.state_lims = static_cast<float>(d);
    else if (key == "fov_divider") p.fov_divider = static_cast<float>(d);
    else if (key == "buffer") p.buffer = static_cast<float>(d);
    else if (key == "speed_discrete") p.speed_discrete = static_cast<int>(d);
    else if (key == "steer_discrete") p.steer_discrete = static_cast<int>(d);
    else if (key == "steer_max") p.steer_max = d;
    else if (key == "traj_discrete") p.traj_discrete = static_cast<int>(d);
    else if (key == "lookahead") p.lookahead = static_cast<float>(d);
    else if (key == "gap_mode") p.gap_mode = static_cast<int>(d);
    else if (key == "steer_rate_max") p.steer_rate_max = d;
    else if (key == "state_box") p.state_box = d != 0.0;
  }
  return p;
}
}  // namespace f110
