"""Synthetic workloads for the batched MPC solve (numpy only — no oracle, no GPU).

Definitions follow SURVEY.md §8(d) / BASELINE.md §2.  The per-QP parameter record is
    x0[3] | u_lin[2] = (v, steer) | l1[3] | l2[3] | ref[3*N]          (11 + 3N doubles)
i.e. exactly what MPC::Update receives per cycle (reference src/mpc.cpp:69-80): the current state,
the linearisation input, the two half-plane lines from Constraints::FindHalfSpaces and the
desired state trajectory (first N states of the chosen mini-path, world frame, ori = 0.0 as in
src/project.cpp:145-149).
"""
import numpy as np

DT_F32 = float(np.float32(0.01))       # float dt_ widened to double (mpc.h:48, mpc.cpp:73)
CAR_LENGTH = 0.35                      # model.cpp:2


def record_doubles(N):
    return 11 + 3 * N


def traj_table(steer_max=0.4, steer_discrete=30, traj_discrete=50, speed_max=4.5, dt=0.01):
    """Traj_Plan::generate_traj_table (trajectory_planner.cpp:26-72): (steer_discrete+1, traj_discrete, 3)."""
    P = steer_discrete + 1
    out = np.zeros((P, traj_discrete, 3))
    ds = 2 * steer_max / steer_discrete
    for i in range(P):
        steer = -steer_max + i * ds
        s = np.zeros(3)
        for k in range(traj_discrete - 1):
            d = np.array([speed_max * np.cos(s[2]), speed_max * np.sin(s[2]), np.tan(steer) * speed_max / CAR_LENGTH])
            s = s + d * dt
            out[i, k + 1] = s
    return out


def yaw_pose(x, y, yaw):
    """geometry_msgs::Pose (px,py,pz,qx,qy,qz,qw) for a planar yaw."""
    return np.array([x, y, 0.0, 0.0, 0.0, np.sin(yaw / 2.0), np.cos(yaw / 2.0)])


def path_to_world(path_xy, x, y, yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    wx = c * path_xy[:, 0] - s * path_xy[:, 1] + x
    wy = s * path_xy[:, 0] + c * path_xy[:, 1] + y
    return np.stack([wx, wy], axis=1)


def tracking_batch(B, N=30, seed=20240905, gaps=False, table=None):
    """Config-5 style batch: random poses, a random mini-path as reference, random start error and
    previous steering.  gaps=True also fills plausible half-plane lines (config 3 style)."""
    rng = np.random.default_rng(seed)
    if table is None:
        table = traj_table(traj_discrete=max(50, N))
    P = table.shape[0]
    recs = np.zeros((B, record_doubles(N)))
    for b in range(B):
        x, y = rng.uniform(-5, 5, 2)
        yaw = rng.uniform(-np.pi, np.pi)
        pi = rng.integers(0, P)
        ref = path_to_world(table[pi, :, :2], x, y, yaw)[:N]
        lat = rng.uniform(-0.3, 0.3)
        dyaw = rng.uniform(-0.2, 0.2)
        x0 = np.array([x - np.sin(yaw) * lat, y + np.cos(yaw) * lat, yaw + dyaw])
        recs[b, 0:3] = x0
        recs[b, 3] = 4.5                               # project.cpp:170
        recs[b, 4] = rng.uniform(-0.4, 0.4)
        if gaps:
            # two lines through the car position opening +-(0.5..1.0) rad around the heading
            a1 = yaw + rng.uniform(0.5, 1.0)
            a2 = yaw - rng.uniform(0.5, 1.0)
            for k, a in enumerate((a1, a2)):
                p1 = np.array([x0[0] + 5 * np.cos(a), x0[1] + 5 * np.sin(a)])
                p2 = np.array([x0[0] + 5 * np.cos(2 * yaw - a), x0[1] + 5 * np.sin(2 * yaw - a)])
                aa = x0[1] - p1[1]
                bb = p1[0] - x0[0]
                cc = x0[0] * p1[1] - x0[1] * p1[0]
                if aa * p2[0] + bb * p2[1] + cc < 0:
                    aa, bb, cc = -aa, -bb, -cc
                recs[b, 5 + 3 * k: 8 + 3 * k] = (np.float32(aa), np.float32(bb), np.float32(cc) + 0.5)
        else:
            recs[b, 5:8] = (0.3, -0.8, 1.5)
            recs[b, 8:11] = (-0.4, 0.7, 2.0)
        r = np.zeros((N, 3))
        r[:, :2] = ref
        recs[b, 11:] = r.reshape(-1)
    return recs
