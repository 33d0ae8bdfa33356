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


# ---- scenes for the mini-path collision check ---------------------------------
def golden_dir():
    import os
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def reference_data():
    """Fixtures derived from the reference's csv/ files (scripts/make_golden.py)."""
    import os
    return np.load(os.path.join(golden_dir(), "reference_data.npz"))


def skirk_waypoints():
    """Trajectory::ReadCSV (trajectory.cpp:18-55): float x, y and heading atan2f from the previous point."""
    xy = reference_data()["skirk_xy"].astype(np.float32)
    prev = np.roll(xy, 1, axis=0)
    # `temp[(i-1)%temp.size()]` with a 32-bit unsigned i (trajectory.cpp:40-43): at i = 0 the index is
    # (2^32 - 1) % size, not size - 1
    prev[0] = xy[(2 ** 32 - 1) % len(xy)]
    ori = np.arctan2((xy[:, 1] - prev[:, 1]).astype(np.float32), (xy[:, 0] - prev[:, 0]).astype(np.float32)).astype(np.float32)
    return xy, ori


SCAN_BEAMS = 1080
SCAN_ANGLE_MIN = float(np.float32(-2.35))
SCAN_ANGLE_INC = float(np.float32(4.7 / 1079))
SCAN_ANGLE_MAX = float(np.float32(np.float32(-2.35) + np.float32(1079) * np.float32(4.7 / 1079)))


def render_scan(x, y, yaw, boxes, max_range=10.0):
    """Ray-cast axis-aligned boxes (cx, cy, half) into a 1080-beam scan taken from the car pose."""
    ang = SCAN_ANGLE_MIN + np.arange(SCAN_BEAMS) * SCAN_ANGLE_INC + yaw
    dx, dy = np.cos(ang), np.sin(ang)
    r = np.full(SCAN_BEAMS, max_range)
    for (cx, cy, h) in boxes:
        with np.errstate(divide="ignore", invalid="ignore"):
            tx1, tx2 = (cx - h - x) / dx, (cx + h - x) / dx
            ty1, ty2 = (cy - h - y) / dy, (cy + h - y) / dy
        tmin = np.maximum(np.minimum(tx1, tx2), np.minimum(ty1, ty2))
        tmax = np.minimum(np.maximum(tx1, tx2), np.maximum(ty1, ty2))
        hit = (tmax >= tmin) & (tmax > 0) & (tmin > 0)
        r = np.where(hit & (tmin < r), tmin, r)
    return r.astype(np.float32)


def scene_batch(S, seed=20240902):
    """S scenes: pose = a skirk waypoint, 1-3 box obstacles 0.3-0.6 m wide 0.8-2.0 m ahead.
    Returns poses (S,7), yaw (S,), scans (S,1080) float32."""
    rng = np.random.default_rng(seed)
    xy, ori = skirk_waypoints()
    poses = np.zeros((S, 7))
    yaws = np.zeros(S)
    scans = np.zeros((S, SCAN_BEAMS), dtype=np.float32)
    for s in range(S):
        i = rng.integers(0, len(xy))
        x, y, yaw = float(xy[i, 0]), float(xy[i, 1]), float(ori[i])
        boxes = []
        for _ in range(rng.integers(1, 4)):
            d = rng.uniform(0.8, 2.0)
            lat = rng.uniform(-0.8, 0.8)
            boxes.append((x + d * np.cos(yaw) - lat * np.sin(yaw), y + d * np.sin(yaw) + lat * np.cos(yaw), rng.uniform(0.15, 0.3)))
        poses[s] = yaw_pose(x, y, yaw)
        yaws[s] = yaw
        scans[s] = render_scan(x, y, yaw, boxes)
    return poses, yaws, scans


# ---- BASELINE.json configurations as record batches ---------------------------------------------
def config1_records(n=500, N=30):
    """Config 1: one tracking QP per skirk waypoint i = 0..n-1 (pose on the raceline, free scan -> the straight-ahead mini-path is
    valid and closest to the look-ahead point on straights; here the reference of cycle i is the mini-path whose end point is
    nearest the raceline 2.5 m ahead), previous steering filled in by the caller (warm-started sequence)."""
    xy, ori = skirk_waypoints()
    table = traj_table()
    ends = table[:, -1, :2]
    recs = np.zeros((n, record_doubles(N)))
    for i in range(n):
        x, y, yaw = float(xy[i % len(xy), 0]), float(xy[i % len(xy), 1]), float(ori[i % len(xy)])
        # look-ahead target: first raceline point at least 2.5 m ahead along the index order
        j = i
        while np.hypot(xy[j % len(xy), 0] - x, xy[j % len(xy), 1] - y) < 2.5 and j < i + len(xy):
            j += 1
        tx, ty = float(xy[j % len(xy), 0]) - x, float(xy[j % len(xy), 1]) - y
        c, sn = np.cos(yaw), np.sin(yaw)
        tcar = np.array([c * tx + sn * ty, -sn * tx + c * ty])
        p = int(np.argmin(np.hypot(ends[:, 0] - tcar[0], ends[:, 1] - tcar[1])))
        ref = np.zeros((N, 3))
        ref[:, :2] = path_to_world(table[p, :N, :2], x, y, yaw)
        recs[i] = np.concatenate([[x, y, yaw], [4.5, 0.0], np.zeros(6), ref.reshape(-1)])
    return recs


def config3_scans(B=1024, seed=20240903):
    """Config 3: B synthetic scans — ranges U(0.5, 2.8) with 1-3 gaps (U(3.5, 10)) 8..200 beams wide inside the field of view."""
    rng = np.random.default_rng(seed)
    scans = np.zeros((B, SCAN_BEAMS), dtype=np.float32)
    for b in range(B):
        r = rng.uniform(0.5, 2.8, SCAN_BEAMS).astype(np.float32)
        for _ in range(rng.integers(1, 4)):
            a = rng.integers(150, 880)
            w = rng.integers(8, 201)
            r[a:a + w] = rng.uniform(3.5, 10.0)
        scans[b] = r
    return scans


def config4_records(n_sc=64, N=30):
    """Config 4: 7 lanes x 20 mini-paths x 64 scenarios = 8960 QPs.  Lanes are not implemented in the reference (README:18);
    SURVEY 8d's synthetic definition: lane l = skirk shifted l x 0.25 m along the left normal, scenario s = station s x (500/64),
    QP (s, l, p) = ego on lane l at that station tracking mini-path p.  Scenario-major, so a rank's shard is contiguous."""
    xy, ori = skirk_waypoints()
    head = reference_data()["skirk_heading"]
    table20 = traj_table(steer_discrete=19)
    recs = []
    for sc in range(n_sc):
        i = int(sc * (500 / 64))
        for lane in range(7):
            nx, ny = -np.sin(head[i]), np.cos(head[i])
            x, y, yaw = float(xy[i, 0]) + lane * 0.25 * nx, float(xy[i, 1]) + lane * 0.25 * ny, float(ori[i])
            for pidx in range(20):
                ref = np.zeros((N, 3))
                ref[:, :2] = path_to_world(table20[pidx, :N, :2], x, y, yaw)
                recs.append(np.concatenate([[x, y, yaw], [4.5, 0.0], [0.3, -0.8, 1.5], [-0.4, 0.7, 2.0], ref.reshape(-1)]))
    return np.array(recs)
