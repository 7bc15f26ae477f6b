"""CPU oracle for the kbot-joystick rollout control step.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the arithmetic of the reference's hot path
(`/root/reference/train.py`, `convert.py`) plus the documented behaviour of the
un-vendored packages it calls (ksim fork b-vm/ksim@e88d8bc, xax 0.4.2, equinox 0.12.2,
distrax 0.1.5, jax 0.6.0 -- none of which are installable here).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and neither
JAX nor ksim can be imported in this image, so this oracle could not be checked against
outputs of the reference itself.  Everything marked [R] below follows train.py line by
line; everything marked [U] is the best available restatement of a third-party helper and
is isolated behind a named parameter (`OracleParams`) so a correction is a one-line change.
What pins the oracle instead: independent re-implementations (torch.nn.LSTMCell,
torch.distributions, scipy Rotation) in tests/golden/make_golden.py, fp64-vs-fp32
conditioning checks, and the structural invariants train.py asserts.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.  The product (kbot-joystick_b200/) never does.

Conventions
-----------
* Arrays are "AoS with leading batch axes": a per-env quantity of width F has shape
  [..., F] (the reference writes single-env code and ksim vmaps it).  Trajectory-wise
  functions take [T, ..., F] with time leading, like `ksim.Trajectory`.
* All arithmetic runs in the dtype of the inputs (float32 for parity, float64 for the
  conditioning cross-check).  Python scalars are weakly typed in NumPy 2 and do not upcast.
* Quaternions are (w, x, y, z).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------
# Constants [R]  train.py:22-70
# --------------------------------------------------------------------------------------

JOINT_NAMES = (
    "dof_left_hip_pitch_04",
    "dof_left_hip_roll_03",
    "dof_left_hip_yaw_03",
    "dof_left_knee_04",
    "dof_left_ankle_02",
    "dof_right_hip_pitch_04",
    "dof_right_hip_roll_03",
    "dof_right_hip_yaw_03",
    "dof_right_knee_04",
    "dof_right_ankle_02",
    "dof_right_shoulder_pitch_03",
    "dof_right_shoulder_roll_03",
    "dof_right_shoulder_yaw_02",
    "dof_right_elbow_02",
    "dof_right_wrist_00",
    "dof_left_shoulder_pitch_03",
    "dof_left_shoulder_roll_03",
    "dof_left_shoulder_yaw_02",
    "dof_left_elbow_02",
    "dof_left_wrist_00",
)

_BIAS_DEG = (20.0, 0.0, 0.0, 50.0, -30.0, -20.0, -0.0, 0.0, -50.0, 30.0,
             0.0, -10.0, 0.0, 90.0, 0.0, 0.0, 10.0, 0.0, -90.0, 0.0)
# train.py:24-45: math.radians(.) evaluated in double, cast to fp32 by jnp.array.
JOINT_BIASES64 = np.array([math.radians(d) for d in _BIAS_DEG], dtype=np.float64)

JOINT_LIMITS64 = np.array(
    [
        (-1.047198, 2.216568), (-0.20944, 2.268928), (-1.570796, 1.570796), (0.0, 2.70526),
        (-1.134464, 0.261799), (-2.216568, 1.047198), (-2.268928, 0.20944), (-1.570796, 1.570796),
        (-2.70526, 0.0), (-0.261799, 1.134464), (-3.490658, 1.047198), (-1.658063, 0.436332),
        (-1.671886, 1.671886), (0.0, 2.478368), (-1.37881, 1.37881), (-1.047198, 3.490658),
        (-0.436332, 1.658063), (-1.671886, 1.671886), (-2.478368, 0.0), (-1.37881, 1.37881),
    ],
    dtype=np.float64,
)  # train.py:47-68

NUM_JOINTS = 20
NUM_COMMANDS = 16
ACTOR_OBS = 65   # train.py:1290-1295
CRITIC_OBS = 475  # train.py:1297-1312

# robot/kbot/robot.mjcf body / sensordata indices (SURVEY 8a constants) [R]
BODY_BASE, BODY_LFOOT, BODY_RFOOT, NBODY = 1, 7, 12, 24
SD_GYRO, SD_IMU_QUAT, SD_TOUCH_L, SD_TOUCH_R, NSENSORDATA = 19, 28, 47, 48, 49

# robot/kbot/metadata.json kp/kd in NN order; ctrlrange from robot.mjcf:4-19 [R]
KP64 = np.array([150, 200, 100, 150, 40, 150, 200, 100, 150, 40,
                 100, 100, 40, 40, 20, 100, 100, 40, 40, 20], dtype=np.float64)
KD64 = np.array([24.722, 26.387, 3.419, 8.654, 0.990, 24.722, 26.387, 3.419, 8.654, 0.990,
                 8.284, 8.257, 0.945, 1.266, 0.295, 8.284, 8.257, 0.945, 1.266, 0.295], dtype=np.float64)
CTRL_LIMIT64 = np.array([120, 60, 60, 120, 17, 120, 60, 60, 120, 17,
                         60, 60, 17, 17, 14, 60, 60, 17, 17, 14], dtype=np.float64)

# get_rewards table train.py:1224-1256: name -> (scale, error_scale)
REWARD_NAMES = ("linvel", "angvel", "roll_pitch", "base_height", "arm_pos", "single_contact",
                "no_contact_p", "feet_airtime", "feet_orient", "com_distance", "base_accel", "torque")
REWARD_SCALES = (0.2, 0.1, 0.2, 0.2, 0.2, 0.1, 0.1, 1.5, 0.1, 0.05, 0.1, 0.1)


@dataclass
class OracleParams:
    """Every scalar the path depends on.  [U] entries are the unverified ones."""

    ctrl_dt: float = 0.02                     # train.py:1776
    hidden_size: int = 256                    # train.py:1773
    depth: int = 2                            # train.py:82-85
    min_std: float = 0.01                     # train.py:1320
    max_std: float = 1.0                      # train.py:1321
    var_scale: float = 0.5                    # train.py:86-89
    cutoff_frequency: float = 10.0            # train.py:90-93
    lpf_form: str = "rc"                      # [U] ksim.lowpass_one_pole coefficient form: "rc" | "exp"
    gamma: float = 0.94                       # train.py:1769
    lam: float = 0.94                         # train.py:1770
    normalize_advantages: int = 0             # [U] 0 none, 1 per-trajectory a/(std+eps)
    adv_eps: float = 1e-6                     # [U]
    actor_mirror_loss_scale: float = 0.0      # train.py:1771
    critic_mirror_loss_scale: float = 0.0     # train.py:1772
    # observation noise, train.py:1158-1198
    jpos_noise_mag: float = math.radians(3)
    jvel_noise_mag: float = math.radians(15)
    gyro_noise_std: float = math.radians(10)
    pg_noise_std: float = math.radians(3)
    gravity: float = 9.81                     # [U] ksim ProjectedGravityObservation
    # terminations train.py:1258-1269
    unhealthy_z: float = 0.4
    max_tilt: float = math.radians(45)
    max_length_sec: float = 12.0
    # command law train.py:1211-1221
    vx_range: tuple = (-0.5, 1.2)
    vy_range: tuple = (-0.5, 0.5)
    wz_range: tuple = (-1.0, 1.0)
    bh_range: tuple = (-0.25, 0.05)
    rx_range: tuple = (-0.25, 0.25)
    ry_range: tuple = (-0.25, 0.25)
    # rewards train.py:1224-1256
    reward_scales: tuple = REWARD_SCALES
    linvel_es: float = 0.2
    angvel_es: float = 0.2
    rp_es: float = 0.03
    rp_es_zero: float = 0.01
    bh_es: float = 0.02
    bh_standard: float = 0.80
    bh_foot_origin: float = 0.06
    arm_es: float = 0.1
    grace_period: float = 2.0
    touchdown_penalty: float = 0.4
    feet_es: float = 0.02
    com_es: float = 0.04
    acc_es: float = 5.0
    torque_es: float = 5.0
    eps_quat: float = 1e-6                    # [U] xax quaternion helpers

    @property
    def switch_prob(self) -> float:           # train.py:1220
        return self.ctrl_dt / 5

    @property
    def lpf_alpha(self) -> float:
        """[U] one-pole low-pass coefficient y' = y + alpha (x - y).  SURVEY Appendix F."""
        w = 2.0 * math.pi * self.cutoff_frequency
        if self.lpf_form == "rc":
            return w * self.ctrl_dt / (1.0 + w * self.ctrl_dt)
        if self.lpf_form == "exp":
            return 1.0 - math.exp(-w * self.ctrl_dt)
        raise ValueError(self.lpf_form)


def _c(x: np.ndarray, v):
    """Constant `v` in the dtype of x."""
    return np.asarray(v, dtype=x.dtype)


def joint_biases(dtype=np.float32) -> np.ndarray:
    return JOINT_BIASES64.astype(dtype)


def max_joint_range(dtype=np.float32) -> np.ndarray:
    """train.py:1330-1332, evaluated in `dtype` like jnp would."""
    b = JOINT_BIASES64.astype(dtype)
    lim = JOINT_LIMITS64.astype(dtype)
    return np.maximum(b - lim[:, 0], lim[:, 1] - b)


# --------------------------------------------------------------------------------------
# xax quaternion helpers [U]  
# --------------------------------------------------------------------------------------


def quat_to_euler(q: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    q = q / (np.sqrt(np.sum(q * q, axis=-1, keepdims=True)) + _c(q, eps))
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two, one = _c(q, 2.0), _c(q, 1.0)
    roll = np.arctan2(two * (w * x + y * z), one - two * (x * x + y * y))
    sinp = two * (w * y - z * x)
    with np.errstate(invalid="ignore"):
        pitch = np.where(np.abs(sinp) >= one, np.sign(sinp) * _c(q, np.pi / 2.0), np.arcsin(np.clip(sinp, -1, 1)))
    yaw = np.arctan2(two * (w * z + x * y), one - two * (y * y + z * z))
    return np.stack([roll, pitch, yaw], axis=-1).astype(q.dtype)


def euler_to_quat(e: np.ndarray) -> np.ndarray:
    half = _c(e, 0.5)
    r, p, y = e[..., 0] * half, e[..., 1] * half, e[..., 2] * half
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    w = cr * cp * cy + sr * sp * sy
    x = sr * cp * cy - cr * sp * sy
    yq = cr * sp * cy + sr * cp * sy
    z = cr * cp * sy - sr * sp * cy
    q = np.stack([w, x, yq, z], axis=-1)
    return (q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))).astype(e.dtype)


def rotate_vector_by_quat(v: np.ndarray, q: np.ndarray, inverse: bool = False, eps: float = 1e-6) -> np.ndarray:
    q = q / (np.sqrt(np.sum(q * q, axis=-1, keepdims=True)) + _c(q, eps))
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    if inverse:
        x, y, z = -x, -y, -z
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    t = _c(q, 2.0)
    xx = w * w * vx + t * y * w * vz - t * z * w * vy + x * x * vx + t * y * x * vy + t * z * x * vz - z * z * vx - y * y * vx
    yy = t * x * y * vx + y * y * vy + t * z * y * vz + t * w * z * vx - z * z * vy + w * w * vy - t * w * x * vz - x * x * vy
    zz = t * x * z * vx + t * y * z * vy + z * z * vz - t * w * y * vx + w * w * vz + t * w * x * vy - y * y * vz - x * x * vz
    return np.stack([xx, yy, zz], axis=-1).astype(q.dtype)


def zero_cmd_mask(cmd: np.ndarray) -> np.ndarray:
    """||cmd[0:3]||_2 < 1e-3  (train.py:142,164,210,289,332,467,505,1365)."""
    c3 = cmd[..., :3]
    return np.sqrt(np.sum(c3 * c3, axis=-1)) < _c(cmd, 1e-3)


# --------------------------------------------------------------------------------------
# Observations  O1..O11
# --------------------------------------------------------------------------------------


def projected_gravity(imu_quat: np.ndarray, p: OracleParams) -> np.ndarray:
    """[U] ksim.ProjectedGravityObservation: rotate (0,0,-g) into the IMU frame.  train.py:1199-1202."""
    g = np.zeros(imu_quat.shape[:-1] + (3,), dtype=imu_quat.dtype)
    g[..., 2] = -p.gravity
    return rotate_vector_by_quat(g, imu_quat, inverse=True, eps=p.eps_quat)


def feet_position_obs(xpos: np.ndarray, xquat: np.ndarray, p: OracleParams) -> np.ndarray:
    """[R] FeetPositionObservation.observe, train.py:682-699.  xpos [...,24,3], xquat [...,24,4] -> [...,6]."""
    base_pos = xpos[..., BODY_BASE, :]
    yaw = quat_to_euler(xquat[..., BODY_BASE, :], p.eps_quat)[..., 2]
    e = np.stack([np.zeros_like(yaw), np.zeros_like(yaw), yaw], axis=-1)
    qyaw = euler_to_quat(e)
    fl = rotate_vector_by_quat(xpos[..., BODY_LFOOT, :] - base_pos, qyaw, inverse=True, eps=p.eps_quat)
    fr = rotate_vector_by_quat(xpos[..., BODY_RFOOT, :] - base_pos, qyaw, inverse=True, eps=p.eps_quat)
    return np.concatenate([fl, fr], axis=-1)


def get_observations(state: dict, noise: dict | None, episode: dict | None, pg_carry: np.ndarray | None,
                     p: OracleParams, reset: np.ndarray | None = None) -> tuple[dict, np.ndarray | None]:
    """[R]+[U] the get_observations table, train.py:1155-1204 (21 observations + 4 noisy twins).

    state: qpos[...,27] qvel[...,26] qacc[...,26] sensordata[...,49] xpos[...,24,3] xquat[...,24,4]
           cinert[...,24,10] cvel[...,24,6] actuator_force[...,20] com_distance[...]
    noise: eps_jpos[...,20] eps_jvel[...,20] in U(-1,1); eps_gyro[...,3] eps_pg[...,3] ~ N(0,1)
    episode: jpos_bias[...,20], pg_lag[...], pg_bias[...,3]  (per-episode randomisation)
    pg_carry: [...,3] EMA state of the lagged projected gravity (O4)
    reset: [...] bool or None -- where True the EMA state is re-initialised to the current projected gravity
           (first step of a new episode; [U] ksim re-inits observation carries on done)
    Returns (obs dict, new pg_carry).  com_distance is consumed as a recorded scalar.
    """
    qpos, qvel, sd = state["qpos"], state["qvel"], state["sensordata"]
    o: dict[str, np.ndarray] = {}
    o["joint_position"] = qpos[..., 7:]
    o["joint_velocity"] = qvel[..., 6:]
    o["actuator_force"] = state["actuator_force"]
    o["center_of_mass_inertia"] = state["cinert"][..., 1:, :].reshape(qpos.shape[:-1] + (230,))
    o["center_of_mass_velocity"] = state["cvel"][..., 1:, :].reshape(qpos.shape[:-1] + (138,))
    o["base_position"] = qpos[..., 0:3]
    o["base_orientation"] = qpos[..., 3:7]
    o["base_linear_velocity"] = qvel[..., 0:3]
    o["base_angular_velocity"] = qvel[..., 3:6]
    if "qacc" in state:
        qacc = state["qacc"]
        o["base_linear_acceleration"] = qacc[..., 0:3]
        o["base_angular_acceleration"] = qacc[..., 3:6]
        o["actuator_acceleration"] = qacc[..., 6:]
    o["imu_gyro"] = sd[..., SD_GYRO:SD_GYRO + 3]
    o["left_foot_touch"] = sd[..., SD_TOUCH_L:SD_TOUCH_L + 1]
    o["right_foot_touch"] = sd[..., SD_TOUCH_R:SD_TOUCH_R + 1]
    o["feet_position"] = feet_position_obs(state["xpos"], state["xquat"], p)
    o["base_height"] = state["xpos"][..., BODY_BASE, 2:]            # train.py:706-707
    g_b = projected_gravity(sd[..., SD_IMU_QUAT:SD_IMU_QUAT + 4], p)
    o["projected_gravity"] = g_b
    o["imu_projected_gravity"] = g_b
    o["com_distance"] = state["com_distance"]
    o["biased_joint_position"] = o["joint_position"]
    new_carry = pg_carry
    if noise is not None:
        x = o["joint_position"]
        jb = episode["jpos_bias"] if episode is not None else np.zeros_like(x)
        # O5 [U fork]: qpos[7:] + per-episode bias + U(+-mag) step noise
        o["biased_joint_position"] = x + jb
        o["noisy_biased_joint_position"] = (x + jb) + _c(x, p.jpos_noise_mag) * noise["eps_jpos"]
        o["noisy_joint_velocity"] = o["joint_velocity"] + _c(x, p.jvel_noise_mag) * noise["eps_jvel"]
        o["noisy_imu_gyro"] = o["imu_gyro"] + _c(x, p.gyro_noise_std) * noise["eps_gyro"]
        # O4 [U]: EMA lag, then bias, then gaussian noise
        lag = episode["pg_lag"][..., None] if episode is not None else np.zeros_like(g_b[..., :1])
        pgb = episode["pg_bias"] if episode is not None else np.zeros_like(g_b)
        prev = pg_carry if pg_carry is not None else g_b
        if reset is not None:
            prev = np.where(reset[..., None], g_b, prev)
        new_carry = lag * prev + (_c(x, 1.0) - lag) * g_b
        o["imu_projected_gravity"] = new_carry + pgb
        o["noisy_imu_projected_gravity"] = (new_carry + pgb) + _c(x, p.pg_noise_std) * noise["eps_pg"]
    return o, new_carry


def normalize_joint_pos(q: np.ndarray) -> np.ndarray:
    """[R] train.py:1329-1333."""
    return (q - joint_biases(q.dtype)) / max_joint_range(q.dtype)


def normalize_joint_vel(v: np.ndarray) -> np.ndarray:
    """[R] train.py:1335-1336."""
    return v / _c(v, 10.0)


def encode_projected_gravity(g: np.ndarray) -> np.ndarray:
    """[R] train.py:1338-1349."""
    gx, gy, gz = g[..., 0], g[..., 1], g[..., 2]
    roll = np.arctan2(gy, -gz)
    pitch = np.arctan2(-gx, np.sqrt(gy * gy + gz * gz))
    unit = g / np.sqrt(np.sum(g * g, axis=-1, keepdims=True))
    return np.concatenate([roll[..., None], pitch[..., None], unit], axis=-1).astype(g.dtype)


def actor_obs(joint_pos, joint_vel, proj_grav, gyro, cmd) -> np.ndarray:
    """[R] run_actor concat, train.py:1365-1376 (identical to convert.py:93-105). -> [...,65]."""
    zc = zero_cmd_mask(cmd)[..., None].astype(cmd.dtype)
    return np.concatenate(
        [normalize_joint_pos(joint_pos), normalize_joint_vel(joint_vel), encode_projected_gravity(proj_grav),
         gyro, zc, cmd], axis=-1)


def actor_obs_from_dict(o: dict, cmd: np.ndarray) -> np.ndarray:
    return actor_obs(o["noisy_biased_joint_position"], o["noisy_joint_velocity"],
                     o["noisy_imu_projected_gravity"], o["noisy_imu_gyro"], cmd)


def critic_obs_from_dict(o: dict, cmd: np.ndarray) -> np.ndarray:
    """[R] run_critic concat, train.py:1388-1431. -> [...,475]."""
    a = actor_obs(o["joint_position"], o["joint_velocity"], o["projected_gravity"], o["imu_gyro"], cmd)
    return np.concatenate(
        [a, o["left_foot_touch"], o["right_foot_touch"], o["feet_position"], o["base_position"],
         o["base_orientation"], o["center_of_mass_inertia"], o["center_of_mass_velocity"],
         o["base_linear_velocity"], o["base_angular_velocity"], o["actuator_force"] / _c(a, 4.0),
         o["base_height"]], axis=-1)


# --------------------------------------------------------------------------------------
# O12 COMDistanceObservation  train.py:509-659  [R]
# --------------------------------------------------------------------------------------


def _com_distance_single(geom1, geom2, pos, com, dt):
    """One env.  geom1/geom2 int [ncon], pos [ncon,3], com [3] (subtree_com[2]).  Loops follow the jnp code line by line:
    observe (train.py:635-659), monotone_chain_hull (561-633), polygon_centroid_masked (519-559)."""
    n = geom1.shape[0]
    f = dt.type
    # num_unique (train.py:642-647): distinct values of contact.geom2, padding entries included -- as written
    sx = np.sort(geom2.ravel())
    unique = 0 if n == 0 else int(np.sum(sx[1:] != sx[:-1])) + 1
    if unique < 3:
        return f(-1.0)
    # feet_to_floor_contacts (train.py:637-638): non-floor rows become the origin and STAY in the point set
    pts = np.where((geom1 == 0)[:, None], pos, np.zeros_like(pos))[:, :2].astype(dt)
    order = np.lexsort((pts[:, 1], pts[:, 0])).astype(np.int32)              # by x then y, stable
    spts = pts[order]

    def cross(a, b, c):
        return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])

    def build(indices):
        stack = -np.ones((n,), np.int32)
        ptr = 0
        for idx in indices:
            while ptr >= 2 and cross(spts[stack[ptr - 2]], spts[stack[ptr - 1]], spts[idx]) <= 0:
                stack[ptr - 1] = -1
                ptr -= 1
            stack[ptr] = idx
            ptr += 1
        return stack, ptr

    sorted_idxs = np.arange(n, dtype=np.int32)
    stack_l, ptr_l = build(sorted_idxs)
    stack_u, ptr_u = build(sorted_idxs[::-1])
    idxs = np.arange(n, dtype=np.int32)
    lower_mask = idxs < max(ptr_l - 1, 0)
    upper_mask = idxs < max(ptr_u - 1, 0)
    hull_pos = np.concatenate([np.where(lower_mask, stack_l, -1), np.where(upper_mask, stack_u, -1)])
    hull_mask = np.concatenate([lower_mask, upper_mask])
    hull_pts = pts[order[np.where(hull_mask, hull_pos, 0)]]
    # polygon_centroid_masked
    L = hull_pts.shape[0]
    ii = np.arange(L, dtype=np.int32)
    count = int(hull_mask.sum())
    valid = np.zeros((L,), np.int64)
    nz = np.nonzero(hull_mask)[0]
    valid[:nz.shape[0]] = nz                                                  # jnp.nonzero(size=L) pads with 0
    packed = hull_pts[valid]
    nxt = packed[np.where(ii + 1 < count, ii + 1, 0)]
    edge = ((ii < max(count - 1, 0)) | ((ii == max(count - 1, 0)) & (count > 0))).astype(dt)
    x, y, x1, y1 = packed[:, 0], packed[:, 1], nxt[:, 0], nxt[:, 1]
    cr = (x * y1 - x1 * y) * edge
    area = f(0.5) * np.sum(cr, dtype=dt)
    first = (ii < max(count, 0)).astype(dt)
    count_f = max(first.sum(dtype=dt), f(1.0))
    mean_pt = np.sum(packed * first[:, None], axis=0, dtype=dt) / count_f
    if abs(area) < 1e-12:
        cx, cy = mean_pt[0], mean_pt[1]
    else:
        cx = np.sum((x + x1) * cr, dtype=dt) / (f(6.0) * area)
        cy = np.sum((y + y1) * cr, dtype=dt) / (f(6.0) * area)
    d = np.array([cx - com[0], cy - com[1]], dt)
    return np.sqrt(np.sum(d * d, dtype=dt))


def com_distance_observation(geom1, geom2, pos, subtree_com_base):
    """[R] COMDistanceObservation.observe, train.py:635-659, batched over leading axes.
    geom1, geom2 int [..., ncon]; pos [..., ncon, 3]; subtree_com_base [..., 3] (= data.subtree_com[2]) -> [...]."""
    dt = pos.dtype
    lead = geom1.shape[:-1]
    g1 = geom1.reshape(-1, geom1.shape[-1]); g2 = geom2.reshape(-1, geom2.shape[-1])
    ps = pos.reshape(-1, pos.shape[-2], 3); cm = subtree_com_base.reshape(-1, 3)
    out = np.empty((g1.shape[0],), dt)
    for i in range(g1.shape[0]):
        out[i] = _com_distance_single(g1[i], g2[i], ps[i], cm[i], np.dtype(dt))
    return out.reshape(lead)


# --------------------------------------------------------------------------------------
# Commands C1, C2   train.py:710-785, ranges 1206-1222
# --------------------------------------------------------------------------------------


def initial_command(mode: np.ndarray, u6: np.ndarray, u_arms: np.ndarray, p: OracleParams) -> np.ndarray:
    """[R] UnifiedCommand.initial_command with the randomness made explicit.

    mode   [...]    int in 0..5      (jax.random.randint(rng_a, (), 0, 6))
    u6     [...,6]  uniforms in [0,1) for vx, vy, wz, bh, rx, ry (keys b..g)
    u_arms [...,10] uniforms in [0,1) from key h.  train.py:734-737 feeds the SAME key to
           uniform() and bernoulli(), so [U: jax.random semantics] the mask is (u_arms < 0.5) on
           the very same draws: arms = (lo + u*(hi-lo)) * (u < 0.5).
    """
    dt = u6.dtype
    rng = [p.vx_range, p.vy_range, p.wz_range, p.bh_range, p.rx_range, p.ry_range]
    vals = [np.asarray(lo, dt) + u6[..., k] * (np.asarray(hi, dt) - np.asarray(lo, dt)) for k, (lo, hi) in enumerate(rng)]
    vx, vy, wz, bh, rx, ry = vals
    lim = JOINT_LIMITS64.astype(dt)[10:20]
    arms = (lim[:, 0] + u_arms * (lim[:, 1] - lim[:, 0])) * (u_arms < np.asarray(0.5, dt)).astype(dt)
    z = np.zeros_like(vx)
    za = np.zeros_like(arms)

    def cat(a, b, c, d, e, f, g):
        return np.concatenate([np.stack([a, b, c, d, e, f], axis=-1), g], axis=-1)

    cmds = [cat(vx, z, z, z, z, z, za), cat(z, vy, z, z, z, z, za), cat(z, z, wz, z, z, z, za),
            cat(vx, vy, wz, z, z, z, arms), cat(z, z, z, bh, rx, ry, arms), cat(z, z, z, z, z, z, za)]
    out = np.zeros_like(cmds[0])
    for m in range(6):
        out = np.where((mode == m)[..., None], cmds[m], out)
    return out


def command_step(prev_cmd, u_switch, mode, u6, u_arms, p: OracleParams) -> np.ndarray:
    """[R] UnifiedCommand.__call__, train.py:768-785: keep prev unless bernoulli(ctrl_dt/5)."""
    new = initial_command(mode, u6, u_arms, p)
    switch = u_switch < np.asarray(p.switch_prob, prev_cmd.dtype)   # [U] jax bernoulli = uniform < p
    return np.where(switch[..., None], new, prev_cmd)


# --------------------------------------------------------------------------------------
# Networks M1..M6   train.py:847-1046, 1435-1572
# --------------------------------------------------------------------------------------


def sigmoid(x):
    return _c(x, 1.0) / (_c(x, 1.0) + np.exp(-x))


def softplus(x):
    """[U] jax.nn.softplus = logaddexp(x, 0)."""
    return np.maximum(x, _c(x, 0.0)) + np.log1p(np.exp(-np.abs(x)))


def linear(w, b, x):
    """[U] eqx.nn.Linear: W @ x + b, W:[out,in]."""
    return x @ w.T + b


def lstm_cell(w_ih, w_hh, b, x, h, c):
    """[U] eqx.nn.LSTMCell 0.12.2: lin = W_ih x + W_hh h + b; i,f,g,o = split(lin, 4)."""
    lin = x @ w_ih.T + h @ w_hh.T + b
    hs = h.shape[-1]
    i, f, g, o = lin[..., :hs], lin[..., hs:2 * hs], lin[..., 2 * hs:3 * hs], lin[..., 3 * hs:]
    c2 = sigmoid(f) * c + sigmoid(i) * np.tanh(g)
    h2 = sigmoid(o) * np.tanh(c2)
    return h2, c2


def init_net_weights(rng: np.random.Generator, num_inputs: int, num_outputs: int, hidden: int, depth: int,
                     dtype=np.float32) -> dict:
    """Random weights in eqx layout with eqx's U(+-1/sqrt(fan_in)) init law (values are synthetic)."""
    def u(shape, fan_in):
        lim = 1.0 / math.sqrt(fan_in)
        return rng.uniform(-lim, lim, size=shape).astype(dtype)

    w = {"w_in": u((hidden, num_inputs), num_inputs), "b_in": u((hidden,), num_inputs),
         "w_out": u((num_outputs, hidden), hidden), "b_out": u((num_outputs,), hidden), "layers": []}
    for _ in range(depth):
        w["layers"].append({"w_ih": u((4 * hidden, hidden), hidden), "w_hh": u((4 * hidden, hidden), hidden),
                            "b": u((4 * hidden,), hidden)})
    return w


def trunk_forward(w: dict, obs, carry):
    """input_proj -> depth x LSTMCell -> output_proj.  carry [..., depth, 2, H].  train.py:916-922, 996-1002."""
    x = linear(w["w_in"], w["b_in"], obs)
    new_carry = np.empty_like(carry)
    for li, lw in enumerate(w["layers"]):
        h, c = lstm_cell(lw["w_ih"], lw["w_hh"], lw["b"], x, carry[..., li, 0, :], carry[..., li, 1, :])
        new_carry[..., li, 0, :] = h
        new_carry[..., li, 1, :] = c
        x = h
    return linear(w["w_out"], w["b_out"], x), new_carry


def actor_forward(w: dict, obs, carry, lpf, p: OracleParams):
    """[R] Actor.forward train.py:913-941.  Returns (mean, std, new_carry, new_lpf)."""
    out, new_carry = trunk_forward(w, obs, carry)
    mean = out[..., :NUM_JOINTS]
    std = out[..., NUM_JOINTS:]
    std = np.minimum((softplus(std) + _c(out, p.min_std)) * _c(out, p.var_scale), _c(out, p.max_std))
    arm_bias = np.concatenate([np.zeros_like(obs[..., :10]), obs[..., -10:]], axis=-1)
    mean = mean + joint_biases(out.dtype) + arm_bias
    # [U] ksim.lowpass_one_pole: y' = y + alpha (x - y), state = 20 floats (convert.py:71)
    new_lpf = lpf + _c(out, p.lpf_alpha) * (mean - lpf)
    return new_lpf, std, new_carry, new_lpf


def critic_forward(w: dict, obs, carry):
    """[R] Critic.forward train.py:993-1004. Returns (value[...,1], new_carry)."""
    return trunk_forward(w, obs, carry)


LOG_2PI = math.log(2.0 * math.pi)


def mvn_log_prob(mean, std, a):
    """[U] distrax.MultivariateNormalDiag.log_prob."""
    z = (a - mean) / std
    return np.sum(_c(a, -0.5) * z * z - _c(a, 0.5 * LOG_2PI), axis=-1) - np.sum(np.log(std), axis=-1)


def mvn_entropy(std):
    """[U] distrax.MultivariateNormalDiag.entropy."""
    return np.sum(np.log(std), axis=-1) + _c(std, std.shape[-1] * (0.5 + 0.5 * LOG_2PI))


def sample_action(w_actor, obs, carry, lpf, eps, argmax: bool, p: OracleParams):
    """[R] sample_action train.py:1545-1572.  eps ~ N(0, I_20) supplied explicitly."""
    mean, std, nc, nl = actor_forward(w_actor, obs, carry, lpf, p)
    a = mean if argmax else mean + std * eps
    return a, mean, std, nc, nl


def policy_step(w_actor, joint_angles, joint_vel, proj_grav, gyro, command, carry_flat, p: OracleParams):
    """[R] convert.py:84-119 step_fn.  carry_flat [..., depth*2*H + 20] = ravel(actor_carry, lpf)."""
    hs, d = p.hidden_size, p.depth
    obs = actor_obs(joint_angles, joint_vel, proj_grav, gyro, command)
    carry = carry_flat[..., : d * 2 * hs].reshape(carry_flat.shape[:-1] + (d, 2, hs))
    lpf = carry_flat[..., d * 2 * hs:]
    mean, _, nc, nl = actor_forward(w_actor, obs, carry, lpf, p)
    return mean, np.concatenate([nc.reshape(carry_flat.shape[:-1] + (-1,)), nl], axis=-1)


# ---- mirror maps X1  train.py:1574-1756 ----------------------------------------------


def mirror_joints(j):
    """[R] train.py:1574-1582: negate all, swap legs only (arms stay in place -- as written)."""
    return -np.concatenate([j[..., 5:10], j[..., 0:5], j[..., 10:15], j[..., 15:20]], axis=-1)


def _sgn(x, signs):
    return x * np.asarray(signs, dtype=x.dtype)


def mirror_obs(o: dict) -> dict:
    """[R] train.py:1584-1733."""
    ci = o["center_of_mass_inertia"].reshape(o["center_of_mass_inertia"].shape[:-1] + (23, 10))
    cv = o["center_of_mass_velocity"].reshape(o["center_of_mass_velocity"].shape[:-1] + (23, 6))
    fp = o["feet_position"]
    return {
        "noisy_biased_joint_position": mirror_joints(o["noisy_biased_joint_position"]),
        "noisy_joint_velocity": mirror_joints(o["noisy_joint_velocity"]),
        "noisy_imu_gyro": _sgn(o["noisy_imu_gyro"], (-1, 1, -1)),
        "noisy_imu_projected_gravity": _sgn(o["noisy_imu_projected_gravity"], (1, -1, 1)),
        "joint_position": mirror_joints(o["joint_position"]),
        "joint_velocity": mirror_joints(o["joint_velocity"]),
        "imu_gyro": _sgn(o["imu_gyro"], (-1, 1, -1)),
        "projected_gravity": _sgn(o["projected_gravity"], (1, -1, 1)),
        "left_foot_touch": o["right_foot_touch"],
        "right_foot_touch": o["left_foot_touch"],
        "feet_position": _sgn(np.concatenate([fp[..., 3:6], fp[..., 0:3]], axis=-1), (1, -1, 1, 1, -1, 1)),
        "base_position": o["base_position"],
        "base_orientation": _sgn(o["base_orientation"], (1, -1, -1, 1)),
        "center_of_mass_inertia": _sgn(ci, (1, 1, -1, 1, 1, 1, 1, -1, 1, -1)).reshape(o["center_of_mass_inertia"].shape),
        "center_of_mass_velocity": _sgn(cv, (1, -1, 1, -1, 1, -1)).reshape(o["center_of_mass_velocity"].shape),
        "base_linear_velocity": _sgn(o["base_linear_velocity"], (1, -1, 1)),
        "base_angular_velocity": _sgn(o["base_angular_velocity"], (-1, 1, -1)),
        "actuator_force": mirror_joints(o["actuator_force"]),
        "base_height": o["base_height"],
    }


def mirror_cmd(cmd):
    """[R] train.py:1735-1756: (+vx,-vy,-wz,+bh,-rx,+ry, -arms unswapped)."""
    pad = np.concatenate([np.zeros_like(cmd[..., :10]), cmd[..., 6:16]], axis=-1)
    return np.concatenate([cmd[..., 0:1], -cmd[..., 1:2], -cmd[..., 2:3], cmd[..., 3:4], -cmd[..., 4:5],
                           cmd[..., 5:6], mirror_joints(pad)[..., 10:20]], axis=-1)


def initial_model_carry(batch_shape: tuple, p: OracleParams, dtype=np.float32) -> dict:
    """[R] get_initial_model_carry train.py:1526-1543 (zeros; LPF init [U] zeros, 20 floats)."""
    z = lambda: np.zeros(batch_shape + (p.depth, 2, p.hidden_size), dtype)
    return {"actor": z(), "actor_mirror": z(), "critic": z(), "critic_mirror": z(),
            "lpf_params": np.zeros(batch_shape + (NUM_JOINTS,), dtype),
            "lpf_params_mirror": np.zeros(batch_shape + (NUM_JOINTS,), dtype)}


def ppo_scan_step(w_actor, w_critic, carry: dict, obs: dict, cmd, action, done, p: OracleParams, mirror: bool = True):
    """[R] _ppo_scan_fn train.py:1435-1508 on one stored transition (batched over leading axes)."""
    a_obs = actor_obs_from_dict(obs, cmd)
    mean, std, na, nl = actor_forward(w_actor, a_obs, carry["actor"], carry["lpf_params"], p)
    log_prob = mvn_log_prob(mean, std, action)
    value, ncr = critic_forward(w_critic, critic_obs_from_dict(obs, cmd), carry["critic"])
    new = dict(carry)
    new.update(actor=na, critic=ncr, lpf_params=nl)
    out = {"log_probs": log_prob[..., None], "values": value[..., 0], "entropy": mvn_entropy(std)[..., None],
           "action_std": std, "mean": mean}
    if mirror:
        mo, mc = mirror_obs(obs), mirror_cmd(cmd)
        mmean, _, nam, nlm = actor_forward(w_actor, actor_obs_from_dict(mo, mc), carry["actor_mirror"],
                                           carry["lpf_params_mirror"], p)
        dm = mirror_joints(mmean)
        out["action_mirror_loss"] = np.mean((mean - dm) ** 2, axis=-1) * _c(mean, p.actor_mirror_loss_scale)
        mval, ncm = critic_forward(w_critic, critic_obs_from_dict(mo, mc), carry["critic_mirror"])
        out["value_mirror_loss"] = np.mean((value - mval) ** 2, axis=-1) * _c(mean, p.critic_mirror_loss_scale)
        out["mirror_mean"] = mmean
        out["mirror_value"] = mval[..., 0]
        new.update(actor_mirror=nam, critic_mirror=ncm, lpf_params_mirror=nlm)
    # carry <- where(done, initial, new)   train.py:1502-1506
    for k in new:
        d = done.reshape(done.shape + (1,) * (new[k].ndim - done.ndim))
        new[k] = np.where(d, np.zeros_like(new[k]), new[k])
    return new, out


def get_ppo_variables(w_actor, w_critic, traj_obs: list, traj_cmd, traj_action, traj_done, carry: dict,
                      p: OracleParams, mirror: bool = True):
    """[R] get_ppo_variables train.py:1510-1524: scan of ppo_scan_step over T (traj_obs: list of T obs dicts)."""
    outs = []
    for t in range(len(traj_obs)):
        carry, o = ppo_scan_step(w_actor, w_critic, carry, traj_obs[t], traj_cmd[t], traj_action[t], traj_done[t], p, mirror)
        outs.append(o)
    return {k: np.stack([o[k] for o in outs], axis=0) for k in outs[0]}, carry


# --------------------------------------------------------------------------------------
# PPO loss [U: ksim.compute_ppo_loss; entropy_coef train.py:1767]
# --------------------------------------------------------------------------------------


def ppo_loss(log_probs, old_log_probs, advantages, values, old_values, value_targets, entropy, clip_param=0.2,
             value_loss_coef=0.5, entropy_coef=0.004, log_clip_value=10.0, use_clipped_value_loss=True):
    """[U] ksim.compute_ppo_loss restated from its published form (oracle-defined, reference-unverified):
      ratio     = exp(clip(log_probs - old_log_probs, +-log_clip_value))
      policy    = min(ratio A, clip(ratio, 1 - eps, 1 + eps) A)
      value     = 0.5 max((target - v)^2, (target - (v_old + clip(v - v_old, +-eps)))^2)   (unclipped: 0.5 (target - v)^2)
      objective = policy - value_loss_coef value + entropy_coef entropy;    loss = -mean(objective)
    All arrays [T, N] (log-probs / entropy already summed over the 20 action dimensions).
    Returns (loss, mean policy, mean value, mean entropy, per-step objective [T, N])."""
    dt = log_probs.dtype
    c = lambda v: np.asarray(v, dt)
    log_ratio = np.clip(log_probs - old_log_probs, -c(log_clip_value), c(log_clip_value))
    ratio = np.exp(log_ratio)
    pol = np.minimum(ratio * advantages, np.clip(ratio, c(1.0) - c(clip_param), c(1.0) + c(clip_param)) * advantages)
    err = value_targets - values
    val = c(0.5) * err * err
    if use_clipped_value_loss:
        vc = old_values + np.clip(values - old_values, -c(clip_param), c(clip_param))
        errc = value_targets - vc
        val = c(0.5) * np.maximum(err * err, errc * errc)
    obj = pol - c(value_loss_coef) * val + c(entropy_coef) * entropy
    f64 = np.float64
    return (-obj.astype(f64).mean(), pol.astype(f64).mean(), val.astype(f64).mean(), entropy.astype(f64).mean(), obj)


# --------------------------------------------------------------------------------------
# Actuators A1 [U fork]  train.py:1091-1105
# --------------------------------------------------------------------------------------


def position_actuator_torque(action, q, qd, kp=None, kd=None, tau_limit=None, action_bias=None, torque_bias=None):
    """tau = clip(kp (a + b_a - q) - kd qd + b_tau, +-tau_lim).  Gains/limits/biases are per-env inputs (F8)."""
    dt = action.dtype
    kp = KP64.astype(dt) if kp is None else kp
    kd = KD64.astype(dt) if kd is None else kd
    lim = CTRL_LIMIT64.astype(dt) if tau_limit is None else tau_limit
    target = action if action_bias is None else action + action_bias
    tau = kp * (target - q) - kd * qd
    if torque_bias is not None:
        tau = tau + torque_bias
    return np.minimum(np.maximum(tau, -lim), lim)


def position_actuator_substeps(action, prev_action, u_drop, latency, q_sub, qd_sub, sub_dt=0.004, drop_prob=0.05,
                               **gains):
    """[U] the per-physics-sub-step actuator path of ksim's engine (train.py:1775-1781: dt = 0.004,
    ctrl_dt = 0.02, action_latency_range = (0.003, 0.01), drop_action_prob = 0.05) -- oracle-defined, reference-unverified:
      applied = prev_action if u_drop < drop_prob else action              (a dropped command repeats the last applied one)
      sub-step k (time k sub_dt since the control tick) sees `applied` once k sub_dt >= latency, else prev_action
      ctrl_k = PositionActuators.get_ctrl(that action, q_k, qd_k)          (position_actuator_torque)
    action / prev_action [..., 20], u_drop / latency [...], q_sub / qd_sub [S, ..., 20].
    Returns (ctrl [S, ..., 20], new prev_action = applied)."""
    dt = action.dtype
    applied = np.where((u_drop < np.asarray(drop_prob, dt))[..., None], prev_action, action)
    out = []
    for k in range(q_sub.shape[0]):
        seen = (np.asarray(k, dt) * np.asarray(sub_dt, dt) >= latency)[..., None]
        out.append(position_actuator_torque(np.where(seen, applied, prev_action), q_sub[k], qd_sub[k], **gains))
    return np.stack(out), applied


def actuator_randomization(u, kp_scale=1.4, kd_scale=1.4, torque_limit_scale_low=0.5, action_bias_scale=0.02,
                           torque_bias_scale=0.0):
    """[U fork] per-episode randomisation of ksim.PositionActuators (arguments [R] train.py:1097-1105; the sampling law is
    oracle-defined, reference-unverified -- tools/verify_against_ref.py item `position_actuators`):
      kp = kp_nominal U(1/kp_scale, kp_scale), kd likewise, tau_limit = ctrl_limit U(low, 1),
      action_bias = U(-s_a, s_a), torque_bias = U(-s_t, s_t);  u [5, ..., 20] uniforms in [0, 1)."""
    dt = u.dtype
    c = lambda v: np.asarray(v, dt)

    def lerp(uu, lo, hi, nominal):
        return nominal * (c(lo) + uu * (c(hi) - c(lo)))

    return {"kp": lerp(u[0], c(1.0) / c(kp_scale), kp_scale, KP64.astype(dt)),
            "kd": lerp(u[1], c(1.0) / c(kd_scale), kd_scale, KD64.astype(dt)),
            "tau_limit": lerp(u[2], torque_limit_scale_low, 1.0, CTRL_LIMIT64.astype(dt)),
            "action_bias": lerp(u[3], -action_bias_scale, action_bias_scale, c(1.0)),
            "torque_bias": lerp(u[4], -torque_bias_scale, torque_bias_scale, c(1.0))}


# --------------------------------------------------------------------------------------
# Optimiser [R] train.py:1059-1077 (launch config: adam_weight_decay = 1e-5 -> optax.adamw) + ksim's clipping [U]
# --------------------------------------------------------------------------------------


def adamw_update(param, grad, m, v, step: int, lr=5e-4, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-5, grad_scale=1.0,
                 max_grad_norm=10.0):
    """optax.adamw(lr, weight_decay) = chain(scale_by_adam(b1, b2, eps, eps_root = 0), add_decayed_weights(wd),
    scale_by_learning_rate(lr)) applied with eqx.apply_updates, in float64 on flat arrays; `step` = count + 1 (>= 1).
    Around it ksim's update [U]: global-norm clip to max_grad_norm (<= 0: off), and no update at all (state untouched)
    when the norm is not finite.  Returns (param, m, v, applied)."""
    g = np.asarray(grad, np.float64) * grad_scale
    norm = float(np.sqrt(np.sum(g * g)))
    if not np.isfinite(norm):
        return np.asarray(param, np.float64), np.asarray(m, np.float64), np.asarray(v, np.float64), False
    if max_grad_norm > 0 and norm > max_grad_norm:
        g = g * (max_grad_norm / max(norm, 1e-6))
    m2 = b1 * np.asarray(m, np.float64) + (1 - b1) * g
    v2 = b2 * np.asarray(v, np.float64) + (1 - b2) * g * g
    mh, vh = m2 / (1 - b1 ** step), v2 / (1 - b2 ** step)
    p64 = np.asarray(param, np.float64)
    return p64 - lr * (mh / (np.sqrt(vh) + eps) + weight_decay * p64), m2, v2, True


# --------------------------------------------------------------------------------------
# Terminations Z1  train.py:1258-1269, 817-823
# --------------------------------------------------------------------------------------


def bad_z_height(xpos):
    """pre-threshold value of TerrainBadZTermination (train.py:818-822)."""
    return xpos[..., BODY_BASE, 2] - np.minimum(xpos[..., BODY_LFOOT, 2], xpos[..., BODY_RFOOT, 2])


def upright_tilt(quat, p: OracleParams):
    """[U] ksim.NotUprightTermination: angle between body z axis and world z.  pre-threshold value."""
    q = quat / (np.sqrt(np.sum(quat * quat, axis=-1, keepdims=True)) + _c(quat, p.eps_quat))
    x, y = q[..., 1], q[..., 2]
    cz = _c(quat, 1.0) - _c(quat, 2.0) * (x * x + y * y)
    return np.arccos(np.clip(cz, -1.0, 1.0)).astype(quat.dtype)


def terminations(xpos, base_quat, time, p: OracleParams):
    """Returns (codes int32 [...,3] = (bad_z, not_upright, episode_length), done bool, success bool)."""
    bad_z = np.where(bad_z_height(xpos) < _c(xpos, p.unhealthy_z), -1, 0).astype(np.int32)
    tilt = np.where(upright_tilt(base_quat, p) > _c(xpos, p.max_tilt), -1, 0).astype(np.int32)
    ep = np.where(time > _c(time, p.max_length_sec), 1, 0).astype(np.int32)
    codes = np.stack([bad_z, tilt, ep], axis=-1)
    done = np.any(codes != 0, axis=-1)                     # [U] ksim reduce
    success = done & np.all(codes != -1, axis=-1)
    return codes, done, success


# --------------------------------------------------------------------------------------
# Rewards R0..R12   train.py:125-506, table 1224-1256.   Trajectory-wise: time axis leading.
# --------------------------------------------------------------------------------------


def reward_initial_carry(batch_shape: tuple, dtype=np.float32) -> dict:
    """train.py:135-136, 175-178."""
    return {"t_single": np.zeros(batch_shape, dtype), "airtime": np.zeros(batch_shape + (2,), dtype),
            "prev_contact": np.ones(batch_shape + (2,), bool)}


def rewards(traj: dict, carry: dict, p: OracleParams) -> tuple[dict, np.ndarray, dict]:
    """All 12 terms on a trajectory.

    traj: xquat [T,...,24,4], xpos [T,...,24,3], qpos [T,...,27], qvel [T,...,26], ctrl [T,...,20],
          command [T,...,16], touch_l [T,...], touch_r [T,...], com_distance [T,...], done [T,...] bool.
    Returns (components dict name->[T,...], total [T,...] = sum_k scale_k r_k in table order, new carry).
    [U]: ksim's aggregation (no dt scaling, no clipping) is unverified.
    """
    cmd = traj["command"]
    dt = cmd.dtype
    T = cmd.shape[0]
    zc = zero_cmd_mask(cmd)
    one, zero = _c(cmd, 1.0), _c(cmd, 0.0)
    bq = traj["xquat"][..., BODY_BASE, :]
    r: dict[str, np.ndarray] = {}

    # R1 linvel  train.py:274-292
    e = quat_to_euler(bq, p.eps_quat).copy()
    e[..., :2] = 0
    qz = euler_to_quat(e)
    vcmd = np.concatenate([cmd[..., :2], np.zeros_like(cmd[..., :1])], axis=-1)
    g = rotate_vector_by_quat(vcmd, qz, inverse=False, eps=p.eps_quat)
    d = traj["qvel"][..., :2] - g[..., :2]
    verr = np.sqrt(np.sum(d * d, axis=-1))
    r["linvel"] = np.exp(-np.where(zc, verr, verr * verr) / _c(cmd, p.linvel_es))

    # R2 angvel  train.py:301-306
    r["angvel"] = np.exp(-np.abs(traj["qvel"][..., 5] - cmd[..., 2]) / _c(cmd, p.angvel_es))

    # R3 roll_pitch  train.py:316-334
    e = quat_to_euler(bq, p.eps_quat).copy()
    e[..., 2] = 0
    qxy = euler_to_quat(e)
    ce = np.stack([cmd[..., 4], cmd[..., 5], np.zeros_like(cmd[..., 5])], axis=-1)
    qc = euler_to_quat(ce)
    qerr = one - np.sum(qc * qxy, axis=-1) ** 2
    r["roll_pitch"] = np.exp(-qerr / np.where(zc, _c(cmd, p.rp_es_zero), _c(cmd, p.rp_es)))

    # R4 base_height  train.py:377-388
    zl = traj["xpos"][..., BODY_LFOOT, 2] - _c(cmd, p.bh_foot_origin)
    zr = traj["xpos"][..., BODY_RFOOT, 2] - _c(cmd, p.bh_foot_origin)
    h = traj["xpos"][..., BODY_BASE, 2] - np.minimum(zl, zr)
    r["base_height"] = np.exp(-np.abs(h - (cmd[..., 3] + _c(cmd, p.bh_standard))) / _c(cmd, p.bh_es))

    # R5 arm_pos  train.py:261-265  (xax.get_norm(., "l2") = elementwise square [U])
    da = traj["qpos"][..., 17:27] - (cmd[..., 6:16] + joint_biases(dt)[10:20])
    r["arm_pos"] = np.exp(-np.sum(da * da, axis=-1) / _c(cmd, p.arm_es))

    # contacts  train.py:139-141
    cl = traj["touch_l"] > _c(cmd, 0.1)
    cr = traj["touch_r"] > _c(cmd, 0.1)
    done = traj["done"]

    # R6 single_contact (stateful scan)  train.py:138-154
    single = np.logical_xor(cl, cr)
    t_sc = carry["t_single"].astype(dt)
    r6 = np.empty_like(cmd[..., 0])
    for t in range(T):
        t_sc = np.where(single[t], zero, t_sc + _c(cmd, p.ctrl_dt))
        t_sc = np.where(zc[t], _c(cmd, p.grace_period), t_sc)
        r6[t] = np.where(zc[t], one, (t_sc < _c(cmd, p.grace_period)).astype(dt))
    r["single_contact"] = r6

    # R7 no_contact_p  train.py:161-165
    r["no_contact_p"] = np.where(zc, zero, np.where(cl | cr, zero, one))

    # R8 feet_airtime (stateful scan)  train.py:197-213
    contact = np.stack([cl, cr], axis=-1)
    air = carry["airtime"].astype(dt)
    prev_c = carry["prev_contact"]
    r8 = np.empty_like(cmd[..., 0])
    for t in range(T):
        first = contact[t] & ~prev_c & ~done[t][..., None]
        r8[t] = np.sum((air - _c(cmd, p.touchdown_penalty)) * first.astype(dt), axis=-1)   # airtime of step t-1
        air = np.where(contact[t] | done[t][..., None], zero, air + _c(cmd, p.ctrl_dt))
        prev_c = contact[t]
    r["feet_airtime"] = np.where(zc, zero, r8)

    # R9 feet_orient  train.py:418-457
    yaw = quat_to_euler(bq, p.eps_quat)[..., 2]
    pi = _c(cmd, np.pi)
    tgt = np.stack([np.stack([np.full_like(yaw, -np.pi / 2), np.zeros_like(yaw), yaw - pi], axis=-1),
                    np.stack([np.full_like(yaw, np.pi / 2), np.zeros_like(yaw), yaw - pi], axis=-1)], axis=-2)
    fq = np.stack([traj["xquat"][..., BODY_LFOOT, :], traj["xquat"][..., BODY_RFOOT, :]], axis=-2)
    tq = euler_to_quat(tgt)
    rpy_err = np.sum(one - np.sum(tq * fq, axis=-1) ** 2, axis=-1)
    fe = quat_to_euler(fq, p.eps_quat).copy()
    fe[..., 2] = 0
    fq0 = euler_to_quat(fe)
    tgt0 = tgt.copy()
    tgt0[..., 2] = 0
    tq0 = euler_to_quat(tgt0)
    rp_err = np.sum(one - np.sum(tq0 * fq0, axis=-1) ** 2, axis=-1)
    # train.py:455 AS WRITTEN: cmd[:, 2] has shape (T,) so norm(axis=-1) reduces over TIME:
    # is_rotating is one scalar per trajectory, sqrt(sum_t wz_t^2) > 1e-3.
    wz = cmd[..., 2]
    is_rot = np.sqrt(np.sum(wz * wz, axis=0)) > _c(cmd, 1e-3)
    r["feet_orient"] = np.exp(-np.where(is_rot[None], rp_err, rpy_err) / _c(cmd, p.feet_es))

    # R10 com_distance  train.py:466-478
    cd = traj["com_distance"]
    r["com_distance"] = np.where(cd >= zero, np.where(zc, np.exp(-cd / _c(cmd, p.com_es)), zero), zero)

    # R11 base_accel  train.py:487-494  (edge pad: delta_0 = 0 every rollout)
    bv = traj["qvel"][..., :6]
    bvp = np.concatenate([bv[:1], bv], axis=0)
    dp = np.concatenate([done[:1], done], axis=0)
    acc = np.where(dp[:-1][..., None], zero, bvp[1:] - bvp[:-1])
    r["base_accel"] = np.exp(-np.sum(np.abs(acc), axis=-1) / _c(cmd, p.acc_es))

    # R12 torque  train.py:503-506
    r["torque"] = np.where(zc, np.mean(np.exp(-np.abs(traj["ctrl"]) / _c(cmd, p.torque_es)), axis=-1), one)

    total = np.zeros_like(cmd[..., 0])
    for name, s in zip(REWARD_NAMES, p.reward_scales):
        total = total + _c(cmd, s) * r[name].astype(dt)
    new_carry = {"t_single": t_sc, "airtime": air, "prev_contact": contact[-1]}
    return r, total, new_carry


# --------------------------------------------------------------------------------------
# GAE G1 [U] ksim.compute_ppo_inputs
# --------------------------------------------------------------------------------------


def compute_ppo_inputs(values, rewards_t, done, success, p: OracleParams):
    """values/rewards [T,...] float, done/success [T,...] bool -> (advantages, value_targets) [T,...].

    v+_t = v_{t+1} (v+_{T-1} = v_{T-1}); mask = 1-done; r~ = r + gamma v success (truncation bootstrap);
    delta = r~ + gamma v+ mask - v; A_t = delta_t + gamma lam mask_t A_{t+1}; targets = A + v.
    """
    dt = values.dtype
    T = values.shape[0]
    gam, lam = _c(values, p.gamma), _c(values, p.lam)
    vnext = np.concatenate([values[1:], values[-1:]], axis=0)
    mask = np.where(done, _c(values, 0.0), _c(values, 1.0))
    rt = rewards_t + gam * values * success.astype(dt)
    delta = rt + gam * vnext * mask - values
    adv = np.empty_like(values)
    a = np.zeros_like(values[0])
    for t in range(T - 1, -1, -1):
        a = delta[t] + gam * lam * mask[t] * a
        adv[t] = a
    targets = adv + values
    if p.normalize_advantages == 1:
        adv = adv / (np.std(adv, axis=0, keepdims=True) + _c(values, p.adv_eps))
    return adv, targets


# --------------------------------------------------------------------------------------
# The rollout control step on recorded state, stage order of ksim's step_engine around mjx.step
# --------------------------------------------------------------------------------------


def rollout_control_steps(w_actor, w_critic, state: dict, noise: dict, episode: dict, cmd_rand: dict, cmd0, carry: dict,
                          pg_carry, p: OracleParams, argmax: bool = False, with_critic: bool = True) -> dict:
    """T control steps over recorded state [T, N, ...]:  terminations -> observations -> sample_action ->
    PD torque -> (critic value) -> command update, carries reset to initial where done (train.py:1502-1506,1526).

    state/noise: dicts of [T, N, ...]; episode: dict of [N, ...]; cmd_rand: mode/u6/u_arms/u_switch [T, N, ...];
    cmd0 [N,16]; carry: {"actor","critic" [N,depth,2,H], "lpf_params" [N,20]}; pg_carry [N,3].
    Returns per-step stacks and the final carries.
    """
    T = state["qpos"].shape[0]
    ac, cc, lpf = carry["actor"], carry["critic"], carry["lpf_params"]
    cmd = cmd0
    out = {k: [] for k in ("actor_obs", "action", "log_prob", "ctrl", "codes", "done", "success", "value", "command",
                           "mean", "std")}
    prev_done = None
    for t in range(T):
        st = {k: v[t] for k, v in state.items()}
        nz = {k: v[t] for k, v in noise.items()}
        codes, done, success = terminations(st["xpos"], st["qpos"][..., 3:7], st["time"], p)
        o, pg_carry = get_observations(st, nz, episode, pg_carry, p, reset=prev_done)
        a_obs = actor_obs_from_dict(o, cmd)
        act, mean, std, ac, lpf = sample_action(w_actor, a_obs, ac, lpf, nz["eps_action"], argmax, p)
        logp = mvn_log_prob(mean, std, act)
        ctrl = position_actuator_torque(act, st["qpos"][..., 7:], st["qvel"][..., 6:], episode.get("kp"),
                                        episode.get("kd"), episode.get("tau_limit"), episode.get("action_bias"),
                                        episode.get("torque_bias"))
        out["command"].append(cmd)
        if with_critic:
            value, cc = critic_forward(w_critic, critic_obs_from_dict(o, cmd), cc)
            out["value"].append(value[..., 0])
            cc = np.where(done[..., None, None, None], np.zeros_like(cc), cc)
        ac = np.where(done[..., None, None, None], np.zeros_like(ac), ac)
        lpf = np.where(done[..., None], np.zeros_like(lpf), lpf)
        # a finished episode draws initial_command for the next one; otherwise the per-step switch law
        u_sw = np.where(done, np.asarray(-1.0, cmd.dtype), cmd_rand["u_switch"][t])
        cmd = command_step(cmd, u_sw, cmd_rand["mode"][t], cmd_rand["u6"][t], cmd_rand["u_arms"][t], p)
        for k, v in (("actor_obs", a_obs), ("action", act), ("log_prob", logp), ("ctrl", ctrl), ("codes", codes),
                     ("done", done), ("success", success), ("mean", mean), ("std", std)):
            out[k].append(v)
        prev_done = done
    res = {k: np.stack(v, axis=0) for k, v in out.items() if v}
    res["command_next"] = cmd
    res["carry"] = {"actor": ac, "critic": cc, "lpf_params": lpf}
    res["pg_carry"] = pg_carry
    return res
