"""Synthetic kbot-headless state batches (SURVEY 8d distributions) and layout converters.

`make_batch` draws AoS numpy arrays `[T, N, ...]` (the shape the reference's single-env code sees under vmap);
`to_soa` turns them into the library's `[T, F, ld]` SoA tensors.  `make_batch_device` draws the same
distributions directly on the GPU for workloads too large to stage through the host (bench configs 2/3).
"""

from __future__ import annotations

import math

import numpy as np

from . import spec

_THRESH_MARGIN = 1e-3


def _away(x: np.ndarray, thr: float, margin: float = _THRESH_MARGIN) -> np.ndarray:
    """Push samples that fall within `margin` of a comparison threshold to the far side (keeps parity tests of
    bit-exact masks meaningful: fp32 libm differences must not flip a comparison)."""
    near = np.abs(x - thr) < margin
    return np.where(near, thr + np.sign(x - thr + 1e-12) * 2 * margin, x).astype(x.dtype)


def make_commands(rng: np.random.Generator, shape: tuple) -> dict:
    """Raw randomness of the command law (train.py:724-785)."""
    return {"mode": rng.integers(0, 6, size=shape).astype(np.int32),
            "u6": rng.random(size=shape + (6,), dtype=np.float32),
            "u_arms": rng.random(size=shape + (10,), dtype=np.float32),
            "u_switch": rng.random(size=shape, dtype=np.float32)}


def make_batch(seed: int, T: int, N: int) -> dict:
    """State, noise, episode randomisation and command randomness for T steps of N envs (fp32, AoS)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    f = np.float32
    bias = np.array(spec.JOINT_BIASES, f)
    lim = np.array(spec.JOINT_LIMITS, f)

    def nrm(*s):
        return rng.standard_normal(size=s, dtype=f)

    def uni(lo, hi, *s):
        return (lo + (hi - lo) * rng.random(size=s, dtype=f)).astype(f)

    def unit_quat(*s):
        q = nrm(*s, 4)
        return (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(f)

    st = {}
    qpos = np.zeros((T, N, 27), f)
    qpos[..., 0:2] = uni(-1, 1, T, N, 2)
    qpos[..., 2] = 0.8 + 0.05 * nrm(T, N)
    bq = np.array([1, 0, 0, 0], f) + 0.15 * nrm(T, N, 4)
    qpos[..., 3:7] = bq / np.linalg.norm(bq, axis=-1, keepdims=True)
    qpos[..., 7:] = np.clip(bias + uni(-0.3, 0.3, T, N, 20), lim[:, 0], lim[:, 1])
    st["qpos"] = qpos
    qvel = np.concatenate([0.5 * nrm(T, N, 6), 2.0 * nrm(T, N, 20)], axis=-1)
    st["qvel"] = qvel
    sd = nrm(T, N, 49)
    sd[..., spec.SD_IMU_QUAT:spec.SD_IMU_QUAT + 4] = unit_quat(T, N)
    touch = np.where(rng.random(size=(T, N, 2)) < 0.5, 0.0, uni(1, 50, T, N, 2)).astype(f)
    sd[..., spec.SD_TOUCH_L] = touch[..., 0]
    sd[..., spec.SD_TOUCH_R] = touch[..., 1]
    st["sensordata"] = sd
    xpos = uni(-1, 1, T, N, 24, 3)
    xpos[..., spec.BODY_BASE, 2] = qpos[..., 2] - np.where(rng.random(size=(T, N)) < 0.03, 0.45, 0.0).astype(f)
    xpos[..., spec.BODY_LFOOT, 2] = uni(0, 0.15, T, N)
    xpos[..., spec.BODY_RFOOT, 2] = uni(0, 0.15, T, N)
    st["xpos"] = xpos
    xquat = unit_quat(T, N, 24)
    xquat[..., spec.BODY_BASE, :] = qpos[..., 3:7]
    st["xquat"] = xquat
    st["cinert"] = uni(0, 1, T, N, 24, 10)
    st["cvel"] = nrm(T, N, 24, 6)
    st["actuator_force"] = 10.0 * nrm(T, N, 20)
    st["com_distance"] = np.where(rng.random(size=(T, N)) < 0.5, -1.0, uni(0, 0.3, T, N)).astype(f)
    st["time"] = uni(0, 13, T, N)

    # keep comparison thresholds clear of fp32 libm noise (SURVEY 8c "Tolerances")
    h = xpos[..., spec.BODY_BASE, 2] - np.minimum(xpos[..., spec.BODY_LFOOT, 2], xpos[..., spec.BODY_RFOOT, 2])
    fix = np.abs(h - 0.4) < _THRESH_MARGIN
    xpos[..., spec.BODY_BASE, 2] = np.where(fix, xpos[..., spec.BODY_BASE, 2] + 4 * _THRESH_MARGIN,
                                            xpos[..., spec.BODY_BASE, 2])
    st["time"] = _away(st["time"], 12.0)
    q = qpos[..., 3:7]
    tilt = np.arccos(np.clip(1 - 2 * (q[..., 1] ** 2 + q[..., 2] ** 2), -1, 1))
    bad = np.abs(tilt - math.radians(45)) < _THRESH_MARGIN
    qpos[bad, 3:7] = np.array([1, 0, 0, 0], f)
    xquat[..., spec.BODY_BASE, :] = qpos[..., 3:7]

    noise = {"eps_jpos": uni(-1, 1, T, N, 20), "eps_jvel": uni(-1, 1, T, N, 20), "eps_gyro": nrm(T, N, 3),
             "eps_pg": nrm(T, N, 3), "eps_action": nrm(T, N, 20)}
    episode = {"jpos_bias": uni(-math.radians(3), math.radians(3), N, 20), "pg_lag": uni(0, 0.75, N),
               "pg_bias": uni(-math.radians(4), math.radians(4), N, 3),
               "kp": np.array(spec.KP, f) * uni(1 / 1.4, 1.4, N, 20),
               "kd": np.array(spec.KD, f) * uni(1 / 1.4, 1.4, N, 20),
               "tau_limit": np.array(spec.CTRL_LIMIT, f) * uni(0.5, 1.0, N, 20),
               "action_bias": uni(-0.02, 0.02, N, 20), "torque_bias": np.zeros((N, 20), f)}
    cmd_rand = make_commands(rng, (T, N))
    cmd0 = make_commands(rng, (N,))
    return {"state": st, "noise": noise, "episode": episode, "cmd_rand": cmd_rand, "cmd0_rand": cmd0,
            "rng": rng}


def to_soa(x: np.ndarray, n_batch_axes: int, device=None):
    """AoS numpy [*batch, N, *feat] -> SoA torch [*batch, F, ld] (F = prod(feat), ld = round_up4(N))."""
    import torch

    lead = x.shape[:n_batch_axes]
    n = x.shape[n_batch_axes]
    feat = int(np.prod(x.shape[n_batch_axes + 1:], dtype=np.int64)) if x.ndim > n_batch_axes + 1 else 1
    ld = (n + 3) // 4 * 4
    y = x.reshape(lead + (n, feat))
    y = np.moveaxis(y, -1, -2)
    out = np.zeros(lead + (feat, ld), dtype=x.dtype)
    out[..., :n] = y
    if x.ndim == n_batch_axes + 1:
        out = out.reshape(lead + (ld,))
    t = torch.from_numpy(np.ascontiguousarray(out))
    return t.to(device) if device is not None else t


def from_soa(t, n: int, feat_shape: tuple = ()) -> np.ndarray:
    """SoA torch [*batch, F, ld] (or [*batch, ld]) -> AoS numpy [*batch, n, *feat_shape]."""
    a = t.detach().cpu().numpy()
    if not feat_shape:
        return a[..., :n] if a.ndim >= 1 else a
    a = np.moveaxis(a[..., :n], -1, -2)
    return a.reshape(a.shape[:-1] + tuple(feat_shape))


def state_to_soa(st: dict, device, n_batch_axes: int = 1) -> dict:
    return {k: to_soa(v, n_batch_axes, device) for k, v in st.items()}


def make_batch_device(seed: int, T: int, N: int, device) -> dict:
    """Same distributions as make_batch, drawn on the GPU straight into [T, F, ld] SoA (no host staging)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ld = (N + 3) // 4 * 4
    f32 = torch.float32

    def nrm(*s):
        return torch.randn(*s, generator=g, device=device, dtype=f32)

    def uni(lo, hi, *s):
        return lo + (hi - lo) * torch.rand(*s, generator=g, device=device, dtype=f32)

    bias = torch.tensor(spec.JOINT_BIASES, device=device, dtype=f32)[None, :, None]
    lim = torch.tensor(spec.JOINT_LIMITS, device=device, dtype=f32)
    st = {}
    qpos = torch.empty((T, 27, ld), device=device, dtype=f32)
    qpos[:, 0:2] = uni(-1, 1, T, 2, ld)
    qpos[:, 2] = 0.8 + 0.05 * nrm(T, ld)
    bq = 0.15 * nrm(T, 4, ld)
    bq[:, 0] += 1.0
    qpos[:, 3:7] = bq / bq.norm(dim=1, keepdim=True)
    qpos[:, 7:] = torch.minimum(torch.maximum(bias + uni(-0.3, 0.3, T, 20, ld), lim[None, :, 0, None]),
                                lim[None, :, 1, None])
    st["qpos"] = qpos
    st["qvel"] = torch.cat([0.5 * nrm(T, 6, ld), 2.0 * nrm(T, 20, ld)], dim=1)
    sd = nrm(T, 49, ld)
    iq = nrm(T, 4, ld)
    sd[:, spec.SD_IMU_QUAT:spec.SD_IMU_QUAT + 4] = iq / iq.norm(dim=1, keepdim=True)
    touch = torch.where(torch.rand(T, 2, ld, generator=g, device=device) < 0.5, torch.zeros((), device=device),
                        uni(1, 50, T, 2, ld))
    sd[:, spec.SD_TOUCH_L] = touch[:, 0]
    sd[:, spec.SD_TOUCH_R] = touch[:, 1]
    st["sensordata"] = sd
    xpos = uni(-1, 1, T, 72, ld)
    xpos[:, 3 * spec.BODY_BASE + 2] = qpos[:, 2]
    xpos[:, 3 * spec.BODY_LFOOT + 2] = uni(0, 0.15, T, ld)
    xpos[:, 3 * spec.BODY_RFOOT + 2] = uni(0, 0.15, T, ld)
    st["xpos"] = xpos
    xq = nrm(T, 24, 4, ld)
    xq = xq / xq.norm(dim=2, keepdim=True)
    xq[:, spec.BODY_BASE] = qpos[:, 3:7]
    st["xquat"] = xq.reshape(T, 96, ld).contiguous()
    st["cinert"] = uni(0, 1, T, 240, ld)
    st["cvel"] = nrm(T, 144, ld)
    st["actuator_force"] = 10.0 * nrm(T, 20, ld)
    st["com_distance"] = torch.where(torch.rand(T, ld, generator=g, device=device) < 0.5,
                                     -torch.ones((), device=device), uni(0, 0.3, T, ld))
    st["time"] = uni(0, 13, T, ld)
    noise = {"eps_jpos": uni(-1, 1, T, 20, ld), "eps_jvel": uni(-1, 1, T, 20, ld), "eps_gyro": nrm(T, 3, ld),
             "eps_pg": nrm(T, 3, ld)}
    kp = torch.tensor(spec.KP, device=device, dtype=f32)[:, None]
    kd = torch.tensor(spec.KD, device=device, dtype=f32)[:, None]
    tl = torch.tensor(spec.CTRL_LIMIT, device=device, dtype=f32)[:, None]
    episode = {"jpos_bias": uni(-math.radians(3), math.radians(3), 20, ld), "pg_lag": uni(0, 0.75, ld),
               "pg_bias": uni(-math.radians(4), math.radians(4), 3, ld), "kp": kp * uni(1 / 1.4, 1.4, 20, ld),
               "kd": kd * uni(1 / 1.4, 1.4, 20, ld), "tau_limit": tl * uni(0.5, 1.0, 20, ld),
               "action_bias": uni(-0.02, 0.02, 20, ld), "torque_bias": torch.zeros(20, ld, device=device)}
    out = {"state": st, "noise": noise, "episode": episode,
           "eps_action": nrm(T, 20, ld), "u_switch": torch.rand(T, ld, generator=g, device=device),
           "cmd_mode": torch.randint(0, 6, (T, ld), generator=g, device=device, dtype=torch.int32),
           "cmd_u6": torch.rand(T, 6, ld, generator=g, device=device),
           "cmd_u_arms": torch.rand(T, 10, ld, generator=g, device=device)}
    return out


def make_weights(seed: int, num_in: int, num_out: int, hidden: int, depth: int) -> dict:
    """Random eqx-layout weights, U(+-1/sqrt(fan_in)) (eqx init law; values synthetic -- no checkpoints)."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def u(shape, fan_in):
        lim = 1.0 / math.sqrt(fan_in)
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    w = {"w_in": u((hidden, num_in), num_in), "b_in": u((hidden,), num_in),
         "w_out": u((num_out, hidden), hidden), "b_out": u((num_out,), hidden), "layers": []}
    for _ in range(depth):
        w["layers"].append({"w_ih": u((4 * hidden, hidden), hidden), "w_hh": u((4 * hidden, hidden), hidden),
                            "b": u((4 * hidden,), hidden)})
    return w


def weights_to_device(w: dict, device) -> dict:
    import torch

    out = {k: torch.from_numpy(v).to(device) for k, v in w.items() if k != "layers"}
    out["layers"] = [{k: torch.from_numpy(v).to(device) for k, v in lw.items()} for lw in w["layers"]]
    return out
