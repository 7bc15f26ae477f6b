"""Shared helpers of the GPU parity tests: one seeded batch presented to both sides -- AoS numpy for the oracle,
SoA CUDA tensors for the library -- and the comparison rules.

Tolerances (BASELINE.json north_star): bit-exact for termination codes / done / success / indices; 1e-5 relative
for fp32 outputs.  A pure relative bound is meaningless at zero crossings, so every float comparison is
|gpu - oracle| <= RTOL * |oracle| + ATOL with ATOL = 1e-6 x the natural scale of the quantity (stated per call).
"""

from __future__ import annotations

import numpy as np
import torch

import kbot_oracle as O
from kbot_joystick_b200 import _lib as L
from kbot_joystick_b200 import synth
from kbot_joystick_b200.engine import KbotStep

RTOL = 1e-5
P = O.OracleParams()


def close(gpu, ref, name, rtol=RTOL, atol=1e-6):
    gpu = np.asarray(gpu)
    ref = np.asarray(ref)
    assert gpu.shape == ref.shape, f"{name}: shape {gpu.shape} vs {ref.shape}"
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64))
    bound = rtol * np.abs(ref.astype(np.float64)) + atol
    bad = err > bound
    if bad.any():
        i = np.unravel_index(np.argmax(err - bound), err.shape)
        raise AssertionError(f"{name}: {bad.sum()}/{bad.size} outside rtol={rtol} atol={atol}; worst at {i}: "
                             f"gpu={gpu[i]!r} oracle={ref[i]!r} err={err[i]:.3e}")


def exact(gpu, ref, name):
    gpu = np.asarray(gpu)
    ref = np.asarray(ref)
    assert gpu.shape == ref.shape, f"{name}: shape {gpu.shape} vs {ref.shape}"
    if not np.array_equal(gpu, ref):
        bad = gpu != ref
        i = np.unravel_index(np.argmax(bad), bad.shape)
        raise AssertionError(f"{name}: {bad.sum()}/{bad.size} differ (bit-exact required); first at {i}: "
                             f"gpu={gpu[i]!r} oracle={ref[i]!r}")


class Batch:
    """T steps x N envs of synthetic state in both representations."""

    def __init__(self, seed: int, T: int, N: int, device):
        self.T, self.N, self.dev = T, N, device
        self.ld = (N + 3) // 4 * 4
        self.np = synth.make_batch(seed, T, N)
        b = self.np
        self.state = {k: synth.to_soa(v, 1, device) for k, v in b["state"].items()}
        self.noise = {k: synth.to_soa(v, 1, device) for k, v in b["noise"].items()}
        self.episode = {k: synth.to_soa(v, 0, device) for k, v in b["episode"].items()}
        self.cmd_rand = {k: synth.to_soa(v, 1, device) for k, v in b["cmd_rand"].items()}
        r0 = b["cmd0_rand"]
        self.cmd0_np = O.initial_command(r0["mode"], r0["u6"], r0["u_arms"], P)
        self.cmd0 = synth.to_soa(self.cmd0_np, 0, device)

    def state_at(self, t):
        return {k: v[t] for k, v in self.state.items()}

    def noise_at(self, t):
        return {k: v[t] for k, v in self.noise.items() if k != "eps_action"}

    def np_state_at(self, t):
        return {k: v[t] for k, v in self.np["state"].items()}

    def np_noise_at(self, t):
        return {k: v[t] for k, v in self.np["noise"].items()}


def make_engine(hidden=256, depth=2, gemm_path=L.GEMM_SIMT_FP32, seed=77, device=None, **overrides):
    """Engine + eqx-layout weights (numpy for the oracle, packed on the device for the library)."""
    eng = KbotStep(hidden_size=hidden, depth=depth, gemm_path=gemm_path, overrides=overrides)
    wa = synth.make_weights(seed, 65, 40, hidden, depth)
    wc = synth.make_weights(seed + 1, 475, 1, hidden, depth)
    eng.pack_weights(L.NET_ACTOR, synth.weights_to_device(wa, device))
    eng.pack_weights(L.NET_CRITIC, synth.weights_to_device(wc, device))
    return eng, wa, wc


def carry_to_np(c: torch.Tensor, n: int) -> np.ndarray:
    """library AoS [depth, 2, n, H] -> oracle [n, depth, 2, H]."""
    return np.moveaxis(c.detach().cpu().numpy(), 2, 0)[:n]


def carry_from_np(c: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(np.moveaxis(c, 0, 2))).to(device)
