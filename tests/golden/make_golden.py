"""Generates tests/golden/*.npz -- known-answer vectors that pin the oracle's [U] helper restatements.

The reference (kscalelabs/kbot-joystick) ships no tests or fixtures and JAX/ksim/xax/equinox/distrax are not
installable in this image, so the reference itself cannot produce goldens ("parity unpinned").
The next best pin is INDEPENDENT implementations of the same published algorithms that ARE available here:

  lstm.npz      torch.nn.LSTMCell (bias_hh = 0), gate order i,f,g,o == equinox.nn.LSTMCell 0.12.2
  mvn.npz       torch.distributions.Independent(Normal) log_prob / entropy == distrax.MultivariateNormalDiag
  quat.npz      scipy.spatial.transform.Rotation == xax.quat_to_euler / euler_to_quat / rotate_vector_by_quat
  softplus.npz  torch.nn.functional.softplus in float64 == jax.nn.softplus
  gae.npz       a 4-step example worked by hand in exact rational arithmetic (fractions.Fraction)
  structure.npz structural known answers asserted by train.py itself (sizes, FLOP counts)

Run:  python tests/golden/make_golden.py      (CPU only; writes next to this file)
"""

from fractions import Fraction
from pathlib import Path

import numpy as np
import torch
from scipy.spatial.transform import Rotation

OUT = Path(__file__).resolve().parent


def lstm():
    torch.manual_seed(1234)
    H, B = 32, 5
    cell = torch.nn.LSTMCell(H, H, bias=True, dtype=torch.float64)
    with torch.no_grad():
        cell.bias_hh.zero_()
    x, h, c = (torch.randn(B, H, dtype=torch.float64) for _ in range(3))
    h2, c2 = cell(x, (h, c))
    np.savez(OUT / "lstm.npz", w_ih=cell.weight_ih.detach().numpy(), w_hh=cell.weight_hh.detach().numpy(),
             b=cell.bias_ih.detach().numpy(), x=x.numpy(), h=h.numpy(), c=c.numpy(), h2=h2.detach().numpy(),
             c2=c2.detach().numpy())


def mvn():
    torch.manual_seed(5)
    mean = torch.randn(7, 20, dtype=torch.float64)
    std = torch.rand(7, 20, dtype=torch.float64) * 0.9 + 0.05
    a = mean + std * torch.randn(7, 20, dtype=torch.float64)
    d = torch.distributions.Independent(torch.distributions.Normal(mean, std), 1)
    np.savez(OUT / "mvn.npz", mean=mean.numpy(), std=std.numpy(), a=a.numpy(), log_prob=d.log_prob(a).numpy(),
             entropy=d.entropy().numpy())


def quat():
    rng = np.random.default_rng(7)
    q_wxyz = rng.standard_normal((64, 4))
    q_wxyz /= np.linalg.norm(q_wxyz, axis=-1, keepdims=True)
    rot = Rotation.from_quat(q_wxyz[:, [1, 2, 3, 0]])  # scipy is scalar-last
    euler_rpy = rot.as_euler("xyz")                     # extrinsic x,y,z == R = Rz(yaw) Ry(pitch) Rx(roll)
    v = rng.standard_normal((64, 3))
    e = rng.uniform(-1.5, 1.5, size=(64, 3))
    q_from_e = Rotation.from_euler("xyz", e).as_quat()[:, [3, 0, 1, 2]]
    q_from_e *= np.sign(q_from_e[:, :1])                # canonical sign w >= 0 (|roll|,|pitch|,|yaw| < pi)
    np.savez(OUT / "quat.npz", q=q_wxyz, euler=euler_rpy, v=v, v_rot=rot.apply(v), v_rot_inv=rot.inv().apply(v),
             e=e, q_from_e=q_from_e, g_body=rot.inv().apply(np.array([0, 0, -9.81])))


def softplus():
    x = torch.linspace(-30, 30, 121, dtype=torch.float64)
    np.savez(OUT / "softplus.npz", x=x.numpy(), y=torch.nn.functional.softplus(x, threshold=1e9).numpy())


def gae():
    # gamma = lam = 1/2 so every quantity is an exact dyadic rational.
    g = lam = Fraction(1, 2)
    v = [Fraction(1), Fraction(2), Fraction(-1), Fraction(3)]
    r = [Fraction(1, 2), Fraction(1), Fraction(0), Fraction(2)]
    done = [0, 1, 0, 0]
    succ = [0, 1, 0, 0]           # step 1 ends by time-out: bootstrap with gamma * v
    T = 4
    vn = v[1:] + v[-1:]
    a = Fraction(0)
    adv = [None] * T
    for t in reversed(range(T)):
        mask = 1 - done[t]
        delta = (r[t] + g * v[t] * succ[t]) + g * vn[t] * mask - v[t]
        a = delta + g * lam * mask * a
        adv[t] = a
    np.savez(OUT / "gae.npz", gamma=float(g), lam=float(lam), values=np.array([float(x) for x in v]),
             rewards=np.array([float(x) for x in r]), done=np.array(done, bool), success=np.array(succ, bool),
             adv=np.array([float(x) for x in adv]), targets=np.array([float(x + y) for x, y in zip(adv, v)]))


def structure():
    # train.py:1279-1312 input-size arithmetic, convert.py:71 carry size, SURVEY 8d FLOP counts
    np.savez(OUT / "structure.npz", actor_obs=65, critic_obs=475, carry_size_h256=2 * 2 * 256 + 20,
             actor_flops_h256=2150912, critic_flops_h256=2340864, actor_params_h256=1077800,
             critic_params_h256=1172737)


if __name__ == "__main__":
    lstm(); mvn(); quat(); softplus(); gae(); structure()
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))
