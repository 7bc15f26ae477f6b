"""GPU parity tests: every stage of the control step, called through the C-ABI of libkbotstep.so, against the CPU
oracle on the same seeded inputs.  Bit-exact for termination codes / done / success / contact flags; 1e-5 relative
(+ a stated absolute floor) for fp32 outputs.  Sizes cover BASELINE config 1 (N=32), ragged and tiny N."""

import ctypes as C

import numpy as np
import pytest
import torch

import harness as Hn
import kbot_oracle as O
from harness import Batch, P, close, exact
from kbot_joystick_b200 import _lib as L
from kbot_joystick_b200 import synth
from kbot_joystick_b200.engine import COMPUTED_OBS_ROWS

pytestmark = pytest.mark.gpu

S = synth.from_soa
PATHS = [pytest.param(L.GEMM_SIMT_FP32, id="simt"), pytest.param(L.GEMM_TC_3XTF32, id="tc3xtf32"),
         pytest.param(L.GEMM_TC_2XF16, id="tc2xf16")]


@pytest.fixture(scope="module")
def dev(cuda_device, lib_built):
    return cuda_device


@pytest.fixture(scope="module")
def eng(dev):
    e, wa, wc = Hn.make_engine(device=dev)
    yield e
    e.close()


@pytest.mark.parametrize("N", [1, 3, 32, 130, 1001])
def test_observations(eng, dev, N):
    b = Batch(100 + N, 2, N, dev)
    ld = b.ld
    pg = torch.zeros((3, ld), device=dev)
    pg_np = np.zeros((N, 3), np.float32)
    prev_done = None
    for t in range(2):
        comp = torch.zeros((78, ld), device=dev)
        aobs = torch.zeros((65, ld), device=dev)
        cobs = torch.zeros((475, ld), device=dev)
        reset = None
        if t == 1:
            prev_done = (np.arange(N) % 3 == 0)
            reset = synth.to_soa(prev_done.astype(np.uint8), 0, dev)
        eng.observations(b.state_at(t), b.cmd0, b.noise_at(t), b.episode, pg, comp, aobs, cobs, N, pg_reset=reset)
        o, pg_np = O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], pg_np, P, reset=prev_done)
        for name, (r0, r1) in COMPUTED_OBS_ROWS.items():
            close(S(comp[r0:r1], N, (r1 - r0,)), o[name], name, atol=1e-5 if "gravity" in name else 1e-6)
        close(S(aobs, N, (65,)), O.actor_obs_from_dict(o, b.cmd0_np), f"actor_obs t={t}")
        close(S(cobs, N, (475,)), O.critic_obs_from_dict(o, b.cmd0_np), f"critic_obs t={t}")
        close(S(pg, N, (3,)), pg_np, "pg_carry", atol=1e-5)
    # pure slices of the table are bit-exact copies
    exact(S(cobs[80:310], N, (230,)), o["center_of_mass_inertia"], "cinert dump")
    exact(S(cobs[310:448], N, (138,)), o["center_of_mass_velocity"], "cvel dump")
    exact(S(aobs[48], N), O.zero_cmd_mask(b.cmd0_np).astype(np.float32), "zero_cmd flag")


@pytest.mark.parametrize("T,N", [(3, 50), (2, 1001)])
def test_mirror_observations(eng, dev, T, N):
    """X1: mirror_obs / mirror_cmd / mirror_joints (train.py:1574-1756) + the concatenations on the mirrored
    observations, from what a rollout stores (raw observations incl. the noisy twins, state, commands)."""
    b = Batch(150 + N, T, N, dev)
    ld = b.ld
    comp = torch.zeros((T, 78, ld), device=dev)
    cmd_np = np.stack([b.cmd0_np * (1.0 + 0.25 * t) for t in range(T)]).astype(np.float32)
    cmd = synth.to_soa(cmd_np, 1, dev)
    ref_a, ref_c, ref_cmd = [], [], []
    for t in range(T):
        eng.observations(b.state_at(t), cmd[t], b.noise_at(t), b.episode, None, comp[t], None, None, N)
        o, _ = O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], None, P)
        mo, mc = O.mirror_obs(o), O.mirror_cmd(cmd_np[t])
        ref_a.append(O.actor_obs_from_dict(mo, mc))
        ref_c.append(O.critic_obs_from_dict(mo, mc))
        ref_cmd.append(mc)
    aobs = torch.full((T, 65, ld), float("nan"), device=dev)
    cobs = torch.full((T, 475, ld), float("nan"), device=dev)
    cout = torch.full((T, 16, ld), float("nan"), device=dev)
    eng.mirror_observations(b.state, comp, cmd, aobs, cobs, cout, n_envs=N)
    close(S(aobs, N, (65,)), np.stack(ref_a), "mirrored actor_obs", atol=1e-5)
    close(S(cobs, N, (475,)), np.stack(ref_c), "mirrored critic_obs", atol=1e-5)
    exact(S(cout, N, (16,)), np.stack(ref_cmd), "mirror_cmd")
    # pure sign/permutation rows are bit-exact
    exact(S(cobs[:, 80:448], N, (368,)), np.stack(ref_c)[..., 80:448], "mirrored cinert/cvel dump")
    j = torch.randn((T, 20, ld), device=dev)
    exact(S(eng.mirror_joints(j, n_envs=N), N, (20,)), O.mirror_joints(S(j, N, (20,))), "mirror_joints")
    # involution: mirroring twice is the identity (size-independent property)
    exact(S(eng.mirror_joints(eng.mirror_joints(j, n_envs=N), n_envs=N), N, (20,)), S(j, N, (20,)), "mirror_joints twice")


@pytest.mark.parametrize("ncon", [4, 8, 16])
def test_com_distance_observation(eng, dev, ncon):
    """O12 COMDistanceObservation (train.py:509-659): convex hull of the floor contacts (monotone chain) -> centroid ->
    distance to subtree_com[2]; the < 3 distinct geoms branch (-1) and the selection are exact, the distance fp32."""
    T, N = 3, 700
    rng = np.random.default_rng(50 + ncon)
    geom1 = rng.choice([0, 0, 0, 5], size=(T, N, ncon)).astype(np.int32)             # 0 = floor
    geom2 = rng.choice([-1, 30, 31, 32, 33, 34, 35], size=(T, N, ncon)).astype(np.int32)
    geom2[:, ::7] = 30                                                                # a block with < 3 distinct geoms
    geom2[:, ::7, 0] = 31
    pos = np.concatenate([rng.uniform(-0.3, 0.3, (T, N, ncon, 2)), rng.uniform(0, 0.02, (T, N, ncon, 1))], -1).astype(np.float32)
    pos[:, 1::11, :, 1] = pos[:, 1::11, :, 0] * 0.5                                   # collinear: degenerate hull, fallback mean
    pos[:, 2::13, 1:] = pos[:, 2::13, :1]                                             # all contacts at one point
    com = np.concatenate([rng.uniform(-0.2, 0.2, (T, N, 2)), np.full((T, N, 1), 0.8)], -1).astype(np.float32)
    ref = O.com_distance_observation(geom1, geom2, pos, com)
    ref64 = O.com_distance_observation(geom1, geom2, pos.astype(np.float64), com.astype(np.float64))
    out = eng.com_distance(synth.to_soa(geom1, 1, dev), synth.to_soa(geom2, 1, dev),
                           synth.to_soa(pos.reshape(T, N, 3 * ncon), 1, dev), synth.to_soa(com, 1, dev), n_envs=N)
    got = S(out, N)
    exact(got < 0, ref < 0, "fewer than 3 distinct contact geoms -> -1")
    assert (ref < 0).sum() > 0 and (ref >= 0).sum() > N
    # the centroid divides by the hull area: conditioning ~ |coords|^3 / area; bound the fp32 oracle by the fp64 one too
    tol = np.maximum(2e-5, 4 * np.abs(ref - ref64))
    err = np.abs(got - ref)
    assert (err <= tol + 1e-5 * np.abs(ref)).all(), (err.max(), np.argmax(err - tol))
    # size-independent property: with all-floor contacts, translating every contact and the COM together leaves the
    # distance unchanged (up to the conditioning of the centroid, bounded per env by the fp32-vs-fp64 oracle gap)
    g1 = np.zeros_like(geom1)
    shift = np.array([0.25, -0.5, 0.0], np.float32)
    pos_s, com_s = pos + shift, com + shift

    def gpu(p_, c_):
        return S(eng.com_distance(synth.to_soa(g1, 1, dev), synth.to_soa(geom2, 1, dev),
                                  synth.to_soa(p_.reshape(T, N, 3 * ncon), 1, dev), synth.to_soa(c_, 1, dev), n_envs=N), N)

    a, bsh = gpu(pos, com), gpu(pos_s, com_s)
    gap = (np.abs(O.com_distance_observation(g1, geom2, pos, com) -
                  O.com_distance_observation(g1, geom2, pos.astype(np.float64), com.astype(np.float64))) +
           np.abs(O.com_distance_observation(g1, geom2, pos_s, com_s) -
                  O.com_distance_observation(g1, geom2, pos_s.astype(np.float64), com_s.astype(np.float64))))
    regular = np.ones((T, N), bool)
    regular[:, 1::11] = False          # the collinear / coincident constructions switch between the area formula and the
    regular[:, 2::13] = False          # mean fallback under translation (|area| < 1e-12 is not translation invariant)
    ok = (a >= 0) & regular
    assert ok.sum() > N and (np.abs(a - bsh)[ok] <= 4 * gap[ok] + 5e-5).all()


def test_observations_without_noise_twins_equal_clean(eng, dev):
    b = Batch(7, 1, 64, dev)
    aobs = torch.zeros((65, b.ld), device=dev)
    cobs = torch.zeros((475, b.ld), device=dev)
    eng.observations(b.state_at(0), b.cmd0, None, None, None, None, aobs, cobs, 64)
    exact(aobs.cpu().numpy(), cobs[:65].cpu().numpy(), "noise-free actor block == critic block")


@pytest.mark.parametrize("N", [2, 32, 777])
def test_command_update(eng, dev, N):
    b = Batch(200 + N, 3, N, dev)
    cmd = b.cmd0.clone()
    ref = b.cmd0_np
    # initial_command (u_switch NULL) reproduces cmd0
    init = torch.zeros_like(cmd)
    r0 = {k: synth.to_soa(v, 0, dev) for k, v in b.np["cmd0_rand"].items()}
    eng.command_update(init, r0["mode"], r0["u6"], r0["u_arms"], None, N)
    exact(S(init, N, (16,)), ref, "initial_command")
    r = b.np["cmd_rand"]
    for t in range(3):
        us = (r["u_switch"][t] * 0.02).astype(np.float32)       # ~20 % of envs switch
        eng.command_update(cmd, b.cmd_rand["mode"][t], b.cmd_rand["u6"][t], b.cmd_rand["u_arms"][t],
                           synth.to_soa(us, 0, dev), N)
        ref = O.command_step(ref, us, r["mode"][t], r["u6"][t], r["u_arms"][t], P)
        exact(S(cmd, N, (16,)), ref, f"command t={t}")


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("hidden,N", [(256, 32), (256, 200), (128, 77), (256, 1)])
def test_actor_and_critic_step(dev, path, hidden, N):
    e, wa, wc = Hn.make_engine(hidden=hidden, gemm_path=path, device=dev)
    p = O.OracleParams(hidden_size=hidden)
    b = Batch(300 + N, 3, N, dev)
    ld = b.ld
    ac = torch.zeros((2, 2, N, hidden), device=dev)
    cc = torch.zeros((2, 2, N, hidden), device=dev)
    lpf = torch.zeros((20, ld), device=dev)
    ac_np = np.zeros((N, 2, 2, hidden), np.float32)
    cc_np = np.zeros_like(ac_np)
    lpf_np = np.zeros((N, 20), np.float32)
    rng = np.random.default_rng(9)
    for t in range(3):
        o, _ = O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], None, P)
        a_obs, c_obs = O.actor_obs_from_dict(o, b.cmd0_np), O.critic_obs_from_dict(o, b.cmd0_np)
        done_np = rng.random(N) < 0.3
        done = synth.to_soa(done_np.astype(np.uint8), 0, dev)
        eps = b.noise["eps_action"][t]
        out = e.actor_step(synth.to_soa(a_obs, 0, dev), ac, lpf, eps=eps, done=done, n_envs=N)
        act, mean, std, ac_np, lpf_np = O.sample_action(wa, a_obs, ac_np, lpf_np, b.np["noise"]["eps_action"][t], False, p)
        close(S(out["mean"], N, (20,)), mean, f"mean t={t}")
        close(S(out["std"], N, (20,)), std, f"std t={t}")
        close(S(out["action"], N, (20,)), act, f"action t={t}")
        close(S(out["log_prob"], N), O.mvn_log_prob(mean, std, act), f"log_prob t={t}", atol=1e-5)
        close(S(out["entropy"], N), O.mvn_entropy(std), f"entropy t={t}", atol=1e-5)
        ac_np = np.where(done_np[:, None, None, None], 0, ac_np).astype(np.float32)
        lpf_np = np.where(done_np[:, None], 0, lpf_np).astype(np.float32)
        close(Hn.carry_to_np(ac, N), ac_np, f"actor carry t={t}")
        close(S(lpf, N, (20,)), lpf_np, f"lpf t={t}")
        val = e.critic_step(synth.to_soa(c_obs, 0, dev), cc, done=done, n_envs=N)
        v_np, cc_np = O.critic_forward(wc, c_obs, cc_np)
        close(S(val, N), v_np[:, 0], f"value t={t}", atol=1e-5)       # see harness.compare_rollout on the floor
        cc_np = np.where(done_np[:, None, None, None], 0, cc_np).astype(np.float32)
        close(Hn.carry_to_np(cc, N), cc_np, f"critic carry t={t}", atol=1e-5)
    # stored-transition path of _ppo_scan_fn: log-prob of a given action, argmax mode
    a_in = (mean + 0.3 * rng.standard_normal(mean.shape)).astype(np.float32)
    ac2, lpf2 = ac.clone(), lpf.clone()
    out = e.actor_step(synth.to_soa(a_obs, 0, dev), ac2, lpf2, eps=None, action_in=synth.to_soa(a_in, 0, dev), n_envs=N)
    _, mean2, std2, _, _ = O.sample_action(wa, a_obs, ac_np, lpf_np, None, True, p)
    close(S(out["action"], N, (20,)), mean2, "mode()")
    close(S(out["log_prob"], N), O.mvn_log_prob(mean2, std2, a_in), "log_prob(action_in)", atol=1e-5)
    e.close()


@pytest.mark.parametrize("N", [5, 32, 515])
def test_torque(eng, dev, N):
    b = Batch(400 + N, 1, N, dev)
    act_np = b.np["state"]["qpos"][0][:, 7:] + 0.2 * b.np["noise"]["eps_action"][0]
    act = synth.to_soa(act_np, 0, dev)
    ep = b.np["episode"]
    q, qd = b.np["state"]["qpos"][0][:, 7:], b.np["state"]["qvel"][0][:, 6:]
    ctrl = eng.torque(act, b.state_at(0), b.episode, n_envs=N)
    ref = O.position_actuator_torque(act_np, q, qd, ep["kp"], ep["kd"], ep["tau_limit"], ep["action_bias"], ep["torque_bias"])
    close(S(ctrl, N, (20,)), ref, "ctrl (randomised gains)", atol=1e-4)
    assert (np.abs(ref) == ep["tau_limit"]).any(), "clip branch not exercised"
    ctrl = eng.torque(act, b.state_at(0), None, n_envs=N)
    close(S(ctrl, N, (20,)), O.position_actuator_torque(act_np, q, qd), "ctrl (nominal gains)", atol=1e-4)


@pytest.mark.parametrize("N", [5, 640])
def test_torque_substeps_latency_and_drop(eng, dev, N):
    """8f-2: five PD evaluations per control step on the latency / drop-selected action (train.py:1775-1781)."""
    S_ = 5
    rng = np.random.default_rng(300 + N)
    f = np.float32
    b = Batch(310 + N, 1, N, dev)
    ep = b.np["episode"]
    act = (0.5 * rng.normal(size=(N, 20))).astype(f)
    prev = (0.5 * rng.normal(size=(N, 20))).astype(f)
    u_drop = rng.random(N).astype(f)
    lat = rng.uniform(0.003, 0.01, N).astype(f)
    lat[:: 7] = 0.004                                                  # exactly on a sub-step boundary: k dt >= latency
    q = rng.normal(0, 0.5, (S_, N, 20)).astype(f)
    qd = rng.normal(0, 2.0, (S_, N, 20)).astype(f)
    ref, applied = O.position_actuator_substeps(act, prev, u_drop, lat, q, qd, kp=ep["kp"], kd=ep["kd"], tau_limit=ep["tau_limit"],
                                                action_bias=ep["action_bias"], torque_bias=ep["torque_bias"])
    prev_d = synth.to_soa(prev, 0, dev)
    ctrl = eng.torque_substeps(synth.to_soa(act, 0, dev), prev_d, synth.to_soa(u_drop, 0, dev), synth.to_soa(lat, 0, dev),
                               synth.to_soa(q, 1, dev), synth.to_soa(qd, 1, dev), b.episode, n_envs=N)
    close(S(ctrl, N, (20,)), ref, "sub-step torques", atol=1e-4)
    exact(S(prev_d, N, (20,)), applied, "applied action (drop = repeat the previous one)")
    assert (u_drop < 0.05).sum() >= 0 and ((np.arange(S_)[:, None] * f(0.004) >= lat[None]).sum(0) >= 2).all()


@pytest.mark.parametrize("N", [1, 32, 4099])
def test_terminations_bit_exact(eng, dev, N):
    b = Batch(500 + N, 1, N, dev)
    out = eng.terminate(b.state_at(0), N, want_pre=True)
    st = b.np_state_at(0)
    codes, done, succ = O.terminations(st["xpos"], st["qpos"][:, 3:7], st["time"], P)
    exact(S(out["codes"], N, (3,)), codes, "codes")
    exact(S(out["done"], N).astype(bool), done, "done")
    exact(S(out["success"], N).astype(bool), succ, "success")
    close(S(out["pre"][0], N), O.bad_z_height(st["xpos"]), "height")
    close(S(out["pre"][1], N), O.upright_tilt(st["qpos"][:, 3:7], P), "tilt", atol=1e-5)
    if N >= 32:
        assert done.any() and (~done).any()


def _traj(b, seed=11):
    """Recorded trajectory for the reward tests (command evolves by the command law; ~10 % done)."""
    rng = np.random.default_rng(seed)
    st = b.np["state"]
    T, N = b.T, b.N
    r = b.np["cmd_rand"]
    cmd = np.empty((T, N, 16), np.float32)
    c = b.cmd0_np
    for t in range(T):
        cmd[t] = c
        c = O.command_step(c, (r["u_switch"][t] * 0.05).astype(np.float32), r["mode"][t], r["u6"][t], r["u_arms"][t], P)
    return {"xquat": st["xquat"], "xpos": st["xpos"], "qpos": st["qpos"], "qvel": st["qvel"],
            "ctrl": (20 * rng.standard_normal((T, N, 20))).astype(np.float32), "command": cmd,
            "touch_l": st["sensordata"][..., O.SD_TOUCH_L], "touch_r": st["sensordata"][..., O.SD_TOUCH_R],
            "com_distance": st["com_distance"], "done": rng.random((T, N)) < 0.1}


# conditioning of 1 - (q.q)^2 with error scales 0.01-0.03 amplifies 1-ulp libm differences ~1e2-1e3x: these two terms
# are checked at 2e-3 relative against the fp32 oracle (the fp32 oracle itself is only 2e-3 from fp64, see
# tests/test_oracle_cpu.py::test_fp32_oracle_tracks_fp64) AND at the same bound against the fp64 oracle.
ILL_CONDITIONED = {"roll_pitch": 2e-3, "feet_orient": 2e-3}


@pytest.mark.parametrize("T,N", [(16, 32), (40, 130), (1, 7)])
def test_rewards(eng, dev, T, N):
    b = Batch(600 + N, T, N, dev)
    tr = _traj(b)
    ld = b.ld
    c0 = O.reward_initial_carry((N,))
    c0["t_single"] = (np.random.default_rng(3).random(N) * 3).astype(np.float32)
    comp_np, total_np, c1 = O.rewards(tr, c0, P)
    carry = {"t_single": synth.to_soa(c0["t_single"], 0, dev), "airtime": synth.to_soa(c0["airtime"], 0, dev),
             "prev_contact": synth.to_soa(c0["prev_contact"].astype(np.uint8), 0, dev)}
    comp = torch.zeros((T, 12, ld), device=dev)
    total = eng.rewards(b.state, synth.to_soa(tr["command"], 1, dev), synth.to_soa(tr["ctrl"], 1, dev),
                        synth.to_soa(tr["done"].astype(np.uint8), 1, dev), carry, components=comp, n_envs=N)
    tr64 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in tr.items()}
    c064 = {k: (v.astype(np.float64) if v.dtype == np.float32 else v) for k, v in c0.items()}
    comp64, total64, _ = O.rewards(tr64, c064, P)
    for i, name in enumerate(O.REWARD_NAMES):
        g = S(comp[:, i], N)
        if name in ILL_CONDITIONED:
            close(g, comp_np[name], name, rtol=ILL_CONDITIONED[name])
            close(g, comp64[name], name + " vs fp64", rtol=ILL_CONDITIONED[name])
        elif name in ("single_contact", "no_contact_p"):
            exact(g, comp_np[name].astype(np.float32), name)
        else:
            close(g, comp_np[name], name)
    close(S(total, N), total_np, "total", rtol=4e-4)     # bounded by the two ill-conditioned terms (scale 0.2, 0.1)
    close(S(total, N), total64, "total vs fp64", rtol=4e-4)
    close(S(carry["t_single"], N), c1["t_single"], "carry t_single")
    close(S(carry["airtime"], N, (2,)), c1["airtime"], "carry airtime")
    exact(S(carry["prev_contact"], N, (2,)).astype(bool), c1["prev_contact"], "carry prev_contact")


def test_rewards_streaming_equivalence_full_size(eng, dev):
    """Property at BASELINE config-2 size (4096 envs x 100 steps): two half-rollouts chained through the reward carry
    reproduce the stateful terms of the whole rollout bit for bit."""
    T, N = 100, 4096
    d = synth.make_batch_device(21, T, N, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    cmd = torch.zeros((T, 16, N), device=dev)
    cmd[:, 0] = (torch.rand(T, N, generator=g, device=dev) < 0.6).float() * 0.7
    ctrl = 20 * torch.randn(T, 20, N, generator=g, device=dev)
    done = (torch.rand(T, N, generator=g, device=dev) < 0.02).to(torch.uint8)

    def carry0():
        return {"t_single": torch.zeros(N, device=dev), "airtime": torch.zeros((2, N), device=dev),
                "prev_contact": torch.ones((2, N), device=dev, dtype=torch.uint8)}

    cw = carry0()
    comp_w = torch.zeros((T, 12, N), device=dev)
    eng.rewards(d["state"], cmd, ctrl, done, cw, components=comp_w, n_envs=N)
    ch = carry0()
    h = T // 2
    comp_a = torch.zeros((h, 12, N), device=dev)
    comp_b = torch.zeros((T - h, 12, N), device=dev)
    eng.rewards({k: v[:h].contiguous() for k, v in d["state"].items()}, cmd[:h].contiguous(), ctrl[:h].contiguous(),
                done[:h].contiguous(), ch, components=comp_a, n_envs=N)
    eng.rewards({k: v[h:].contiguous() for k, v in d["state"].items()}, cmd[h:].contiguous(), ctrl[h:].contiguous(),
                done[h:].contiguous(), ch, components=comp_b, n_envs=N)
    for i in (5, 7):   # single_contact, feet_airtime
        assert torch.equal(torch.cat([comp_a[:, i], comp_b[:, i]]), comp_w[:, i]), O.REWARD_NAMES[i]
    for k in cw:
        assert torch.equal(cw[k], ch[k]), k
    assert torch.isfinite(comp_w).all()


@pytest.mark.parametrize("T,N,norm", [(4, 1, 0), (16, 32, 0), (100, 130, 0), (256, 33, 0), (16, 32, 1)])
def test_gae(dev, T, N, norm):
    e, _, _ = Hn.make_engine(device=dev, normalize_advantages=norm)
    p = O.OracleParams(normalize_advantages=norm)
    rng = np.random.default_rng(T * 1000 + N)
    v = rng.standard_normal((T, N)).astype(np.float32)
    r = (1.5 * rng.random((T, N))).astype(np.float32)
    done = rng.random((T, N)) < 0.05
    succ = done & (rng.random((T, N)) < 0.3)
    adv, tgt = e.gae(synth.to_soa(v, 1, dev), synth.to_soa(r, 1, dev), synth.to_soa(done.astype(np.uint8), 1, dev),
                     synth.to_soa(succ.astype(np.uint8), 1, dev), n_envs=N)
    a_np, t_np = O.compute_ppo_inputs(v, r, done, succ, p)
    close(S(adv, N), a_np, "advantages", atol=1e-5)
    close(S(tgt, N), t_np, "value_targets", atol=1e-5)
    e.close()


def test_gae_golden_exact(dev):
    from pathlib import Path

    g = np.load(Path(__file__).resolve().parent / "golden" / "gae.npz")
    e, _, _ = Hn.make_engine(device=dev, gamma=float(g["gamma"]), lam=float(g["lam"]))
    f = lambda k, dt: synth.to_soa(g[k].astype(dt)[:, None], 1, dev)
    adv, tgt = e.gae(f("values", np.float32), f("rewards", np.float32), f("done", np.uint8), f("success", np.uint8),
                     n_envs=1)
    exact(S(adv, 1)[:, 0], g["adv"].astype(np.float32), "golden adv")        # dyadic rationals: exact in fp32
    exact(S(tgt, 1)[:, 0], g["targets"].astype(np.float32), "golden targets")
    e.close()


def test_gae_linearity_full_size(eng, dev):
    """Property at 16384 envs x 256 steps (BASELINE config 3): with done = 0, GAE is linear in (values, rewards)."""
    T, N = 256, 16384
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    v1, v2, r1, r2 = (torch.randn(T, N, generator=g, device=dev) for _ in range(4))
    z = torch.zeros((T, N), device=dev, dtype=torch.uint8)
    a1, _ = eng.gae(v1, r1, z, z)
    a2, _ = eng.gae(v2, r2, z, z)
    a12, t12 = eng.gae(v1 + v2, r1 + r2, z, z)
    assert torch.allclose(a12, a1 + a2, rtol=1e-4, atol=1e-4)
    assert torch.allclose(t12, a12 + (v1 + v2), rtol=0, atol=1e-5)
    # done everywhere: A_t = r_t - v_t exactly (mask kills bootstrap and recursion)
    one = torch.ones_like(z)
    a, _ = eng.gae(v1, r1, one, z)
    assert torch.equal(a, r1 - v1)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("N", [1, 33, 1024, 4100])
def test_policy_step(dev, path, N):
    """convert.py step_fn: AoS inputs + flat carry -> (mode, carry)."""
    e, wa, _ = Hn.make_engine(gemm_path=path, device=dev)
    p = O.OracleParams()
    b = Batch(700 + N, 2, N, dev)
    flat = np.zeros((N, 2 * 2 * 256 + 20), np.float32)
    flat_d = torch.from_numpy(flat).to(dev)
    for t in range(2):
        o, _ = O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], None, P)
        args = (o["noisy_biased_joint_position"], o["noisy_joint_velocity"], o["noisy_imu_projected_gravity"],
                o["noisy_imu_gyro"], b.cmd0_np)
        act, flat = O.policy_step(wa, *args, flat, p)
        act_d, flat_d = e.policy_step(*(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in args), flat_d)
        close(act_d.cpu().numpy(), act, f"policy action t={t}")
        close(flat_d.cpu().numpy(), flat, f"policy carry t={t}")
    e.close()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("T,N,hidden", [(16, 32, 256), (6, 130, 256), (5, 50, 128), (7, 1100, 256), (1, 300, 256),
                                        (70, 40, 256),    # 70 steps: three 32-step chunks of the command / reward scans
                                        (2, 1, 256), (3, 5, 256)])   # ld = 4 / 8: TMA boxes far wider than the env axis
def test_rollout_fused(dev, path, T, N, hidden):
    """BASELINE config 1 shape (num_envs = 32, short rollout): the fused control step over T recorded steps."""
    res = Hn.run_rollout_case(seed=800 + N, T=T, N=N, hidden=hidden, device=dev, gemm_path=path)
    bad = {k: v for k, v in res["errors"].items() if not v <= 1.0}
    assert not bad, f"scaled errors > 1: {bad} (all: {res['errors']})"
    assert res["launches"] > 0


@pytest.mark.parametrize("path", [pytest.param(L.GEMM_TC_2XF16, id="tc2xf16"), pytest.param(L.GEMM_TC_3XTF32, id="tc3xtf32")])
@pytest.mark.parametrize("T,N", [(100, 4096), (256, 512)])
def test_rollout_fused_at_reference_length_against_oracle(dev, path, T, N):
    """the fused rollout against the ORACLE (not kernel-vs-kernel) at BASELINE configs[1]'s full size
    (4 096 envs x the reference's 100-step rollout, train.py:1763-1766, 1776) and at configs[2]'s 256 steps: the truncating
    tcgen05 accumulation is the error that grows with T (tools/error_growth.py prints its distribution over time,
    profiles/r02_error_growth.md)."""
    res = Hn.run_rollout_case(seed=4000 + N + T, T=T, N=N, hidden=256, device=dev, gemm_path=path)
    bad = {k: v for k, v in res["errors"].items() if not v <= 1.0}
    assert not bad, f"scaled errors > 1: {bad} (all: {res['errors']})"


def test_rollout_configs2_size_by_env_tiling(dev):
    """BASELINE configs[2] (16 384 envs x 256 steps) at full size through a size-independent property: environments are
    independent, so a batch made of 32 copies of a 512-env batch (which IS compared with the oracle at T = 256 above) must
    give, bit for bit, 32 copies of that batch's results -- whatever the panel a copy lands in, the order the 128-env panels
    are scheduled in, and the 256-step depth of the dependency counters."""
    T, N0, R = 256, 512, 32
    b = Batch(5150, T, N0, dev)
    small = Hn.rollout_buffers(b, 256, 2)
    e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
    e.rollout(small, N0)
    torch.cuda.synchronize()
    assert e.device_status() == 0

    def tile(v):
        if isinstance(v, dict):
            return {k: tile(x) for k, x in v.items()}
        if torch.is_tensor(v) and v.shape[-1] == N0:
            return v.repeat(*([1] * (v.dim() - 1)), R)
        return v

    fresh = Hn.rollout_buffers(b, 256, 2)                  # untouched inputs / zeroed outputs of the same batch
    big = {k: tile(v) for k, v in fresh.items() if k not in ("actor_carry", "critic_carry")}
    for k in ("actor_carry", "critic_carry"):
        big[k] = torch.zeros((2, 2, N0 * R, 256), device=dev)
    del fresh
    e.rollout(big, N0 * R)
    torch.cuda.synchronize()
    assert e.device_status() == 0
    for k in ("action", "log_prob", "value", "ctrl", "done", "success", "term_codes", "actor_obs", "lpf", "pg_carry", "command"):
        got = big[k].reshape(*big[k].shape[:-1], R, N0)
        ref = small[k].unsqueeze(-2).expand_as(got)
        assert torch.equal(got, ref), f"{k}: a copy of the batch differs from the 512-env run"
    for k in ("actor_carry", "critic_carry"):
        got = big[k].reshape(2, 2, R, N0, 256)
        assert torch.equal(got, small[k].unsqueeze(2).expand_as(got)), k
    e.close()


@pytest.mark.parametrize("T,N", [(12, 2048), (40, 4096)])
def test_rollout_persistent_vs_per_step_launches(dev, T, N, monkeypatch):
    """Size-independent property at BASELINE configs[1] scale: the persistent recurrence kernel (one launch, CTAs
    synchronised through progress counters) and the per-step launch sequence (kernel boundaries as the only
    synchronisation) run the same split-precision tcgen05 datapath with different tile shapes (128 x 256 vs 128 x 128
    tiles, so the truncating accumulation order differs): they agree to rounding level, a missed dependency would show
    as an O(0.1) error.  The heads differ in summation order (tensor-core vs FFMA).  Each path on its own is bitwise
    reproducible run to run (one MMA-issuing thread per CTA): checked by running the persistent kernel twice."""
    b = Batch(4242, T, N, dev)
    outs = []
    for per_step in ("0", "1"):
        monkeypatch.setenv("KBS_TC_PER_STEP", per_step)
        e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
        io = Hn.rollout_buffers(b, 256, 2)
        l0 = e.launches
        e.rollout(io, N)
        torch.cuda.synchronize()
        assert e.device_status() == 0
        outs.append((io, e.launches - l0))
        e.close()
    (a, la), (c, lc) = outs
    assert la < lc and lc - la >= 3 * T - 10, (la, lc)          # one launch instead of 3 per step
    monkeypatch.setenv("KBS_TC_PER_STEP", "0")
    e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
    again = Hn.rollout_buffers(b, 256, 2)
    e.rollout(again, N)
    torch.cuda.synchronize()
    e.close()
    for k in ("actor_carry", "critic_carry", "action", "log_prob", "value", "ctrl", "lpf"):
        assert torch.equal(a[k], again[k]), f"{k}: the persistent kernel is not bitwise reproducible"
    close(a["actor_carry"].cpu().numpy(), c["actor_carry"].cpu().numpy(), "actor carry", atol=2e-6)
    close(a["critic_carry"].cpu().numpy(), c["critic_carry"].cpu().numpy(), "critic carry", atol=2e-6)
    close(S(a["action"], N, (20,)), S(c["action"], N, (20,)), "action")
    close(S(a["log_prob"], N), S(c["log_prob"], N), "log_prob", atol=1e-5)
    close(S(a["value"], N), S(c["value"], N), "value", atol=1e-5)
    close(S(a["ctrl"], N, (20,)), S(c["ctrl"], N, (20,)), "ctrl", atol=1e-4)
    close(S(a["lpf"], N, (20,)), S(c["lpf"], N, (20,)), "lpf")


@pytest.mark.parametrize("T,N", [(30, 4096)])
def test_rollout_folded_input_projection_vs_projection_launch(dev, T, N, monkeypatch):
    """The actor's input projection folded into LSTM layer 0 of the persistent kernel ((W_ih W_in) o, packed once) against
    the reference's two-step form (x = W_in o + b_in as its own tensor-core launch, then W_ih x) at BASELINE configs[1]
    width: same mathematics, different rounding points -- they must agree far inside the 1e-5 tolerance."""
    b = Batch(777, T, N, dev)
    outs = []
    for off in ("0", "1"):
        monkeypatch.setenv("KBS_NO_FUSED_INPUT", off)
        e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
        io = Hn.rollout_buffers(b, 256, 2)
        e.profile(True)
        e.rollout(io, N)
        torch.cuda.synchronize()
        prof = e.profile_read()
        e.profile(False)
        assert e.device_status() == 0
        outs.append((io, prof))
        e.close()
    (a, pa), (c, pc) = outs
    assert pa["proj_tc_kernel"][1] == 1 and pc["proj_tc_kernel"][1] == 2   # critic: a launch in both; actor: only unfolded
    close(a["actor_carry"].cpu().numpy(), c["actor_carry"].cpu().numpy(), "actor carry", atol=2e-6)
    assert torch.equal(a["critic_carry"], c["critic_carry"])            # the critic's datapath is untouched
    close(S(a["action"], N, (20,)), S(c["action"], N, (20,)), "action")
    close(S(a["log_prob"], N), S(c["log_prob"], N), "log_prob", atol=1e-5)
    close(S(a["ctrl"], N, (20,)), S(c["ctrl"], N, (20,)), "ctrl", atol=1e-4)
    close(S(a["lpf"], N, (20,)), S(c["lpf"], N, (20,)), "lpf")


@pytest.mark.parametrize("T,N", [(20, 4096), (3, 200)])
def test_rollout_input_projection_from_soa_vs_staged(dev, T, N, monkeypatch):
    """input_proj_fused_kernel (feature rows bulk-copied straight from the env-major observations / the recorded cinert /
    cvel arrays, converted to the MMA operand in shared memory, 256-column tile) against the staged form (pack kernel ->
    SB buffer -> 128-column MODE_PROJ launch): same products, different accumulator tiling -> rounding-level agreement."""
    b = Batch(778, T, N, dev)
    outs = []
    for staged in ("0", "1"):
        monkeypatch.setenv("KBS_PROJ_STAGED", staged)
        e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
        io = Hn.rollout_buffers(b, 256, 2)
        l0 = e.launches
        e.rollout(io, N)
        torch.cuda.synchronize()
        assert e.device_status() == 0
        outs.append((io, e.launches - l0))
        e.close()
    (a, la), (c, lc) = outs
    assert lc == la + 1, (la, lc)                                       # the critic's pack launch is gone
    assert torch.equal(a["actor_carry"], c["actor_carry"])              # the actor's datapath is untouched
    close(a["critic_carry"].cpu().numpy(), c["critic_carry"].cpu().numpy(), "critic carry", atol=2e-6)
    close(S(a["value"], N), S(c["value"], N), "value", atol=1e-5)


def test_upload_state_moves_every_row_the_path_reads(dev):
    """kbs_upload_state copies 473 of the 676 MuJoCo rows.  Device arrays pre-filled with NaN + that upload must give
    the same rollout, rewards and GAE, bit for bit, as the fully populated state (any row the path reads but the
    upload skips would poison the outputs)."""
    T, N = 5, 260
    b = Batch(4343, T, N, dev)
    e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)

    def run(state):
        io = Hn.rollout_buffers(b, 256, 2)
        io["state"] = state
        e.rollout(io, N)
        carry = {"t_single": torch.zeros(b.ld, device=dev), "airtime": torch.zeros((2, b.ld), device=dev),
                 "prev_contact": torch.ones((2, b.ld), device=dev, dtype=torch.uint8)}
        total = e.rewards(state, io["command"][:T].contiguous(), io["ctrl"], io["done"], carry, n_envs=N)
        adv, tgt = e.gae(io["value"], total, io["done"], io["success"], n_envs=N)
        torch.cuda.synchronize()
        return {"action": io["action"], "value": io["value"], "done": io["done"], "total": total, "adv": adv, "tgt": tgt,
                "ctrl": io["ctrl"], "log_prob": io["log_prob"], "term": io["term_codes"]}

    ref = run(b.state)
    host = {k: v.cpu().pin_memory() for k, v in b.state.items()}
    poisoned = {k: torch.full_like(v, float("nan")) for k, v in b.state.items()}
    nbytes = e.upload_state(host, poisoned)
    torch.cuda.synchronize()
    assert nbytes == 473 * 4 * b.ld * T
    out = run(poisoned)
    for k in ref:
        assert torch.equal(ref[k][..., :N], out[k][..., :N]), k
    e.close()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("depth,with_critic,N", [(1, True, 140), (2, False, 200), (1, False, 33)])
def test_rollout_fused_other_shapes(dev, path, depth, with_critic, N):
    """Shapes off the launch configuration: one LSTM layer, actor only (the persistent kernel's item list and counter
    protocol depend on both)."""
    res = Hn.run_rollout_case(seed=860 + N, T=5, N=N, hidden=256, device=dev, gemm_path=path, depth=depth,
                              with_critic=with_critic)
    bad = {k: v for k, v in res["errors"].items() if not v <= 1.0}
    assert not bad, f"scaled errors > 1: {bad} (all: {res['errors']})"


def test_rollout_then_rewards_gae_chain(dev):
    """Whole path once: rollout -> rewards -> GAE on its outputs, against the same chain in the oracle."""
    T, N = 12, 64
    b = Batch(901, T, N, dev)
    e, wa, wc = Hn.make_engine(device=dev)
    io = Hn.rollout_buffers(b, 256, 2)
    e.rollout(io, N)
    ref = Hn.oracle_rollout(b, wa, wc, 256, 2)
    c0 = O.reward_initial_carry((N,))
    st = b.np["state"]
    tr = {"xquat": st["xquat"], "xpos": st["xpos"], "qpos": st["qpos"], "qvel": st["qvel"], "ctrl": ref["ctrl"],
          "command": ref["command"], "touch_l": st["sensordata"][..., O.SD_TOUCH_L],
          "touch_r": st["sensordata"][..., O.SD_TOUCH_R], "com_distance": st["com_distance"], "done": ref["done"]}
    _, total_np, _ = O.rewards(tr, c0, P)
    a_np, t_np = O.compute_ppo_inputs(ref["value"], total_np, ref["done"], ref["success"], P)
    carry = {"t_single": torch.zeros(b.ld, device=dev), "airtime": torch.zeros((2, b.ld), device=dev),
             "prev_contact": torch.ones((2, b.ld), device=dev, dtype=torch.uint8)}
    total = e.rewards(b.state, io["command"][:T].contiguous(), io["ctrl"], io["done"], carry, n_envs=N)
    adv, tgt = e.gae(io["value"], total, io["done"], io["success"], n_envs=N)
    close(S(total, N), total_np, "total reward", rtol=4e-4)
    close(S(adv, N), a_np, "advantages", rtol=4e-4, atol=1e-4)
    close(S(tgt, N), t_np, "value targets", rtol=4e-4, atol=1e-4)
    e.close()


@pytest.mark.parametrize("T,N,clipped", [(7, 33, True), (100, 4096, True), (20, 1001, False)])
def test_ppo_loss(eng, dev, T, N, clipped):
    """ksim.compute_ppo_loss (clipped surrogate + value loss + entropy bonus) as one deterministic reduction."""
    rng = np.random.default_rng(77 + N)
    f = np.float32
    old_lp = rng.normal(-20, 5, (T, N)).astype(f)
    lp = (old_lp + rng.normal(0, 0.3, (T, N))).astype(f)
    lp[0, 0] = old_lp[0, 0] + 30.0                                    # beyond log_clip_value
    adv = rng.normal(0, 1, (T, N)).astype(f)
    v_old = rng.normal(0, 1, (T, N)).astype(f)
    v = (v_old + rng.normal(0, 0.3, (T, N))).astype(f)
    tgt = (v_old + rng.normal(0, 0.5, (T, N))).astype(f)
    ent = rng.normal(25, 3, (T, N)).astype(f)
    ref = O.ppo_loss(lp, old_lp, adv, v, v_old, tgt, ent, use_clipped_value_loss=clipped)
    d = lambda a: synth.to_soa(a, 1, dev)
    per = torch.full((T, (N + 3) // 4 * 4), float("nan"), device=dev)
    args = (d(lp), d(old_lp), d(adv), d(v), d(v_old), d(tgt), d(ent))
    out = eng.ppo_loss(*args, per_step=per, n_envs=N, use_clipped_value_loss=int(clipped))
    got = out.cpu().numpy()
    close(got, np.array(ref[:4], np.float32), "loss, mean policy / value / entropy", atol=1e-6)
    close(S(per, N), ref[4], "per-step objective", atol=1e-5)
    again = eng.ppo_loss(*args, n_envs=N, use_clipped_value_loss=int(clipped))
    assert torch.equal(out, again), "the reduction order is fixed: bitwise reproducible"


_ppo_batch = Hn.ppo_batch


@pytest.mark.parametrize("path", [pytest.param(L.GEMM_SIMT_FP32, id="simt"), pytest.param(L.GEMM_TC_2XF16, id="tc2xf16"),
                                  pytest.param(L.GEMM_TC_3XTF32, id="tc3xtf32")])
@pytest.mark.parametrize("T,N,hidden", [(5, 24, 128), (7, 70, 256), (4, 300, 256)])
def test_ppo_grad_against_autograd(dev, T, N, hidden, path):
    """8f-1: back-propagation through time through both LSTMs, the projections, the actor head (incl. the low-pass filter's
    recurrence and the done-resets) and the PPO loss, against torch.autograd (float64) on the same minibatch.  On the
    FP16-split datapath this is the persistent form (rollout_persist_kernel<SAVE> + bptt_persist_kernel + split-K tcgen05
    weight-gradient GEMMs); the other two datapaths keep the per-step launch sequence."""
    _ppo_grad_case(dev, T, N, hidden, path)


@pytest.mark.parametrize("T,N,hidden", [(7, 70, 256), (33, 300, 256), (6, 130, 128)])
def test_ppo_grad_per_step_launches_against_autograd(dev, T, N, hidden, monkeypatch):
    """The per-step launch sequence of the FP16-split datapath (KBS_PPO_PER_STEP=1): the cross-check of the persistent form."""
    monkeypatch.setenv("KBS_PPO_PER_STEP", "1")
    _ppo_grad_case(dev, T, N, hidden, L.GEMM_TC_2XF16)


@pytest.mark.parametrize("T,N", [(9, 256), (26, 384)])
def test_ppo_grad_in_place_weight_gradients_vs_repacked(dev, T, N, monkeypatch):
    """dw_gemm_kernel (dG / x / h_in read in place as MN-major UMMA operands, whole panels only) against the K-major GEMM on
    re-packed operands (KBS_DW_DIRECT=0): the same split products over the same accumulation runs -- the gradients must agree
    far below the tolerance against autograd (which both forms are also held to)."""
    _ppo_grad_case(dev, T, N, 256, L.GEMM_TC_2XF16)
    e, wa, wc = Hn.make_engine(hidden=256, gemm_path=L.GEMM_TC_2XF16, device=dev)
    p = O.OracleParams(hidden_size=256)
    b = _ppo_batch(np.random.default_rng(7 + N), T, N, 256, p, wa, wc)
    d = lambda a, dt=None: synth.to_soa(a if dt is None else a.astype(dt), 1, dev)
    batch = {"actor_obs": d(b["actor_obs"]), "critic_obs": d(b["critic_obs"]), "action": d(b["action"]),
             "done": d(b["done"], np.uint8), "old_log_probs": d(b["old_log_probs"]), "advantages": d(b["advantages"]),
             "value_targets": d(b["value_targets"]), "old_values": d(b["old_values"])}

    def grads():
        z = lambda a: torch.full(a.shape, float("nan"), device=dev)
        mk = lambda w: {"w_in": z(w["w_in"]), "b_in": z(w["b_in"]), "w_out": z(w["w_out"]), "b_out": z(w["b_out"]),
                        "layers": [{k: z(l[k]) for k in ("w_ih", "w_hh", "b")} for l in w["layers"]]}
        ga, gc = mk(wa), mk(wc)
        e.ppo_grad(batch, ga, gc, n_envs=N)
        torch.cuda.synchronize()
        assert e.device_status() == 0
        return [g["layers"][l][k].clone() for g in (ga, gc) for l in range(2) for k in ("w_ih", "w_hh", "b")]

    direct = grads()
    monkeypatch.setenv("KBS_DW_DIRECT", "0")
    repacked = grads()
    monkeypatch.delenv("KBS_DW_DIRECT")
    e.close()
    for a, r in zip(direct, repacked):
        assert torch.isfinite(a).all()
        scale = float(r.abs().max())
        assert float((a - r).abs().max()) <= 2e-6 * scale + 1e-12, (float((a - r).abs().max()), scale)


def test_ppo_grad_large_minibatch_by_tiling(dev):
    """BASELINE configs[3]'s per-GPU share of the 65 536-env batch (8 192 trajectories x 100 steps) through a size-independent
    property: the loss is a mean over trajectories, so a minibatch made of 16 copies of a 512-trajectory minibatch (compared
    with autograd above) has the same gradient.  At this size the recurrence kernels run many items per CTA and slot and the
    weight-gradient GEMMs 256 accumulation runs: only the order of the fp32 sums differs."""
    T, N0, R, H = 100, 512, 16, 256
    e, wa, wc = Hn.make_engine(hidden=H, gemm_path=L.GEMM_TC_2XF16, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    rn = lambda *s, sc=1.0: torch.randn(s, generator=g, device=dev) * sc
    small = {"actor_obs": rn(T, 65, N0, sc=0.7), "critic_obs": rn(T, 475, N0, sc=0.7), "action": rn(T, 20, N0, sc=0.3),
             "done": (torch.rand((T, N0), generator=g, device=dev) < 0.02).to(torch.uint8),
             "advantages": rn(T, N0), "value_targets": rn(T, N0, sc=0.5)}
    z = lambda n: torch.zeros((2, 2, n, H), device=dev)
    fwd = e.ppo_variables(small["actor_obs"], small["action"], small["done"], z(N0), torch.zeros((20, N0), device=dev),
                          small["critic_obs"], z(N0), want_std=False, n_envs=N0)
    small["old_log_probs"] = fwd["log_probs"] + rn(T, N0, sc=0.1)
    small["old_values"] = fwd["values"] + rn(T, N0, sc=0.2)

    def grads(batch, n):
        mk = lambda w: {"w_in": torch.zeros_like(w["w_in"]), "b_in": torch.zeros_like(w["b_in"]), "w_out": torch.zeros_like(w["w_out"]),
                        "b_out": torch.zeros_like(w["b_out"]),
                        "layers": [{k: torch.zeros_like(l[k]) for k in ("w_ih", "w_hh", "b")} for l in w["layers"]]}
        ga, gc = mk(synth.weights_to_device(wa, dev)), mk(synth.weights_to_device(wc, dev))
        out = e.ppo_grad(batch, ga, gc, n_envs=n)
        torch.cuda.synchronize()
        assert e.device_status() == 0
        flat = [x[k] for x in (ga, gc) for k in ("w_in", "b_in", "w_out", "b_out")]
        flat += [x["layers"][l][k] for x in (ga, gc) for l in range(2) for k in ("w_ih", "w_hh", "b")]
        return [t.clone() for t in flat], out["stats"].clone()

    ref, ref_stats = grads(small, N0)
    big = {k: v.repeat(*([1] * (v.dim() - 1)), R) for k, v in small.items()}
    got, got_stats = grads(big, N0 * R)
    e.close()
    assert torch.allclose(got_stats, ref_stats, rtol=2e-5, atol=1e-6), (got_stats, ref_stats)
    for a, r in zip(got, ref):
        scale = float(r.abs().max())
        assert torch.isfinite(a).all() and float((a - r).abs().max()) <= 2e-5 * scale + 1e-12, (float((a - r).abs().max()), scale)


@pytest.mark.parametrize("T,N", [(1, 128), (1, 5), (2, 129), (3, 640)])
def test_ppo_grad_edge_shapes(dev, T, N):
    """One-step rollouts, a handful of trajectories, a panel plus one row, whole panels with T x panels below one accumulation
    run (the in-place weight-gradient GEMM then has empty runs to zero-fill)."""
    _ppo_grad_case(dev, T, N, 256, L.GEMM_TC_2XF16)


@pytest.mark.parametrize("T,N", [(33, 300), (100, 512)])
def test_ppo_grad_persistent_longer_rollouts(dev, T, N):
    """The persistent update at the reference's rollout length (T = 100, 512 trajectories: train.py:1764-1766)."""
    _ppo_grad_case(dev, T, N, 256, L.GEMM_TC_2XF16)


def _ppo_grad_case(dev, T, N, hidden, path):
    bad = Hn.run_ppo_grad_case(dev, T, N, hidden, path)
    assert not bad, bad


def test_ppo_update_adamw_step_decreases_loss(dev):
    """optax.adamw step (kbs_adamw_step) on the flat parameter vector: matches the closed form on the first step and a few
    updates on a fixed minibatch decrease the PPO loss."""
    from kbot_joystick_b200.ppo import PpoUpdater

    T, N, H = 6, 48, 128
    e, wa, wc = Hn.make_engine(hidden=H, gemm_path=L.GEMM_TC_2XF16, device=dev)
    p = O.OracleParams(hidden_size=H)
    rng = np.random.default_rng(5)
    b = _ppo_batch(rng, T, N, H, p, wa, wc)
    d = lambda a, dt=None: synth.to_soa(a if dt is None else a.astype(dt), 1, dev)
    batch = {"actor_obs": d(b["actor_obs"]), "critic_obs": d(b["critic_obs"]), "action": d(b["action"]),
             "done": d(b["done"], np.uint8), "old_log_probs": d(b["old_log_probs"]), "advantages": d(b["advantages"]),
             "value_targets": d(b["value_targets"]), "old_values": d(b["old_values"])}
    up = PpoUpdater(e, wa, wc, lr=1e-3)
    p0 = up.param.clone()
    first = up.update(batch, N)
    g = up.grad.clone()
    # first AdamW step: m/(1-b1) = g, v/(1-b2) = g^2  ->  delta = -lr (g / (|g| + eps) + wd p)   (|g|_2 << the clip norm)
    assert float(up.norm) < 10.0 and up.step_count == 1
    expect = p0 - 1e-3 * (g / (g.abs() + 1e-8) + 1e-5 * p0)
    assert torch.allclose(up.param, expect, rtol=1e-5, atol=1e-7)
    losses = [float(first["stats"][0])]
    for _ in range(6):
        losses.append(float(up.update(batch, N)["stats"][0]))
    assert losses[-1] < losses[0], losses
    e.close()


@pytest.mark.parametrize("count", [1, 1000, 2250537])
def test_adamw_and_grad_norm_against_float64(eng, dev, count):
    """kbs_grad_norm + kbs_adamw_step against the float64 restatement of optax.adamw + ksim's clip (oracle.adamw_update) and
    against torch.optim.AdamW (decoupled decay = optax.adamw semantics) over several steps: unclipped, clipped, device-side
    step counter, and the non-finite skip (parameters, moments and the counter untouched)."""
    rng = np.random.default_rng(count)
    p0 = rng.normal(0, 0.3, count).astype(np.float32)
    opt = dict(lr=5e-4, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-5)
    for max_norm, gscale in ((0.0, 1.0), (0.5, 0.125)):
        param = torch.from_numpy(p0.copy()).to(dev)
        m, v = torch.zeros_like(param), torch.zeros_like(param)
        step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        norm = torch.zeros(1, device=dev)
        rp, rm, rv = p0.astype(np.float64), np.zeros(count), np.zeros(count)
        tp = torch.from_numpy(p0.astype(np.float64)).requires_grad_(True)
        topt = torch.optim.AdamW([tp], lr=opt["lr"], betas=(opt["b1"], opt["b2"]), eps=opt["eps"], weight_decay=opt["weight_decay"])
        for step in range(1, 6):
            g = (rng.normal(0, 1.0, count) * (0.02 if step != 3 else 3.0)).astype(np.float32)
            gd = torch.from_numpy(g).to(dev)
            eng.grad_norm(gd, out=norm)
            ref_norm = np.sqrt(np.sum(g.astype(np.float64) ** 2))
            assert abs(float(norm) - ref_norm) <= 2e-7 * ref_norm + 1e-30          # double accumulation, one fp32 rounding
            n2 = eng.grad_norm(gd)
            assert torch.equal(n2, norm)                                           # fixed reduction order: bitwise reproducible
            eng.adamw_step(param, gd, m, v, grad_norm=norm, step_dev=step_dev, grad_scale=gscale, max_grad_norm=max_norm, **opt)
            rp, rm, rv, applied = O.adamw_update(rp, g, rm, rv, step, grad_scale=gscale, max_grad_norm=max_norm, **opt)
            assert applied and int(step_dev) == step
            ge = g.astype(np.float64) * gscale
            ne = ref_norm * gscale
            if max_norm > 0 and ne > max_norm:
                ge = ge * (max_norm / ne)
            tp.grad = torch.from_numpy(ge)
            topt.step()
            close(param.cpu().numpy(), rp, f"adamw param step {step}", rtol=1e-6, atol=2e-7)
            close(m.cpu().numpy(), rm, f"adamw m step {step}", rtol=1e-5, atol=1e-9)
            close(v.cpu().numpy(), rv, f"adamw v step {step}", rtol=1e-5, atol=1e-12)
            np.testing.assert_allclose(rp, tp.detach().numpy(), rtol=0, atol=1e-12)   # the restatement == torch AdamW
        # non-finite gradient: ksim skips the update; nothing moves, the counter does not advance
        before = (param.clone(), m.clone(), v.clone())
        g = rng.normal(0, 0.02, count).astype(np.float32)
        g[count // 2] = np.inf
        gd = torch.from_numpy(g).to(dev)
        eng.grad_norm(gd, out=norm)
        assert not np.isfinite(float(norm))
        eng.adamw_step(param, gd, m, v, grad_norm=norm, step_dev=step_dev, grad_scale=gscale, max_grad_norm=max_norm, **opt)
        assert int(step_dev) == 5
        assert torch.equal(param, before[0]) and torch.equal(m, before[1]) and torch.equal(v, before[2])
    # weight_decay = 0, no norm, host-side step: the optax.adam branch (kbs_adam_step) and kbs_adamw_step agree bitwise
    g = torch.from_numpy(rng.normal(0, 0.02, count).astype(np.float32)).to(dev)
    pa, pb = torch.from_numpy(p0.copy()).to(dev), torch.from_numpy(p0.copy()).to(dev)
    ma, va, mb, vb = (torch.zeros_like(pa) for _ in range(4))
    eng.adam_step(pa, g, ma, va, 3)
    eng.adamw_step(pb, g, mb, vb, step=3, weight_decay=0.0, max_grad_norm=0.0)
    close(pb.cpu().numpy(), pa.cpu().numpy(), "adamw(wd=0) vs adam", rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("N", [3, 64, 1001])
def test_actuator_randomization_sampler(eng, dev, N):
    """8f-2: per-episode PositionActuators randomisation (train.py:1097-1105; sampling law [U]) -> the kp / kd / tau_limit /
    action_bias / torque_bias arrays kbs_torque consumes; partial resample with a reset mask."""
    ld = (N + 3) // 4 * 4
    rng = np.random.default_rng(N)
    u = rng.random((5, N, 20)).astype(np.float32)
    ref = O.actuator_randomization(u)
    ud = synth.to_soa(u, 1, dev)
    ep = {k: torch.full((20, ld), float("nan"), device=dev) for k in ("kp", "kd", "tau_limit", "action_bias", "torque_bias")}
    eng.sample_actuator_randomization(ud, ep, n_envs=N)
    for k in ep:
        close(S(ep[k], N, (20,)), ref[k], k, atol=1e-7)
    assert (S(ep["tau_limit"], N, (20,)) <= np.asarray(O.CTRL_LIMIT64, np.float32) + 1e-4).all()
    assert (np.abs(S(ep["action_bias"], N, (20,))) <= 0.02 + 1e-7).all() and (S(ep["torque_bias"], N, (20,)) == 0).all()
    # second draw, only where reset: the other environments keep their episode's values
    u2 = rng.random((5, N, 20)).astype(np.float32)
    reset = rng.random(N) < 0.4
    ref2 = O.actuator_randomization(u2)
    eng.sample_actuator_randomization(synth.to_soa(u2, 1, dev), ep, reset=synth.to_soa(reset.astype(np.uint8), 0, dev), n_envs=N)
    for k in ep:
        close(S(ep[k], N, (20,)), np.where(reset[:, None], ref2[k], ref[k]), k + " (masked)", atol=1e-7)
    # and the torque map accepts them
    act = synth.to_soa(rng.normal(0, 0.5, (N, 20)).astype(np.float32), 0, dev)
    b = Batch(5, 1, N, dev)
    ctrl = eng.torque(act, b.state_at(0), ep, n_envs=N)
    st = b.np_state_at(0)
    want = {k: np.where(reset[:, None], ref2[k], ref[k]) for k in ref}
    close(S(ctrl, N, (20,)), O.position_actuator_torque(S(act, N, (20,)), st["qpos"][:, 7:], st["qvel"][:, 6:], want["kp"],
                                                       want["kd"], want["tau_limit"], want["action_bias"], want["torque_bias"]),
          "ctrl with sampled gains", atol=1e-4)


@pytest.mark.parametrize("path", PATHS)
def test_checkpoint_bytes_to_policy_step(dev, path, tmp_path):
    """8f-4 (convert.py:36-39, 84-119): an eqx leaf stream on disk -> checkpoint.load_policy -> kbs_weights_pack ->
    kbs_policy_step, against oracle.policy_step on the weights that were serialised (init_fn() = zero carry)."""
    from kbot_joystick_b200 import checkpoint
    from kbot_joystick_b200.engine import KbotStep

    H, N = 256, 300
    wa, wc = synth.make_weights(501, 65, 40, H, 2), synth.make_weights(502, 475, 1, H, 2)
    static = (np.asarray(65), np.asarray(40), np.asarray(0.01, np.float32))       # 0-d static fields ride along, as eqx writes them
    path_ckpt = tmp_path / "model.eqx"
    checkpoint.write_leaves(checkpoint.weights_to_leaves(wa, wc, static, static[:2]), path_ckpt)
    la, lc = checkpoint.load_policy(path_ckpt, hidden=H, depth=2)
    e = KbotStep(hidden_size=H, depth=2, gemm_path=path)
    e.pack_weights(L.NET_ACTOR, synth.weights_to_device(la, dev))
    e.pack_weights(L.NET_CRITIC, synth.weights_to_device(lc, dev))
    p = O.OracleParams()
    b = Batch(17, 3, N, dev)
    flat = np.zeros((N, 2 * 2 * H + 20), np.float32)              # init_fn(): convert.py:67-72
    flat_d = torch.from_numpy(flat).to(dev)
    for t in range(3):
        o, _ = O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], None, P)
        args = (o["noisy_biased_joint_position"], o["noisy_joint_velocity"], o["noisy_imu_projected_gravity"],
                o["noisy_imu_gyro"], b.cmd0_np)
        act, flat = O.policy_step(wa, *args, flat, p)
        act_d, flat_d = e.policy_step(*(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in args), flat_d)
        close(act_d.cpu().numpy(), act, f"policy action from checkpoint t={t}")
        close(flat_d.cpu().numpy(), flat, f"policy carry from checkpoint t={t}")
    e.close()


def test_f16_split_range_guard_sets_status(dev):
    """An input beyond the FP16 planes' range (|x| >= 65504 or non-finite) on the 2xFP16-split datapath must not pass
    silently as inf: the health word gets KBS_STATUS_F16_RANGE, the next fused call fails with KBS_E_DEVICE until reset."""
    T, N = 2, 200
    b = Batch(31, T, N, dev)
    e, wa, wc = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
    io = Hn.rollout_buffers(b, 256, 2, True)
    e.rollout(io, N)
    assert e.device_status() == 0
    bad = {k: v.clone() for k, v in b.state.items()}
    bad["cinert"][1, 57, 5] = 1.0e5                              # a raw critic input (train.py:1405): beyond the half range
    io2 = dict(io)
    io2["state"] = bad
    e.rollout(io2, N)
    st = e.device_status()
    assert st & L.STATUS_F16_RANGE and not (st & 3), st
    with pytest.raises(RuntimeError, match="health word"):
        e.rollout(io, N)
    e.device_status_reset()
    for k in ("actor_carry", "critic_carry", "lpf", "pg_carry"):    # the poisoned run left inf / NaN in the carries it advanced
        io[k].zero_()
    e.rollout(io, N)
    assert e.device_status() == 0
    e.close()
    # the 3xTF32 datapath carries the same value without complaint
    e3, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_3XTF32, device=dev)
    io3 = Hn.rollout_buffers(b, 256, 2, True)
    io3["state"] = bad
    e3.rollout(io3, N)
    assert e3.device_status() == 0 and torch.isfinite(io3["value"]).all()
    e3.close()


def test_rollout_phase_a_chunks_on_aux_stream(dev, monkeypatch):
    """KBS_ROLLOUT_CHUNKS > 1 produces phase A on the handle's aux stream: the persistent kernel must wait for every chunk
    (it reads the operands of all T steps).  Same results as the single-stream default, bit for bit."""
    T, N = 12, 2048
    outs = []
    monkeypatch.setenv("KBS_NO_FUSED_INPUT", "1")      # chunked phase A keeps the actor's projection launch: same form for both
    for chunks in ("1", "3"):
        monkeypatch.setenv("KBS_ROLLOUT_CHUNKS", chunks)
        b = Batch(41, T, N, dev)
        e, wa, wc = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
        io = Hn.rollout_buffers(b, 256, 2, True)
        for _ in range(3):                                       # repeated: a race would not lose every time
            io["command"][1:].zero_()
            io["actor_carry"].zero_(); io["critic_carry"].zero_(); io["lpf"].zero_(); io["pg_carry"].zero_()
            e.rollout(io, N)
        torch.cuda.synchronize()
        assert e.device_status() == 0
        outs.append({k: io[k].clone() for k in ("action", "log_prob", "value", "ctrl", "actor_obs", "actor_carry")})
        e.close()
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_generate_rollout_noise_statistics_and_counter_property(eng, dev):
    """kbs_generate_rollout_noise (device-side Philox draws for callers that do not need JAX-threefry bit parity): ranges and
    moments of every array, determinism, and the counter property -- (seed, row, step, env) alone decides a value, so any
    split of a rollout into calls gives the same numbers; a rollout then runs on them."""
    T, N = 16, 4096
    ld = N

    def bufs():
        f = dict(device=dev, dtype=torch.float32)
        return {"noise": {"eps_jpos": torch.full((T, 20, ld), 9.0, **f), "eps_jvel": torch.full((T, 20, ld), 9.0, **f),
                          "eps_gyro": torch.full((T, 3, ld), 9.0, **f), "eps_pg": torch.full((T, 3, ld), 9.0, **f)},
                "eps_action": torch.full((T, 20, ld), 9.0, **f), "u_switch": torch.full((T, ld), 9.0, **f),
                "cmd_mode": torch.full((T, ld), 9, device=dev, dtype=torch.int32), "cmd_u6": torch.full((T, 6, ld), 9.0, **f),
                "cmd_u_arms": torch.full((T, 10, ld), 9.0, **f)}

    a = bufs()
    eng.generate_rollout_noise(a, N, seed=1234, step0=0)
    for k in ("eps_jpos", "eps_jvel"):
        x = a["noise"][k]
        assert float(x.min()) >= -1.0 and float(x.max()) < 1.0
        assert abs(float(x.mean())) < 5e-3 and abs(float(x.var()) - 1.0 / 3.0) < 5e-3
    for x in (a["noise"]["eps_gyro"], a["noise"]["eps_pg"], a["eps_action"]):
        assert torch.isfinite(x).all()
        assert abs(float(x.mean())) < 1e-2 and abs(float(x.var()) - 1.0) < 2e-2
        assert abs(float((x ** 4).mean()) - 3.0) < 0.15                     # Gaussian kurtosis
        assert 4.0 < float(x.abs().max()) < 7.0                             # tails present, Box-Muller bounded by sqrt(-2 ln 2^-24)
    for k in ("u_switch", "cmd_u6", "cmd_u_arms"):
        assert float(a[k].min()) >= 0.0 and float(a[k].max()) < 1.0 and abs(float(a[k].mean()) - 0.5) < 5e-3
    m = a["cmd_mode"]
    assert int(m.min()) == 0 and int(m.max()) == 5
    freq = torch.bincount(m.flatten(), minlength=6).float() / m.numel()
    assert float((freq - 1.0 / 6.0).abs().max()) < 5e-3
    # rows / steps / envs are decorrelated
    g = a["eps_action"]
    assert abs(float((g[0, 0] * g[0, 1]).mean())) < 0.05 and abs(float((g[0, 0] * g[1, 0]).mean())) < 0.05
    assert abs(float((g[0, 0, :-1] * g[0, 0, 1:]).mean())) < 0.05
    # deterministic, and the counter property: steps 5..15 regenerated on their own are the same numbers
    b = bufs()
    eng.generate_rollout_noise(b, N, seed=1234, step0=0)
    assert all(torch.equal(a["noise"][k], b["noise"][k]) for k in a["noise"]) and torch.equal(a["cmd_mode"], b["cmd_mode"])
    T2 = T - 5
    c = {k: (v[:T2] if k != "noise" else {kk: vv[:T2] for kk, vv in v.items()}) for k, v in bufs().items()}
    c = {k: (v.contiguous() if k != "noise" else {kk: vv.contiguous() for kk, vv in v.items()}) for k, v in c.items()}
    eng.generate_rollout_noise(c, N, seed=1234, step0=5)
    assert torch.equal(c["eps_action"], a["eps_action"][5:]) and torch.equal(c["noise"]["eps_pg"], a["noise"]["eps_pg"][5:])
    assert torch.equal(c["cmd_u_arms"], a["cmd_u_arms"][5:]) and torch.equal(c["cmd_mode"], a["cmd_mode"][5:])
    d = bufs()
    eng.generate_rollout_noise(d, N, seed=1235, step0=0)
    assert not torch.equal(d["eps_action"], a["eps_action"])
    # ragged n: nothing beyond ld is touched, envs < n are filled
    e2 = bufs()
    eng.generate_rollout_noise(e2, 4093, seed=7, step0=0)
    assert float(e2["u_switch"][:, :4093].max()) < 1.0


def test_rollout_on_device_generated_noise(dev):
    """The fused rollout consuming kbs_generate_rollout_noise output: the e2e configuration of bench.py (no noise upload)."""
    T, N = 6, 300
    b = Batch(77, T, N, dev)
    e, wa, wc = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
    io = Hn.rollout_buffers(b, 256, 2, True)
    io["noise"] = {k: torch.empty_like(v) for k, v in io["noise"].items()}
    for k in ("eps_action", "u_switch", "cmd_mode", "cmd_u6", "cmd_u_arms"):
        io[k] = torch.empty_like(io[k])
    e.generate_rollout_noise(io, N, seed=99, step0=0)
    e.rollout(io, N)
    torch.cuda.synchronize()
    assert e.device_status() == 0
    # the oracle on the SAME device-generated randomness: parity holds whatever produced the noise
    nz = {k: S(v, N, (v.shape[1],)) for k, v in io["noise"].items()}
    nz["eps_action"] = S(io["eps_action"], N, (20,))
    cr = {"u_switch": S(io["u_switch"], N), "mode": S(io["cmd_mode"], N), "u6": S(io["cmd_u6"], N, (6,)), "u_arms": S(io["cmd_u_arms"], N, (10,))}
    p = O.OracleParams()
    carry = {"actor": np.zeros((N, 2, 2, 256), np.float32), "critic": np.zeros((N, 2, 2, 256), np.float32),
             "lpf_params": np.zeros((N, 20), np.float32)}
    ref = O.rollout_control_steps(wa, wc, b.np["state"], nz, b.np["episode"], cr, b.cmd0_np, carry, np.zeros((N, 3), np.float32), p)
    errs = Hn.compare_rollout(io, ref, N, True)
    assert all(v <= 1.0 for v in errs.values()), errs
    e.close()


def test_error_codes(eng, dev):
    lib = L.load()
    assert lib.kbs_version() == 101
    p = L.default_params()
    p.hidden_size = 100
    h = C.c_void_p()
    assert lib.kbs_create(C.byref(p), C.byref(h)) == -2            # KBS_E_SHAPE
    v = torch.zeros((4, 8), device=dev)
    z = torch.zeros((4, 8), device=dev, dtype=torch.uint8)
    assert lib.kbs_gae(eng._h, None, L.ptr(v), L.ptr(z), L.ptr(z), L.ptr(v), L.ptr(v), 4, 8, 8, None) == -1   # NULL
    assert lib.kbs_gae(eng._h, L.ptr(v), L.ptr(v), L.ptr(z), L.ptr(z), L.ptr(v), L.ptr(v), 0, 8, 8, None) == -2  # T = 0
    b = Batch(1, 1, 8, dev)
    mis = torch.zeros(16 * 8 + 1, device=dev)[1:].view(16, 8)      # 4-byte offset: not 16-byte aligned
    sv = L.KbsStateView()
    for k, t in b.state_at(0).items():
        setattr(sv, k, L.ptr(t))
    sv.ld = 8
    out = torch.zeros((65, 8), device=dev)
    rc = lib.kbs_observations(eng._h, C.byref(sv), None, None, mis.data_ptr(), None, None, None, L.ptr(out), None, 8, None)
    assert rc == -3                                                 # KBS_E_ALIGN
    assert b"aligned" in lib.kbs_error_string(rc)
    # weights not packed -> KBS_E_STATE, never a silent fallback
    from kbot_joystick_b200.engine import KbotStep

    e2 = KbotStep()
    with pytest.raises(RuntimeError, match="not ready"):
        e2.actor_step(out, torch.zeros((2, 2, 8, 256), device=dev), torch.zeros((20, 8), device=dev))
    e2.close()


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("T,N,hidden", [(8, 50, 256), (5, 130, 128)])
def test_ppo_variables(dev, path, T, N, hidden):
    """get_ppo_variables (train.py:1435-1524) on a stored trajectory: log_probs / values / entropy / action_std and the
    final carries, incl. the done-resets; the mirror pass is the same call on mirrored observations."""
    e, wa, wc = Hn.make_engine(hidden=hidden, gemm_path=path, device=dev)
    # mirror-loss scales: the config defaults (train.py:115-122); the launch config multiplies both by 0.0 (train.py:1771-1772)
    p = O.OracleParams(hidden_size=hidden, actor_mirror_loss_scale=1.0, critic_mirror_loss_scale=0.01)
    b = Batch(1000 + N, T, N, dev)
    rng = np.random.default_rng(4)
    obs_list = [O.get_observations(b.np_state_at(t), b.np_noise_at(t), b.np["episode"], None, P)[0] for t in range(T)]
    cmd = np.stack([b.cmd0_np] * T)
    done = rng.random((T, N)) < 0.15
    act = np.stack([o["joint_position"] for o in obs_list]) + (0.2 * rng.standard_normal((T, N, 20))).astype(np.float32)
    carry0 = O.initial_model_carry((N,), p)
    ref, c_end = O.get_ppo_variables(wa, wc, obs_list, cmd, act.astype(np.float32), done, carry0, p, mirror=True)

    def run(olist, cmds, want_mean):
        a_obs = np.stack([O.actor_obs_from_dict(o, c) for o, c in zip(olist, cmds)])
        c_obs = np.stack([O.critic_obs_from_dict(o, c) for o, c in zip(olist, cmds)])
        ac = torch.zeros((2, 2, N, hidden), device=dev)
        cc = torch.zeros((2, 2, N, hidden), device=dev)
        lpf = torch.zeros((20, b.ld), device=dev)
        out = e.ppo_variables(synth.to_soa(a_obs, 1, dev), synth.to_soa(act.astype(np.float32), 1, dev),
                              synth.to_soa(done.astype(np.uint8), 1, dev), ac, lpf, synth.to_soa(c_obs, 1, dev), cc,
                              want_mean=want_mean, n_envs=N)
        return out, ac, cc, lpf

    out, ac, cc, lpf = run(obs_list, cmd, False)
    assert e.device_status() == 0
    close(S(out["log_probs"], N), ref["log_probs"][..., 0], "log_probs", atol=1e-4)     # |log_prob| up to ~1e3 here
    close(S(out["entropy"], N), ref["entropy"][..., 0], "entropy", atol=1e-5)
    close(S(out["values"], N), ref["values"], "values", atol=1e-5)
    close(S(out["action_std"], N, (20,)), ref["action_std"], "action_std")
    close(Hn.carry_to_np(ac, N), c_end["actor"], "actor carry")
    close(Hn.carry_to_np(cc, N), c_end["critic"], "critic carry", atol=1e-5)
    close(S(lpf, N, (20,)), c_end["lpf_params"], "lpf")
    # mirror pass (train.py:1463-1481): same entry point on mirror_obs / mirror_cmd
    m_list = [O.mirror_obs(o) for o in obs_list]
    m_cmd = np.stack([O.mirror_cmd(c) for c in cmd])
    mout, _, _, _ = run(m_list, m_cmd, True)
    close(S(mout["mean"], N, (20,)), ref["mirror_mean"], "mirrored dist.mean()")
    close(S(mout["values"], N), ref["mirror_value"], "mirrored value", atol=1e-5)
    # aux_losses as library outputs (train.py:1462-1481, 1488-1491): one call, the mirrored passes with their own carries
    a_obs = synth.to_soa(np.stack([O.actor_obs_from_dict(o, c) for o, c in zip(obs_list, cmd)]), 1, dev)
    c_obs = synth.to_soa(np.stack([O.critic_obs_from_dict(o, c) for o, c in zip(obs_list, cmd)]), 1, dev)
    am_obs = synth.to_soa(np.stack([O.actor_obs_from_dict(o, c) for o, c in zip(m_list, m_cmd)]), 1, dev)
    cm_obs = synth.to_soa(np.stack([O.critic_obs_from_dict(o, c) for o, c in zip(m_list, m_cmd)]), 1, dev)
    z = lambda: torch.zeros((2, 2, N, hidden), device=dev)
    mir = {"actor_obs": am_obs, "critic_obs": cm_obs, "actor_carry": z(), "critic_carry": z(),
           "lpf": torch.zeros((20, b.ld), device=dev), "actor_scale": p.actor_mirror_loss_scale,
           "critic_scale": p.critic_mirror_loss_scale}
    ac, cc, lpf = z(), z(), torch.zeros((20, b.ld), device=dev)
    full = e.ppo_variables(a_obs, synth.to_soa(act.astype(np.float32), 1, dev), synth.to_soa(done.astype(np.uint8), 1, dev),
                           ac, lpf, c_obs, cc, n_envs=N, mirror=mir)
    assert e.device_status() == 0
    close(S(full["log_probs"], N), ref["log_probs"][..., 0], "log_probs (aux call)", atol=1e-4)
    close(S(full["values"], N), ref["values"], "values (aux call)", atol=1e-5)
    close(S(full["action_mirror_loss"], N), ref["action_mirror_loss"], "action_mirror_loss", rtol=1e-4, atol=1e-6)
    close(S(full["value_mirror_loss"], N), ref["value_mirror_loss"], "value_mirror_loss", rtol=1e-4, atol=2e-7)
    assert float(np.abs(ref["action_mirror_loss"]).max()) > 1e-3 and float(np.abs(ref["value_mirror_loss"]).max()) > 1e-6
    close(Hn.carry_to_np(mir["actor_carry"], N), c_end["actor_mirror"], "actor_mirror carry")
    close(Hn.carry_to_np(mir["critic_carry"], N), c_end["critic_mirror"], "critic_mirror carry", atol=1e-5)
    close(S(mir["lpf"], N, (20,)), c_end["lpf_params_mirror"], "lpf_params_mirror")
    close(Hn.carry_to_np(ac, N), c_end["actor"], "actor carry (aux call)")
    e.close()


def test_task_plugin_surface_end_to_end(dev):
    """The reference-shaped host API (kbot-joystick_b200/task.py): observations -> sample_action -> actuators ->
    terminations through HumanoidWalkingTask, against the oracle's single-env functions."""
    from kbot_joystick_b200.task import HumanoidWalkingTask, HumanoidWalkingTaskConfig

    N = 96
    b = Batch(1100, 2, N, dev)
    task = HumanoidWalkingTask(HumanoidWalkingTaskConfig(num_envs=N))
    wa = synth.make_weights(77, 65, 40, 256, 2)
    wc = synth.make_weights(78, 475, 1, 256, 2)
    task.get_model(synth.weights_to_device(wa, dev), synth.weights_to_device(wc, dev))
    carry = task.get_initial_model_carry(N, dev)
    assert set(carry) == {"actor", "actor_mirror", "critic", "critic_mirror", "lpf_params", "lpf_params_mirror"}
    cmd = task.get_commands(torch.zeros((16, b.ld), device=dev),
                            {k: synth.to_soa(v, 0, dev) for k, v in b.np["cmd0_rand"].items()}, initial=True, n_envs=N)
    exact(S(cmd["unified_command"], N, (16,)), b.cmd0_np, "initial command")
    obs = task.get_observations(b.state_at(0), cmd, b.noise_at(0), b.episode, n_envs=N)
    o_np, _ = O.get_observations(b.np_state_at(0), b.np_noise_at(0), b.np["episode"], None, P)
    for name in ("joint_position", "noisy_biased_joint_position", "noisy_joint_velocity", "noisy_imu_gyro", "feet_position",
                 "projected_gravity", "noisy_imu_projected_gravity", "base_orientation", "center_of_mass_velocity",
                 "left_foot_touch", "base_height", "imu_gyro"):
        ref = o_np[name]
        close(S(obs[name], N, (ref.shape[-1],)), ref, name, atol=1e-5 if "gravity" in name else 1e-6)
    act = task.sample_action(carry, obs, eps=b.noise["eps_action"][0], argmax=False, n_envs=N)
    p = O.OracleParams()
    a_np, mean, std, _, _ = O.sample_action(wa, O.actor_obs_from_dict(o_np, b.cmd0_np), np.zeros((N, 2, 2, 256), np.float32),
                                            np.zeros((N, 20), np.float32), b.np["noise"]["eps_action"][0], False, p)
    close(S(act["action"], N, (20,)), a_np, "sample_action")
    val = task.run_critic(obs, carry["critic"], n_envs=N)
    v_np, _ = O.critic_forward(wc, O.critic_obs_from_dict(o_np, b.cmd0_np), np.zeros((N, 2, 2, 256), np.float32))
    close(S(val, N), v_np[:, 0], "run_critic", atol=1e-5)
    ctrl = task.get_actuators(act["action"], b.state_at(0), b.episode, n_envs=N)
    ep = b.np["episode"]
    st = b.np_state_at(0)
    close(S(ctrl, N, (20,)), O.position_actuator_torque(a_np, st["qpos"][:, 7:], st["qvel"][:, 6:], ep["kp"], ep["kd"],
                                                       ep["tau_limit"], ep["action_bias"], ep["torque_bias"]), "ctrl", atol=1e-4)
    term = task.get_terminations(b.state_at(0), n_envs=N)
    codes, done, succ = O.terminations(st["xpos"], st["qpos"][:, 3:7], st["time"], P)
    exact(S(term["codes"], N, (3,)), codes, "termination codes")
    task.close()
