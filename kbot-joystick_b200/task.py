"""Host-side mirror of the reference's ksim Task plugin surface for the rollout control step.

`HumanoidWalkingTask` exposes the method names of `train.py:1058-1756` (get_observations / get_commands /
get_rewards / get_terminations / get_actuators / get_model / get_initial_model_carry / sample_action /
get_ppo_variables / run_actor / run_critic) with the same argument meaning, but BATCHED over environments: where the
reference is written for one env and ksim vmaps it, every array here carries the env axis last (SoA `[F, ld]`,
trajectories `[T, F, ld]`) and the vmap lives inside the CUDA kernels of libkbotstep.so.  Nothing is computed in
Python: each method marshals torch CUDA tensors into one C-ABI call (`engine.KbotStep`).  There is no CPU fallback.

What maps to what (reference -> here):
  ksim.PhysicsData / Trajectory fields  -> dict of SoA tensors  (engine.STATE_ROWS)
  xax.FrozenDict observations           -> dict name -> row-slice view of the state or of the `computed` block
  distrax.MultivariateNormalDiag        -> dict(mean, std) + the sampled / evaluated quantities the Task reads from it
  Carry TypedDict (train.py:1049-1055)  -> dict with the same keys; LSTM carries are `[depth, 2, n, H]`
  PRNG keys                             -> explicit noise tensors (parity mode: "identical inputs and PRNG-derived noise")
"""

from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib as L
from . import spec
from .engine import COMPUTED_OBS_ROWS, KbotStep, round_up4


@dataclass
class HumanoidWalkingTaskConfig:
    """The fields of train.py:73-122 / 1761-1791 the control step depends on."""

    hidden_size: int = 256          # train.py:1773
    depth: int = 2                  # train.py:82
    num_envs: int = 4096            # train.py:1763
    rollout_length_seconds: float = 2.0   # train.py:1766
    ctrl_dt: float = 0.02           # train.py:1776
    gamma: float = 0.94             # train.py:1769
    lam: float = 0.94               # train.py:1770
    var_scale: float = 0.5          # train.py:86
    gemm_path: int = L.GEMM_TC_2XF16

    @property
    def rollout_steps(self) -> int:
        return int(round(self.rollout_length_seconds / self.ctrl_dt))


class HumanoidWalkingTask:
    """train.py:1058 `HumanoidWalkingTask(ksim.PPOTask)` -- the hooks on the rollout control step, batched."""

    def __init__(self, config: HumanoidWalkingTaskConfig | None = None):
        self.config = config or HumanoidWalkingTaskConfig()
        c = self.config
        self.engine = KbotStep(hidden_size=c.hidden_size, depth=c.depth, gemm_path=c.gemm_path,
                               overrides={"gamma": c.gamma, "lam": c.lam, "var_scale": c.var_scale, "ctrl_dt": c.ctrl_dt,
                                          "switch_prob": c.ctrl_dt / 5})

    # ---- model ------------------------------------------------------------------------------------------------------
    def get_model(self, actor_weights: dict, critic_weights: dict) -> "HumanoidWalkingTask":
        """train.py:1278-1327.  Weights in equinox layout (Linear.weight [out,in]; LSTMCell weight_ih/hh [4H,H], bias
        [4H], gates i,f,g,o): dict(w_in, b_in, w_out, b_out, layers=[dict(w_ih, w_hh, b)]) of CUDA tensors."""
        assert actor_weights["w_in"].shape == (self.config.hidden_size, spec.ACTOR_OBS)     # train.py:1290-1295
        assert critic_weights["w_in"].shape == (self.config.hidden_size, spec.CRITIC_OBS)   # train.py:1297-1312
        self.engine.pack_weights(L.NET_ACTOR, actor_weights)
        self.engine.pack_weights(L.NET_CRITIC, critic_weights)
        return self

    def get_initial_model_carry(self, n_envs: int, device) -> dict:
        """train.py:1526-1543: zeros for 4 nets x depth x (h, c) and the two low-pass filter states."""
        c = self.config
        z = lambda: torch.zeros((c.depth, 2, n_envs, c.hidden_size), device=device)   # noqa: E731
        ld = round_up4(n_envs)
        return {"actor": z(), "actor_mirror": z(), "critic": z(), "critic_mirror": z(),
                "lpf_params": torch.zeros((spec.NUM_JOINTS, ld), device=device),
                "lpf_params_mirror": torch.zeros((spec.NUM_JOINTS, ld), device=device)}

    # ---- observations / commands ---------------------------------------------------------------------------------------
    def get_observations(self, state: dict, commands: dict, noise: dict | None = None, episode: dict | None = None,
                         obs_carry: dict | None = None, n_envs: int | None = None) -> dict:
        """train.py:1155-1204: the 21 named observations + the 4 `noisy_*` twins.  Pure slices of the physics state are
        returned as views; computed ones come from one `kbs_observations` call, which also writes the run_actor /
        run_critic concatenations (keys `actor_obs`, `critic_obs`)."""
        ld = state["qpos"].shape[-1]
        dev = state["qpos"].device
        comp = torch.empty((78, ld), device=dev)
        aobs = torch.empty((spec.ACTOR_OBS, ld), device=dev)
        cobs = torch.empty((spec.CRITIC_OBS, ld), device=dev)
        oc = obs_carry or {}
        self.engine.observations(state, commands["unified_command"], noise, episode, oc.get("pg_carry"), comp, aobs, cobs,
                                 n_envs, pg_reset=oc.get("pg_reset"))
        o = {name: comp[a:b] for name, (a, b) in COMPUTED_OBS_ROWS.items()}
        qpos, qvel, sd = state["qpos"], state["qvel"], state["sensordata"]
        o.update({
            "joint_position": qpos[7:], "joint_velocity": qvel[6:], "actuator_force": state["actuator_force"],
            "center_of_mass_inertia": state["cinert"][10:], "center_of_mass_velocity": state["cvel"][6:],
            "base_position": qpos[0:3], "base_orientation": qpos[3:7], "base_linear_velocity": qvel[0:3],
            "base_angular_velocity": qvel[3:6], "imu_gyro": sd[spec.SD_GYRO:spec.SD_GYRO + 3],
            "left_foot_touch": sd[spec.SD_TOUCH_L:spec.SD_TOUCH_L + 1],
            "right_foot_touch": sd[spec.SD_TOUCH_R:spec.SD_TOUCH_R + 1],
            "base_height": state["xpos"][3 * spec.BODY_BASE + 2:3 * spec.BODY_BASE + 3],
            "com_distance": state["com_distance"], "actor_obs": aobs, "critic_obs": cobs})
        if "qacc" in state:
            o.update({"base_linear_acceleration": state["qacc"][0:3], "base_angular_acceleration": state["qacc"][3:6],
                      "actuator_acceleration": state["qacc"][6:]})
        return o

    def get_commands(self, prev_command, rand: dict, initial: bool = False, n_envs: int | None = None) -> dict:
        """train.py:1206-1222 + UnifiedCommand (724-785).  rand: mode int32 [ld], u6 [6,ld], u_arms [10,ld], u_switch [ld]
        (ignored when `initial`): the explicit form of the 9-way key split.  Updates `prev_command` in place."""
        self.engine.command_update(prev_command, rand["mode"], rand["u6"], rand["u_arms"],
                                   None if initial else rand["u_switch"], n_envs)
        return {"unified_command": prev_command}

    # ---- networks ----------------------------------------------------------------------------------------------------------
    def run_actor(self, observations: dict, carry, lpf_params, eps=None, action_in=None, done=None, n_envs=None) -> dict:
        """train.py:1351-1379: returns the quantities read off the distrax distribution (mean/std/sample/log-prob/entropy);
        `carry` and `lpf_params` are updated in place."""
        return self.engine.actor_step(observations["actor_obs"], carry, lpf_params, eps=eps, action_in=action_in, done=done,
                                      n_envs=n_envs)

    def run_critic(self, observations: dict, carry, done=None, n_envs=None):
        """train.py:1381-1433."""
        return self.engine.critic_step(observations["critic_obs"], carry, done=done, n_envs=n_envs)

    def sample_action(self, model_carry: dict, observations: dict, eps=None, argmax: bool = False, done=None,
                      n_envs: int | None = None) -> dict:
        """train.py:1545-1572: action = dist.mode() if argmax else dist.sample(); only `actor` and `lpf_params` of the
        carry advance."""
        out = self.run_actor(observations, model_carry["actor"], model_carry["lpf_params"], eps=None if argmax else eps,
                             done=done, n_envs=n_envs)
        return {"action": out["action"], "carry": model_carry, "log_prob": out["log_prob"]}

    def get_ppo_variables(self, trajectory: dict, model_carry: dict, mirror: dict | None = None,
                          n_envs: int | None = None) -> tuple[dict, dict]:
        """train.py:1510-1524.  trajectory: actor_obs [T,65,ld], critic_obs [T,475,ld], action [T,20,ld], done [T,ld].
        mirror: the same keys for mirror_obs / mirror_cmd (train.py:1463-1481); when given, the aux losses are returned."""
        e = self.engine
        out = e.ppo_variables(trajectory["actor_obs"], trajectory["action"], trajectory["done"], model_carry["actor"],
                              model_carry["lpf_params"], trajectory.get("critic_obs"), model_carry.get("critic"),
                              want_mean=mirror is not None, n_envs=n_envs)
        ppo = {"log_probs": out["log_probs"].unsqueeze(1), "values": out["values"], "entropy": out["entropy"].unsqueeze(1),
               "action_std": out["action_std"], "aux_losses": {}}
        if mirror is not None:
            m = e.ppo_variables(mirror["actor_obs"], trajectory["action"], trajectory["done"], model_carry["actor_mirror"],
                                model_carry["lpf_params_mirror"], mirror.get("critic_obs"), model_carry.get("critic_mirror"),
                                want_std=False, want_mean=True, n_envs=n_envs)
            dm = e.mirror_joints(m["mean"].contiguous(), n_envs=n_envs)
            ppo["aux_losses"] = {"action_mirror_loss": ((out["mean"] - dm) ** 2).mean(dim=1),
                                 "value_mirror_loss": (out["values"] - m["values"]) ** 2}
        return ppo, model_carry

    def com_distance(self, contact: dict, subtree_com_base, n_envs: int | None = None):
        """COMDistanceObservation.observe (train.py:509-659) for T steps.  contact: geom1 / geom2 int32 [T, ncon, ld],
        pos [T, 3 ncon, ld]; subtree_com_base [T, 3, ld] = data.subtree_com[2].  -> [T, ld] (the `com_distance` state row)."""
        return self.engine.com_distance(contact["geom1"], contact["geom2"], contact["pos"], subtree_com_base, n_envs=n_envs)

    def mirror_joints(self, j, n_envs: int | None = None):
        """train.py:1574-1582 on `[T, 20, ld]`: negate all, swap the LEG halves only (as written)."""
        return self.engine.mirror_joints(j.contiguous(), n_envs=n_envs)

    def mirror_obs(self, trajectory_state: dict, computed, command, n_envs: int | None = None) -> dict:
        """mirror_obs + mirror_cmd (train.py:1584-1756) -> the `mirror` argument of get_ppo_variables.
        computed: [T, 78, ld] raw observations stored at rollout time (kbs_observations `computed`)."""
        T, ld = computed.shape[0], computed.shape[-1]
        out = {"actor_obs": torch.empty((T, spec.ACTOR_OBS, ld), device=computed.device),
               "critic_obs": torch.empty((T, spec.CRITIC_OBS, ld), device=computed.device),
               "command": torch.empty((T, spec.NUM_COMMANDS, ld), device=computed.device)}
        self.engine.mirror_observations(trajectory_state, computed, command, out["actor_obs"], out["critic_obs"],
                                        out["command"], n_envs=n_envs)
        return out

    # ---- actuators / terminations / rewards / GAE -----------------------------------------------------------------------
    def get_actuators(self, action, state: dict, episode: dict | None = None, n_envs: int | None = None):
        """train.py:1091-1105 ksim.PositionActuators.get_ctrl."""
        return self.engine.torque(action, state, episode, n_envs=n_envs)

    def get_terminations(self, state: dict, n_envs: int | None = None) -> dict:
        """train.py:1258-1269: codes rows = bad_z, not_upright, episode_length; done / success reduced ksim-style."""
        return self.engine.terminate(state, n_envs)

    def get_rewards(self, trajectory_state: dict, command, ctrl, done, reward_carry: dict, components=None,
                    n_envs: int | None = None):
        """train.py:1224-1256: sum_k scale_k * reward_k over the 12 terms, trajectory-wise (time leading)."""
        return self.engine.rewards(trajectory_state, command, ctrl, done, reward_carry, components=components, n_envs=n_envs)

    def compute_ppo_inputs(self, values, rewards, done, success, n_envs: int | None = None):
        """ksim.compute_ppo_inputs (GAE), gamma / lam from the config (train.py:1769-1770)."""
        return self.engine.gae(values, rewards, done, success, n_envs=n_envs)

    def compute_ppo_loss(self, ppo_variables: dict, old_ppo_variables: dict, advantages, value_targets,
                         n_envs: int | None = None, **hyper):
        """ksim.compute_ppo_loss (forward): clipped surrogate + value loss + entropy bonus, entropy_coef = 0.004
        (train.py:1767).  *_variables: the dicts get_ppo_variables returns ([T, 1, ld] log_probs / entropy, [T, ld] values).
        -> device tensor [loss, mean policy objective, mean value objective, mean entropy]."""
        sq = lambda x: x.squeeze(1).contiguous() if x.dim() == 3 else x
        return self.engine.ppo_loss(sq(ppo_variables["log_probs"]), sq(old_ppo_variables["log_probs"]), advantages,
                                    ppo_variables["values"], old_ppo_variables["values"], value_targets,
                                    sq(ppo_variables["entropy"]), n_envs=n_envs, **hyper)

    def rollout(self, io: dict, n_envs: int) -> None:
        """The fused control step over T recorded steps (ksim step_engine around mjx.step)."""
        self.engine.rollout(io, n_envs)

    def close(self) -> None:
        self.engine.close()
