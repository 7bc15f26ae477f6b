"""Call-compatible adapter for the reference's ksim Task.

The reference writes every hook for ONE environment and ksim vmaps it (train.py:1155, 1206, 1224, 1258, 1510, 1545).
`KbotFfiTaskMixin` overrides the model hooks of `train.HumanoidWalkingTask` with the SAME signatures; each is a
`jax.custom_batching.custom_vmap` function whose batching rule lowers the vmapped call to ONE XLA-FFI custom call on the
batched, env-major arrays (`jax_ffi.py` -> csrc/kbs_xla_ffi.cc -> include/kbotstep.h).  Un-vmapped calls take the same
path with a batch of one.

    import train                                         # the unmodified reference
    from kbot_joystick_b200 import ksim_adapter, jax_ffi
    jax_ffi.register()
    Task = ksim_adapter.make_task_class(train.HumanoidWalkingTask)
    Task.launch(train.HumanoidWalkingTaskConfig(...))     # as train.py:1759-1792 does

NOT RUN IN THIS IMAGE: jax / ksim are not installable here.  Importing this module needs neither; the CPU
tests check that every override has exactly the reference's parameter list (tests/golden/ref_signatures.json, extracted
from train.py by tools/make_ref_signatures.py) and that every FFI target it calls is registered by jax_ffi.py.
What stays in JAX: the physics (MJX), PRNG key handling (noise is drawn with jax.random exactly as the reference does and
handed to the kernels as explicit arrays: "identical inputs and PRNG-derived noise"), and ksim's runtime around the hooks.
"""
from __future__ import annotations

from . import jax_ffi as kf

NUM_JOINTS = 20
# FFI targets each override lowers to (checked against jax_ffi._TARGETS by the CPU tests)
HOOK_TARGETS = {
    "run_actor": ("kbs_actor_step",), "run_critic": ("kbs_critic_step",), "sample_action": ("kbs_actor_step",),
    "get_ppo_variables": ("kbs_mirror_observations", "kbs_ppo_variables"), "get_initial_model_carry": (),
}


def _round_up4(n: int) -> int:
    return (n + 3) // 4 * 4


def _soa(jnp, x, ld: int):
    """[N, F...] (ksim's vmapped AoS) -> env-major SoA [F, ld] (env axis last, zero-padded to ld)."""
    n = x.shape[0]
    y = jnp.moveaxis(x.reshape(n, -1), 0, -1)
    return jnp.pad(y, ((0, 0), (0, ld - n)))


def _aos(jnp, y, n: int, shape=()):
    """SoA [F, ld] -> [N, *shape]."""
    return jnp.moveaxis(y[..., :n], -1, 0).reshape((n,) + tuple(shape or (y.shape[0],)))


def _batched(single_fn):
    """single_fn(*arrays_with_leading_env_axis) -> pytree with leading env axis.  Returns a function of SINGLE-env arrays that
    vmaps to exactly one call of single_fn on the stacked arrays (custom_vmap), and calls it with N = 1 otherwise."""
    import jax

    @jax.custom_batching.custom_vmap
    def f(*args):
        out = single_fn(*(a[None] for a in args))
        return jax.tree_util.tree_map(lambda o: o[0], out)

    @f.def_vmap
    def _rule(axis_size, in_batched, *args):
        import jax.numpy as jnp

        args = [a if b else jnp.broadcast_to(a[None], (axis_size,) + a.shape) for a, b in zip(args, in_batched)]
        out = single_fn(*args)
        return out, jax.tree_util.tree_map(lambda _: True, out)

    return f


class KbotFfiTaskMixin:
    """Mix in FRONT of the reference Task: `class Task(KbotFfiTaskMixin, train.HumanoidWalkingTask)`.  Needs `kbs_handle`
    (engine.KbotStep.handle_address of a handle whose weights follow the model: `sync_weights`)."""

    kbs_handle: int = 0

    # ---- weights: equinox modules -> kbs_weights_pack (through ctypes: a host-side, once-per-update operation) ------------
    def sync_weights(self, model, engine) -> None:
        """Pack `model.actor` / `model.critic` (train.py:847-1004) into the handle.  eqx layout = the library's input layout."""
        import numpy as np
        import torch

        def net(m):
            t = lambda a: torch.from_numpy(np.asarray(a, np.float32)).cuda()      # noqa: E731
            return {"w_in": t(m.input_proj.weight), "b_in": t(m.input_proj.bias), "w_out": t(m.output_proj.weight),
                    "b_out": t(m.output_proj.bias),
                    "layers": [{"w_ih": t(r.weight_ih), "w_hh": t(r.weight_hh), "b": t(r.bias)} for r in m.rnns]}

        engine.pack_weights(0, net(model.actor))
        engine.pack_weights(1, net(model.critic))
        self.kbs_handle = engine.handle_address

    # ---- train.py:1526 -------------------------------------------------------------------------------------------------
    def get_initial_model_carry(self, model, rng):
        return super().get_initial_model_carry(model, rng)        # zeros (train.py:1526-1543): nothing to accelerate

    # ---- train.py:1351 -------------------------------------------------------------------------------------------------
    def run_actor(self, model, observations, commands, carry, lpf_params):
        """Same contract as train.py:1351-1379: returns (distrax.MultivariateNormalDiag, next carry, next lpf_params); the
        trunk + head run as one kbs_actor_step custom call (mean / std come back, the distribution object is rebuilt)."""
        import distrax
        import jax.numpy as jnp
        import ksim

        obs_n = self._actor_obs(observations, commands)                      # the reference's own concatenation (train.py:1360-1376)
        h = self.kbs_handle
        depth, H = self.config.depth, self.config.hidden_size

        def batched(obs, carry_s, lpf):
            n = obs.shape[0]
            ld = _round_up4(n)
            c = jnp.moveaxis(carry_s, 0, 2)                                 # [N, depth, 2, H] -> [depth, 2, N, H]
            zeros_u8 = jnp.zeros((ld,), jnp.uint8)
            c2, lpf2, _act, mean, std, _lp, _ent = kf.actor_step(h, _soa(jnp, obs, ld), c, _soa(jnp, lpf, ld),
                                                                 jnp.zeros((NUM_JOINTS, ld), jnp.float32), zeros_u8, n, argmax=True)
            return _aos(jnp, mean, n), _aos(jnp, std, n), jnp.moveaxis(c2, 2, 0), _aos(jnp, lpf2, n)

        carry_arr = jnp.stack([jnp.stack(hc) for hc in carry])               # tuple((h, c) per layer) -> [depth, 2, H]
        lpf_arr = self._lpf_array(lpf_params)
        mean, std, c2, lpf2 = _batched(batched)(obs_n, carry_arr, lpf_arr)
        next_carry = tuple((c2[i, 0], c2[i, 1]) for i in range(depth))
        return distrax.MultivariateNormalDiag(loc=mean, scale_diag=std), next_carry, self._lpf_params(lpf_params, lpf2)

    # ---- train.py:1381 -------------------------------------------------------------------------------------------------
    def run_critic(self, model, observations, commands, carry):
        import jax.numpy as jnp

        obs_n = self._critic_obs(observations, commands)                     # train.py:1388-1431
        h = self.kbs_handle
        depth = self.config.depth

        def batched(obs, carry_s):
            n = obs.shape[0]
            ld = _round_up4(n)
            c2, value = kf.critic_step(h, _soa(jnp, obs, ld), jnp.moveaxis(carry_s, 0, 2), jnp.zeros((ld,), jnp.uint8), n)
            return value[:n, None], jnp.moveaxis(c2, 2, 0)

        carry_arr = jnp.stack([jnp.stack(hc) for hc in carry])
        value, c2 = _batched(batched)(obs_n, carry_arr)
        return value, tuple((c2[i, 0], c2[i, 1]) for i in range(depth))

    # ---- train.py:1545 -------------------------------------------------------------------------------------------------
    def sample_action(self, model, model_carry, physics_model, physics_state, observations, commands, curriculum_level, rng,
                      argmax):
        import jax
        import ksim

        action_dist, next_actor_carry, next_lpf = self.run_actor(model=model.actor, observations=observations, commands=commands,
                                                                 carry=model_carry["actor"], lpf_params=model_carry["lpf_params"])
        action = action_dist.mode() if argmax else action_dist.sample(seed=rng)      # the reference's own PRNG draw (train.py:1564)
        next_carry = dict(model_carry)
        next_carry.update(actor=next_actor_carry, lpf_params=next_lpf)
        return ksim.Action(action=action, carry=next_carry)

    # ---- train.py:1510 -------------------------------------------------------------------------------------------------
    def get_ppo_variables(self, model, trajectory, model_carry, rng):
        """xax.scan(_ppo_scan_fn) over the stored trajectory (train.py:1435-1524) as ONE kbs_ppo_variables custom call per
        vmapped batch of trajectories: both networks, the mirrored passes, log-prob / entropy / std / aux losses, carries reset
        where done.  The trajectory's observations are concatenated by the reference's own run_actor / run_critic recipes."""
        import jax
        import jax.numpy as jnp
        import ksim

        h = self.kbs_handle
        depth = self.config.depth
        a_obs = jax.vmap(self._actor_obs)(trajectory.obs, trajectory.command)                  # [T, 65]
        c_obs = jax.vmap(self._critic_obs)(trajectory.obs, trajectory.command)                 # [T, 475]
        m_obs, m_cmd = jax.vmap(self.mirror_obs)(trajectory.obs), jax.vmap(self.mirror_cmd)(trajectory.command)
        am_obs, cm_obs = jax.vmap(self._actor_obs)(m_obs, m_cmd), jax.vmap(self._critic_obs)(m_obs, m_cmd)
        sa, sc = self.config.actor_mirror_loss_scale, self.config.critic_mirror_loss_scale
        stack = lambda c: jnp.stack([jnp.stack(hc) for hc in c])             # noqa: E731

        def batched(a_o, c_o, am_o, cm_o, act, done, ca, cc, lpf, cam, ccm, lpfm):
            n, T = a_o.shape[0], a_o.shape[1]
            ld = _round_up4(n)
            tm = lambda x: jnp.pad(jnp.moveaxis(x, 0, -1), ((0, 0),) * (x.ndim - 1) + ((0, ld - n),))   # [N, T, F] -> [T, F, ld]
            cr = lambda c: jnp.moveaxis(c, 0, 2)
            res = kf.ppo_variables(h, tm(a_o), tm(c_o), tm(am_o), tm(cm_o), tm(act), tm(done.astype(jnp.uint8)), cr(ca), cr(cc),
                                   _soa(jnp, lpf, ld), cr(cam), cr(ccm), _soa(jnp, lpfm, ld), n, sa, sc)
            ca2, cc2, lpf2, cam2, ccm2, lpfm2, lp, val, ent, std, aml, vml = res
            back = lambda y: jnp.moveaxis(y[..., :n], -1, 0)                 # [T, ..., ld] -> [N, T, ...]
            uncr = lambda c: jnp.moveaxis(c, 2, 0)
            return (back(lp)[..., None], back(val), back(ent)[..., None], back(std), back(aml), back(vml), uncr(ca2), uncr(cc2),
                    _aos(jnp, lpf2, n), uncr(cam2), uncr(ccm2), _aos(jnp, lpfm2, n))

        out = _batched(batched)(a_obs, c_obs, am_obs, cm_obs, trajectory.action, trajectory.done, stack(model_carry["actor"]),
                                stack(model_carry["critic"]), self._lpf_array(model_carry["lpf_params"]),
                                stack(model_carry["actor_mirror"]), stack(model_carry["critic_mirror"]),
                                self._lpf_array(model_carry["lpf_params_mirror"]))
        lp, val, ent, std, aml, vml, ca2, cc2, lpf2, cam2, ccm2, lpfm2 = out
        tup = lambda c: tuple((c[i, 0], c[i, 1]) for i in range(depth))     # noqa: E731
        ppo_variables = ksim.PPOVariables(log_probs=lp, values=val, entropy=ent, action_std=std,
                                          aux_losses={"action_mirror_loss": aml, "value_mirror_loss": vml})
        next_carry = {"actor": tup(ca2), "critic": tup(cc2), "actor_mirror": tup(cam2), "critic_mirror": tup(ccm2),
                      "lpf_params": self._lpf_params(model_carry["lpf_params"], lpf2),
                      "lpf_params_mirror": self._lpf_params(model_carry["lpf_params_mirror"], lpfm2)}
        return ppo_variables, next_carry

    # ---- helpers: the reference's own observation recipes, factored out of run_actor / run_critic -------------------------
    def _actor_obs(self, observations, commands):
        """train.py:1360-1376 (identical in convert.py:95-105)."""
        import jax.numpy as jnp

        o = observations
        cmd = commands["unified_command"]
        zero_cmd = (jnp.linalg.norm(cmd[..., :3], axis=-1) < 1e-3)[..., None]
        return jnp.concatenate([self.normalize_joint_pos(o["noisy_biased_joint_position"]),
                                self.normalize_joint_vel(o["noisy_joint_velocity"]),
                                self.encode_projected_gravity(o["noisy_imu_projected_gravity"]), o["noisy_imu_gyro"], zero_cmd, cmd],
                               axis=-1)

    def _critic_obs(self, observations, commands):
        """train.py:1388-1431."""
        import jax.numpy as jnp

        o = observations
        cmd = commands["unified_command"]
        zero_cmd = (jnp.linalg.norm(cmd[..., :3], axis=-1) < 1e-3)[..., None]
        return jnp.concatenate([self.normalize_joint_pos(o["joint_position"]), self.normalize_joint_vel(o["joint_velocity"]),
                                self.encode_projected_gravity(o["projected_gravity"]), o["imu_gyro"], zero_cmd, cmd,
                                o["left_foot_touch"], o["right_foot_touch"], o["feet_position"], o["base_position"],
                                o["base_orientation"], o["center_of_mass_inertia"], o["center_of_mass_velocity"],
                                o["base_linear_velocity"], o["base_angular_velocity"], o["actuator_force"] / 4.0, o["base_height"]],
                               axis=-1)

    @staticmethod
    def _lpf_array(lpf_params):
        import jax

        return jax.tree_util.tree_leaves(lpf_params)[0]                      # LowPassFilterParams: exactly 20 floats (convert.py:71)

    @staticmethod
    def _lpf_params(template, new):
        import jax

        leaves, treedef = jax.tree_util.tree_flatten(template)
        return jax.tree_util.tree_unflatten(treedef, [new] + leaves[1:])


def make_task_class(reference_task_cls):
    """`class Task(KbotFfiTaskMixin, reference_task_cls)`: the reference Task with its model hooks lowered to the library."""
    return type("KbotFfi" + reference_task_cls.__name__, (KbotFfiTaskMixin, reference_task_cls), {})
