"""jax.ffi registration of the XLA-FFI handlers in libkbs_xla_ffi.so (csrc/kbs_xla_ffi.cc): the binding a JAX program
such as the reference (train.py on ksim) uses to call the C-ABI from inside jit on the CUDA platform.

Not importable in this image (no jax / jaxlib: SURVEY F5) and not used by the tests or the bench, which bind the same
C-ABI through ctypes + torch (`_lib.py`, `engine.py`).  Build the shim with `make ffi JAX_INCLUDE=...` first.

    import kbot_joystick_b200.jax_ffi as kf
    kf.register()                                   # once per process
    h = engine.handle_address                       # kbs_create / kbs_weights_pack done through ctypes as usual
    adv, targets = kf.gae(h, values_tn, rewards_tn, done_tn, success_tn, n_envs=N)
"""
from __future__ import annotations

import ctypes
from pathlib import Path

_TARGETS = {"kbs_gae": "KbsGae", "kbs_terminate": "KbsTerminate", "kbs_torque": "KbsTorque",
            "kbs_policy_step": "KbsPolicyStep", "kbs_ppo_variables": "KbsPpoVariables", "kbs_rewards": "KbsRewards"}


def register(path: str | None = None) -> None:
    import jax  # noqa: PLC0415 -- deliberately lazy: this module must import without jax

    lib = ctypes.cdll.LoadLibrary(path or str(Path(__file__).resolve().parent / "libkbs_xla_ffi.so"))
    for target, symbol in _TARGETS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")


def gae(handle: int, values, rewards, done, success, n_envs: int):
    """ksim.compute_ppo_inputs: values / rewards [T, ld] f32, done / success [T, ld] -> (advantages, value_targets)."""
    import jax
    import jax.numpy as jnp
    import numpy as np

    out = (jax.ShapeDtypeStruct(values.shape, jnp.float32),) * 2
    return jax.ffi.ffi_call("kbs_gae", out)(values, rewards, done.astype(jnp.uint8), success.astype(jnp.uint8),
                                            handle=np.int64(handle), n_envs=np.int64(n_envs))


def policy_step(handle: int, joint_angles, joint_vel, projected_gravity, gyro, command, carry):
    """convert.py step_fn batched over envs: AoS [n, .] inputs + flat carry [n, depth*2*H + 20] -> (action [n, 20], carry)."""
    import jax
    import jax.numpy as jnp
    import numpy as np

    n = joint_angles.shape[0]
    out = (jax.ShapeDtypeStruct((n, 20), jnp.float32), jax.ShapeDtypeStruct(carry.shape, jnp.float32))
    return jax.ffi.ffi_call("kbs_policy_step", out)(joint_angles, joint_vel, projected_gravity, gyro, command, carry,
                                                    handle=np.int64(handle))


def ppo_variables(handle: int, actor_obs, critic_obs, action, done, actor_carry, critic_carry, lpf, n_envs: int):
    """get_ppo_variables on a stored trajectory ([T, F, ld] SoA): -> (actor_carry, critic_carry, lpf, log_probs [T, ld],
    values [T, ld], entropy [T, ld], action_std [T, 20, ld]); the carries alias their inputs."""
    import jax
    import jax.numpy as jnp
    import numpy as np

    T, _, ld = actor_obs.shape
    f32 = jnp.float32
    out = (jax.ShapeDtypeStruct(actor_carry.shape, f32), jax.ShapeDtypeStruct(critic_carry.shape, f32),
           jax.ShapeDtypeStruct(lpf.shape, f32), jax.ShapeDtypeStruct((T, ld), f32), jax.ShapeDtypeStruct((T, ld), f32),
           jax.ShapeDtypeStruct((T, ld), f32), jax.ShapeDtypeStruct((T, 20, ld), f32))
    call = jax.ffi.ffi_call("kbs_ppo_variables", out, input_output_aliases={4: 0, 5: 1, 6: 2})
    return call(actor_obs, critic_obs, action, done.astype(jnp.uint8), actor_carry, critic_carry, lpf,
                handle=np.int64(handle), n_envs=np.int64(n_envs))
