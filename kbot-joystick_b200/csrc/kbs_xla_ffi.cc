// kbs_xla_ffi.cc -- XLA FFI (jax.ffi) handlers over the C-ABI of libkbotstep.so: the binding a maintainer of the reference
// (a JAX program) adds so that the jitted ksim Task hooks lower to these kernels on the CUDA platform.
//
// NOT BUILT IN THIS IMAGE: jaxlib and its headers (xla/ffi/api/ffi.h) are not installable here, so this file
// is compiled only where they exist:
//     make ffi JAX_INCLUDE=$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
// It adds nothing to the product path: every handler unpacks buffers and forwards to ONE entry point of include/kbotstep.h
// on XLA's stream; errors come back as ffi::Error (no throw across the ABI, no CPU fallback).
//
// Handle: the Python side creates the kbs_handle once (kbs_create / kbs_weights_pack through ctypes, jax_ffi.py) and
// passes the pointer as the int64 attribute "handle" of every call.  Layout at the boundary: env-major SoA [F][ld]
// (ld = n_envs rounded up to 4), trajectories [T][F][ld] -- see INTEGRATION.md "Layout contract".
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>

#include "kbotstep.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

inline kbs_handle* H(int64_t handle) { return reinterpret_cast<kbs_handle*>(static_cast<intptr_t>(handle)); }
inline ffi::Error Rc(int rc) { return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(kbs_error_string(rc)); }
template <typename B>
inline int64_t Dim(const B& b, int i) { return static_cast<int64_t>(b.dimensions()[i]); }

// ---- ksim.compute_ppo_inputs (GAE), gamma / lam live in the handle's kbs_params (train.py:1769-1770) --------------------
ffi::Error GaeImpl(cudaStream_t stream, int64_t handle, int64_t n_envs, ffi::Buffer<ffi::F32> values,
                   ffi::Buffer<ffi::F32> rewards, ffi::Buffer<ffi::U8> done, ffi::Buffer<ffi::U8> success,
                   ffi::ResultBuffer<ffi::F32> adv, ffi::ResultBuffer<ffi::F32> targets) {
  const int64_t T = Dim(values, 0), ld = Dim(values, 1);
  return Rc(kbs_gae(H(handle), values.typed_data(), rewards.typed_data(), done.typed_data(), success.typed_data(),
                    adv->typed_data(), targets->typed_data(), T, ld, n_envs, stream));
}

// ---- get_terminations (train.py:1258-1269, 817-823): qpos [27][ld], xpos [72][ld], time [ld] -> codes / done / success ----
ffi::Error TerminateImpl(cudaStream_t stream, int64_t handle, int64_t n_envs, ffi::Buffer<ffi::F32> qpos,
                         ffi::Buffer<ffi::F32> xpos, ffi::Buffer<ffi::F32> time, ffi::ResultBuffer<ffi::S32> codes,
                         ffi::ResultBuffer<ffi::U8> done, ffi::ResultBuffer<ffi::U8> success) {
  kbs_state_view s{};
  s.qpos = qpos.typed_data(); s.xpos = xpos.typed_data(); s.time = time.typed_data(); s.ld = Dim(qpos, 1);
  return Rc(kbs_terminate(H(handle), &s, codes->typed_data(), done->typed_data(), success->typed_data(), nullptr, n_envs,
                          stream));
}

// ---- get_actuators().get_ctrl (train.py:1091-1105): action [20][ld], qpos [27][ld], qvel [26][ld] -> ctrl [20][ld] -------
ffi::Error TorqueImpl(cudaStream_t stream, int64_t handle, int64_t n_envs, ffi::Buffer<ffi::F32> action,
                      ffi::Buffer<ffi::F32> qpos, ffi::Buffer<ffi::F32> qvel, ffi::ResultBuffer<ffi::F32> ctrl) {
  kbs_state_view s{};
  s.qpos = qpos.typed_data(); s.qvel = qvel.typed_data(); s.ld = Dim(qpos, 1);
  return Rc(kbs_torque(H(handle), action.typed_data(), &s, nullptr, ctrl->typed_data(), n_envs, stream));
}

// ---- convert.py:84-119 step_fn, batched over envs (AoS rows as the exported function takes them) -------------------------
ffi::Error PolicyStepImpl(cudaStream_t stream, int64_t handle, ffi::Buffer<ffi::F32> joint_angles,
                          ffi::Buffer<ffi::F32> joint_vel, ffi::Buffer<ffi::F32> projected_gravity, ffi::Buffer<ffi::F32> gyro,
                          ffi::Buffer<ffi::F32> command, ffi::Buffer<ffi::F32> carry, ffi::ResultBuffer<ffi::F32> action,
                          ffi::ResultBuffer<ffi::F32> carry_out) {
  return Rc(kbs_policy_step(H(handle), joint_angles.typed_data(), joint_vel.typed_data(), projected_gravity.typed_data(),
                            gyro.typed_data(), command.typed_data(), carry.typed_data(), carry_out->typed_data(),
                            action->typed_data(), Dim(joint_angles, 0), stream));
}

// ---- get_ppo_variables -> xax.scan(_ppo_scan_fn) (train.py:1435-1524) on a stored trajectory ------------------------------
// carries / lpf are in-out in the C-ABI: XLA aliases the inputs onto the results (input_output_aliases in jax_ffi.py).
ffi::Error PpoVariablesImpl(cudaStream_t stream, int64_t handle, int64_t n_envs, ffi::Buffer<ffi::F32> actor_obs,
                            ffi::Buffer<ffi::F32> critic_obs, ffi::Buffer<ffi::F32> action, ffi::Buffer<ffi::U8> done,
                            ffi::Buffer<ffi::F32> actor_carry, ffi::Buffer<ffi::F32> critic_carry, ffi::Buffer<ffi::F32> lpf,
                            ffi::ResultBuffer<ffi::F32> actor_carry_out, ffi::ResultBuffer<ffi::F32> critic_carry_out,
                            ffi::ResultBuffer<ffi::F32> lpf_out, ffi::ResultBuffer<ffi::F32> log_probs,
                            ffi::ResultBuffer<ffi::F32> values, ffi::ResultBuffer<ffi::F32> entropy,
                            ffi::ResultBuffer<ffi::F32> action_std) {
  // aliased in/out pairs share storage; if XLA did not alias them the caller's copies are taken first
  auto same = [&](const void* a, const void* b, size_t bytes) {
    return a == b ? cudaSuccess : cudaMemcpyAsync(const_cast<void*>(b), a, bytes, cudaMemcpyDeviceToDevice, stream);
  };
  if (same(actor_carry.typed_data(), actor_carry_out->typed_data(), actor_carry.size_bytes()) != cudaSuccess ||
      same(critic_carry.typed_data(), critic_carry_out->typed_data(), critic_carry.size_bytes()) != cudaSuccess ||
      same(lpf.typed_data(), lpf_out->typed_data(), lpf.size_bytes()) != cudaSuccess)
    return ffi::Error::Internal("kbs_ppo_variables: carry copy failed");
  kbs_ppo_io io{};
  io.actor_obs = actor_obs.typed_data(); io.critic_obs = critic_obs.typed_data(); io.action = action.typed_data();
  io.done = done.typed_data();
  io.actor_carry = actor_carry_out->typed_data(); io.critic_carry = critic_carry_out->typed_data(); io.lpf = lpf_out->typed_data();
  io.log_probs = log_probs->typed_data(); io.values = values->typed_data(); io.entropy = entropy->typed_data();
  io.action_std = action_std->typed_data(); io.mean = nullptr;
  io.T = Dim(actor_obs, 0); io.ld = Dim(actor_obs, 2);
  return Rc(kbs_ppo_variables(H(handle), &io, n_envs, stream));
}

// ---- get_rewards table + scale-and-sum (train.py:125-506, 1224-1256) on a trajectory --------------------------------------
ffi::Error RewardsImpl(cudaStream_t stream, int64_t handle, int64_t n_envs, ffi::Buffer<ffi::F32> qpos, ffi::Buffer<ffi::F32> qvel,
                       ffi::Buffer<ffi::F32> sensordata, ffi::Buffer<ffi::F32> xpos, ffi::Buffer<ffi::F32> xquat,
                       ffi::Buffer<ffi::F32> com_distance, ffi::Buffer<ffi::F32> command, ffi::Buffer<ffi::F32> ctrl,
                       ffi::Buffer<ffi::U8> done, ffi::Buffer<ffi::F32> t_single, ffi::Buffer<ffi::F32> airtime,
                       ffi::Buffer<ffi::U8> prev_contact, ffi::ResultBuffer<ffi::F32> t_single_out,
                       ffi::ResultBuffer<ffi::F32> airtime_out, ffi::ResultBuffer<ffi::U8> prev_contact_out,
                       ffi::ResultBuffer<ffi::F32> total, ffi::ResultBuffer<ffi::F32> components) {
  auto same = [&](const void* a, const void* b, size_t bytes) {
    return a == b ? cudaSuccess : cudaMemcpyAsync(const_cast<void*>(b), a, bytes, cudaMemcpyDeviceToDevice, stream);
  };
  if (same(t_single.typed_data(), t_single_out->typed_data(), t_single.size_bytes()) != cudaSuccess ||
      same(airtime.typed_data(), airtime_out->typed_data(), airtime.size_bytes()) != cudaSuccess ||
      same(prev_contact.typed_data(), prev_contact_out->typed_data(), prev_contact.size_bytes()) != cudaSuccess)
    return ffi::Error::Internal("kbs_rewards: carry copy failed");
  kbs_traj_view tr{};
  tr.state.qpos = qpos.typed_data(); tr.state.qvel = qvel.typed_data(); tr.state.sensordata = sensordata.typed_data();
  tr.state.xpos = xpos.typed_data(); tr.state.xquat = xquat.typed_data(); tr.state.com_distance = com_distance.typed_data();
  tr.state.ld = Dim(qpos, 2);
  tr.command = command.typed_data(); tr.ctrl = ctrl.typed_data(); tr.done = done.typed_data(); tr.T = Dim(qpos, 0);
  kbs_reward_carry rc{t_single_out->typed_data(), airtime_out->typed_data(), prev_contact_out->typed_data()};
  return Rc(kbs_rewards(H(handle), &tr, &rc, total->typed_data(), components->typed_data(), n_envs, stream));
}

}  // namespace

#define KBS_F32 ffi::Buffer<ffi::F32>
#define KBS_U8 ffi::Buffer<ffi::U8>
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsGae, GaeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle").Attr<int64_t>("n_envs")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_U8>().Arg<KBS_U8>().Ret<KBS_F32>().Ret<KBS_F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsTerminate, TerminateImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle").Attr<int64_t>("n_envs")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Ret<ffi::Buffer<ffi::S32>>().Ret<KBS_U8>().Ret<KBS_U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsTorque, TorqueImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle").Attr<int64_t>("n_envs")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Ret<KBS_F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsPolicyStep, PolicyStepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>()
                                  .Ret<KBS_F32>().Ret<KBS_F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsPpoVariables, PpoVariablesImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle").Attr<int64_t>("n_envs")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_U8>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>()
                                  .Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(KbsRewards, RewardsImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("handle").Attr<int64_t>("n_envs")
                                  .Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_F32>()
                                  .Arg<KBS_F32>().Arg<KBS_U8>().Arg<KBS_F32>().Arg<KBS_F32>().Arg<KBS_U8>()
                                  .Ret<KBS_F32>().Ret<KBS_F32>().Ret<KBS_U8>().Ret<KBS_F32>().Ret<KBS_F32>());
#else
#error "kbs_xla_ffi.cc needs jaxlib's XLA FFI headers: make ffi JAX_INCLUDE=$(python -c 'import jax.ffi; print(jax.ffi.include_dir())')"
#endif
