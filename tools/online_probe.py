"""Per-kernel CUDA-event breakdown of the online form: one kbs_rollout call per control step (T = 1), 4 096 envs."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import kbot_joystick_b200  # noqa: F401
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
dev = torch.device("cuda:0")
H, N, T = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 1
eng = KbotStep(hidden_size=H, depth=2, gemm_path=L.GEMM_TC_2XF16)
eng.pack_weights(L.NET_ACTOR, synth.weights_to_device(synth.make_weights(77, 65, 40, H, 2), dev))
eng.pack_weights(L.NET_CRITIC, synth.weights_to_device(synth.make_weights(78, 475, 1, H, 2), dev))
f32 = dict(device=dev, dtype=torch.float32)
ld = (N + 3) // 4 * 4
d = synth.make_batch_device(1237, T, N, dev)
command = torch.zeros((T + 1, 16, ld), **f32)
eng.command_update(command[0], d["cmd_mode"][0], d["cmd_u6"][0], d["cmd_u_arms"][0], None, N)
io = {"state": d["state"], "noise": d["noise"], "episode": d["episode"], "eps_action": d["eps_action"],
      "u_switch": d["u_switch"], "cmd_mode": d["cmd_mode"], "cmd_u6": d["cmd_u6"], "cmd_u_arms": d["cmd_u_arms"],
      "command": command, "pg_carry": torch.zeros((3, ld), **f32),
      "actor_carry": torch.zeros((2, 2, N, H), **f32), "critic_carry": torch.zeros((2, 2, N, H), **f32),
      "lpf": torch.zeros((20, ld), **f32), "actor_obs": None, "action": torch.zeros((T, 20, ld), **f32),
      "log_prob": torch.zeros((T, ld), **f32), "ctrl": torch.zeros((T, 20, ld), **f32), "term_codes": None,
      "done": torch.zeros((T, ld), device=dev, dtype=torch.uint8),
      "success": torch.zeros((T, ld), device=dev, dtype=torch.uint8), "value": torch.zeros((T, ld), **f32), "T": T}
for _ in range(20):
    eng.rollout(io, N)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    eng.rollout(io, N)
e1.record()
torch.cuda.synchronize()
print(f"kbs_rollout T=1, {N} envs: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per control step")
eng.profile(True)
for _ in range(50):
    eng.rollout(io, N)
torch.cuda.synchronize()
p = eng.profile_read()
eng.profile(False)
print({k: (round(v[0] / 50 * 1e3, 1), v[1] // 50) for k, v in p.items() if k != "_overflow"}, "us per step, launches per step")
eng.close()
