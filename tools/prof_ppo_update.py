import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import torch
import kbot_joystick_b200
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
from kbot_joystick_b200.ppo import PpoUpdater
dev = torch.device("cuda:0")
import os
H, N, T = 256, int(os.environ.get("PROF_ENVS", "512")), 100
ld = N
eng = KbotStep(hidden_size=H, depth=2, gemm_path=L.GEMM_TC_2XF16)
up = PpoUpdater(eng, synth.make_weights(77, 65, 40, H, 2), synth.make_weights(78, 475, 1, H, 2))
g = torch.Generator(device=dev).manual_seed(1)
f32 = dict(device=dev, dtype=torch.float32)
rn = lambda *s, sc=1.0: torch.randn(s, generator=g, **f32) * sc
batch = {"actor_obs": rn(T, 65, ld, sc=0.7), "critic_obs": rn(T, 475, ld, sc=0.7), "action": rn(T, 20, ld, sc=0.3),
         "done": (torch.rand((T, ld), generator=g, device=dev) < 0.01).to(torch.uint8),
         "old_log_probs": rn(T, ld) - 20.0, "advantages": rn(T, ld), "value_targets": rn(T, ld, sc=0.5), "old_values": rn(T, ld, sc=0.5)}
fwd = eng.ppo_variables(batch["actor_obs"], batch["action"], batch["done"], torch.zeros((2, 2, N, H), **f32), torch.zeros((20, ld), **f32),
                        batch["critic_obs"], torch.zeros((2, 2, N, H), **f32), want_std=False, n_envs=N)
batch["old_log_probs"], batch["old_values"] = fwd["log_probs"] + rn(T, ld, sc=0.1), fwd["values"] + rn(T, ld, sc=0.2)
for _ in range(2):
    up.grads(batch, N)
torch.cuda.synchronize()
eng.profile(True)
up.grads(batch, N)
torch.cuda.synchronize()
p = eng.profile_read(); eng.profile(False)
print({k: (round(v[0], 3), v[1]) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0] if isinstance(kv[1], tuple) else 0) if isinstance(v, tuple)}, p.get("_overflow"))
