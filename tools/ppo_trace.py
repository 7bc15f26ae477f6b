"""Per-CTA cycle accounting of the two persistent kernels of the PPO update (rollout_persist_kernel<SAVE>, bptt_persist_kernel)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import kbot_joystick_b200
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
from kbot_joystick_b200.ppo import PpoUpdater
dev = torch.device("cuda:0")
N, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 512, int(sys.argv[2]) if len(sys.argv) > 2 else 100, 256
ld = (N + 3) // 4 * 4
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
tr = torch.zeros((3 * 148 * 16 + 1024,), dtype=torch.int64, device=dev)
for _ in range(2):
    up.grads(batch, N)
eng.lib.kbs_debug_tc_trace_attach(eng._h, tr.data_ptr(), 0, 0)
up.grads(batch, N)
torch.cuda.synchronize()
print("status", eng.device_status())
t = tr.cpu().numpy()
fw = t[:148 * 16].reshape(148, 16).astype(np.float64)
bw = t[2 * 148 * 16:3 * 148 * 16].reshape(148, 16).astype(np.float64)
def show(title, a, names, slots):
    a = a[a[:, 2] > 0]
    print(f"== {title}: {len(a)} CTAs with work, SM clock {np.median(a[:, 0] / a[:, 15]):.3f} GHz, total {np.median(a[:, 15]) / 1e3:.1f} us, "
          f"{np.median(a[:, 0]) / slots:.0f} cycles per slot, items per CTA {np.median(a[:, 2]):.0f}")
    for i, nm in names:
        v = a[:, i]
        print(f"  {nm:38s} median {np.median(v):12.0f}   per item {np.median(v / a[:, 2]):9.0f}   share of total {np.median(v / a[:, 0]):.2f}")
show("forward (SAVE)", fw, [(1, "poller wait"), (3, "issuer wait-for-stage"), (7, "issuer wait-for-TMEM"), (6, "epilogue wait-for-accumulator"),
                            (4, "epilogue cycles (LSTM items)"), (5, "epilogue cycles (head items)"), (14, "producer: wait deps + proxy fence"), (8, "LSTM epi: TMEM pull"),
                            (9, "LSTM epi: acc ready -> end of item"), (12, "producer: wait for a free stage")], T + 2)
show("backward (BPTT)", bw, [(1, "poller wait"), (3, "issuer wait-for-stage"), (4, "issuer wait-for-TMEM"), (5, "epilogue wait-for-deps"),
                             (6, "epilogue wait-for-accumulator"), (7, "epilogue H-tile cycles"), (8, "epilogue X-tile cycles"), (9, "H items")], T + 3)
eng.lib.kbs_debug_tc_trace_attach(eng._h, None, -1, 0)
eng.close()
