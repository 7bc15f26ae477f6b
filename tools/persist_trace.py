"""Per-CTA accounting of rollout_persist_kernel: total cycles, producer dependency-wait cycles, items, issuer wait-for-data cycles."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import kbot_joystick_b200
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
dev = torch.device("cuda:0")
N, T, H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 20, 256
d = synth.make_batch_device(7, T, N, dev)
f32 = dict(device=dev, dtype=torch.float32)
e = KbotStep(hidden_size=H, gemm_path=L.GEMM_TC_2XF16)
e.pack_weights(L.NET_ACTOR, synth.weights_to_device(synth.make_weights(1, 65, 40, H, 2), dev))
e.pack_weights(L.NET_CRITIC, synth.weights_to_device(synth.make_weights(2, 475, 1, H, 2), dev))
io = {"state": d["state"], "noise": d["noise"], "episode": d["episode"], "eps_action": d["eps_action"],
      "u_switch": d["u_switch"], "cmd_mode": d["cmd_mode"], "cmd_u6": d["cmd_u6"], "cmd_u_arms": d["cmd_u_arms"],
      "command": torch.zeros((T + 1, 16, N), **f32), "pg_carry": torch.zeros((3, N), **f32),
      "actor_carry": torch.zeros((2, 2, N, H), **f32), "critic_carry": torch.zeros((2, 2, N, H), **f32),
      "lpf": torch.zeros((20, N), **f32), "actor_obs": None, "action": torch.zeros((T, 20, N), **f32),
      "log_prob": torch.zeros((T, N), **f32), "ctrl": torch.zeros((T, 20, N), **f32), "term_codes": None,
      "done": torch.zeros((T, N), device=dev, dtype=torch.uint8), "success": torch.zeros((T, N), device=dev, dtype=torch.uint8),
      "value": torch.zeros((T, N), **f32), "T": T}
import os
ctas = int(os.environ.get("KBS_PERSIST_GRID", "148"))
tr = torch.zeros((ctas * 16 + 1024,), dtype=torch.int64, device=dev)
e.lib.kbs_debug_tc_trace_attach(e._h, tr.data_ptr(), 0, 0)
for rep in range(3):
    e.rollout(io, N)
torch.cuda.synchronize()
print("status", e.device_status())
t = tr.cpu().numpy()[:ctas * 16].reshape(ctas, 16).astype(np.float64)
names = ("total cycles", "poller wait cycles", "items", "issuer wait-for-stage", "epilogue cycles (LSTM items)",
         "epilogue cycles (head items)", "epilogue wait-for-accumulator", "issuer wait-for-TMEM", "LSTM epi: TMEM pull", "LSTM epi: pull+math+stores",
         "LSTM epi: publish (fence+red)", "LSTM epi: syncwarp+threadfence only", "producer: wait for a free stage", "producer: expect_tx + bulk issue", "producer: wait deps + proxy fence", "-")
for i, name in enumerate(names):
    v = t[:, i]
    print(f"{name:32s} median {np.median(v):12.0f}  min {v.min():12.0f}  max {v.max():12.0f}   per item {np.median(v / t[:, 2]):9.0f}")
print("SM clock during the kernel: %.3f GHz (clock64 / globaltimer)" % np.median(t[:, 0] / t[:, 15]))
slots = T + 2
print("cycles per slot: %.0f (items per CTA per slot %.2f)" % (np.median(t[:, 0]) / slots, np.median(t[:, 2]) / slots))
e.lib.kbs_debug_tc_trace_attach(e._h, None, -1, 0)
e.close()
