"""Timing probes of input_proj_fused_kernel (KBS_FPROJ_DBG bits, see FProjArgs::dbg): CUDA-event time of the projection
launch inside one 4 096 x 100 kbs_rollout.  Results with dbg != 0 are WRONG by construction; timing only."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import kbot_joystick_b200  # noqa: F401
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep

dev = torch.device("cuda:0")
H, N, T = 256, 4096, 100
eng = KbotStep(hidden_size=H, depth=2, gemm_path=L.GEMM_TC_2XF16)
eng.pack_weights(L.NET_ACTOR, synth.weights_to_device(synth.make_weights(77, 65, 40, H, 2), dev))
eng.pack_weights(L.NET_CRITIC, synth.weights_to_device(synth.make_weights(78, 475, 1, H, 2), dev))
f32 = dict(device=dev, dtype=torch.float32)
ld = N
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
import ctypes as C
tr = torch.zeros((2 * 148 + 8) * 16, device=dev, dtype=torch.int64)
eng.lib.kbs_debug_tc_trace_attach(eng._h, C.c_void_p(tr.data_ptr()), 0, 0)
NAMES = ["total", "items", "rawprod:wait raw_empty", "wprod:wait empty", "conv:wait raw_full", "conv:wait empty",
         "issuer:wait full", "issuer:wait acc_empty", "epi:wait acc_full"]
for dbg in [int(x) for x in (sys.argv[1:] or ["0"])]:
    os.environ["KBS_FPROJ_DBG"] = str(dbg)
    for _ in range(2):
        eng.rollout(io, N)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(3):
        eng.rollout(io, N)
    torch.cuda.synchronize()
    p = eng.profile_read()
    eng.profile(False)
    ms, cnt = p["proj_tc_kernel"]
    print(f"dbg={dbg:3d}  proj {ms / cnt * 1e3:8.1f} us per launch   persist {p['rollout_persist_kernel'][0] / 3:6.3f} ms", flush=True)
    t = tr.view(-1, 16)[148:296].double().cpu()
    print("        per-CTA mean cycles: " + ", ".join(f"{n} {t[:, i].mean().item():.0f}" for i, n in enumerate(NAMES)), flush=True)
eng.close()
