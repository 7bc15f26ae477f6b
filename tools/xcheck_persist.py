import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import torch, numpy as np
import harness as Hn
from kbot_joystick_b200 import _lib as L
dev = torch.device("cuda:0")
T, N = int(sys.argv[1]), int(sys.argv[2])
b = Hn.Batch(4242, T, N, dev)
outs = []
for per_step in ("0", "1", "0"):
    os.environ["KBS_TC_PER_STEP"] = per_step
    e, _, _ = Hn.make_engine(gemm_path=L.GEMM_TC_2XF16, device=dev)
    io = Hn.rollout_buffers(b, 256, 2)
    e.rollout(io, N); torch.cuda.synchronize()
    print("status", e.device_status())
    outs.append(io); e.close()
for name, (x, y) in {"persist vs per-step": (outs[0], outs[1]), "persist vs persist": (outs[0], outs[2])}.items():
    for k in ("actor_carry", "critic_carry", "action", "log_prob", "value", "lpf"):
        d = (x[k] - y[k]).abs()
        print(name, k, "max diff %.3e" % d.max().item(), "n diff", int((d > 0).sum().item()), "of", d.numel())
d = (outs[0]["actor_carry"] - outs[1]["actor_carry"]).abs()
idx = (d > 0).nonzero()
print(idx[:10].tolist())
