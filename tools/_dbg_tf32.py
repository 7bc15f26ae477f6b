import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import harness as Hn
from harness import Batch
from kbot_joystick_b200 import _lib as L, synth
dev = torch.device("cuda:0")
np.set_printoptions(linewidth=220, precision=2, suppress=False)
def run(b, N, staged, path):
    os.environ["KBS_PROJ_STAGED"] = staged
    e, _, _ = Hn.make_engine(gemm_path=path, device=dev)
    io = Hn.rollout_buffers(b, 256, 2)
    e.rollout(io, N)
    torch.cuda.synchronize()
    v = io["value"].clone().cpu().numpy()
    e.close()
    return v
path = L.GEMM_TC_3XTF32
T, N = 1, 260
b = Batch(778, T, N, dev)
s0 = run(b, N, "1", path)
for i in range(4):
    f = run(b, N, "0", path)
    d = np.abs(f - s0)[0, :N]
    print("run", i, "max", d.max(), "n bad", (d > 1e-6).sum(), "bad rows", np.nonzero(d > 1e-6)[0][:40], flush=True)
    print("   errs panel0[:16]", d[:16], "panel2", d[256:260])
