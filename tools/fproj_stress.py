"""Run-to-run reproducibility of the fused rollout with input_proj_fused_kernel: fresh engine per repetition, outputs
must be bitwise equal to the first repetition (one MMA-issuing thread per CTA fixes every accumulation order)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import harness as Hn
from harness import Batch
from kbot_joystick_b200 import _lib as L
dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "f16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
path = L.GEMM_TC_2XF16 if kind == "f16" else L.GEMM_TC_3XTF32
for (T, N) in [(1, 260), (6, 132), (20, 4096)]:
    b = Batch(778, T, N, dev)
    ref = None
    bad = []
    for i in range(reps):
        e, _, _ = Hn.make_engine(gemm_path=path, device=dev)
        io = Hn.rollout_buffers(b, 256, 2)
        e.rollout(io, N)
        torch.cuda.synchronize()
        v = io["value"].clone()
        e.close()
        if ref is None:
            ref = v
        elif not torch.equal(v, ref):
            d = (v - ref).abs()[:, :N]
            rows = torch.nonzero(d.max(dim=0).values > 0).flatten().tolist()
            bad.append((i, float(d.max()), rows[:4], rows[-1], len(rows), int(torch.nonzero(d.max(dim=1).values > 0).flatten()[0])))
    print(kind, "T,N", T, N, "reps", reps, "differing reps:", bad, flush=True)
