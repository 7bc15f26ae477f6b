"""Per-CTA phase timing of the tcgen05 LSTM layer kernel (clock64 stamps) inside a real fused rollout."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import kbot_joystick_b200
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
dev = torch.device("cuda:0")
N, T, H = 4096, 20, 256
d = synth.make_batch_device(7, T, N, dev)
f32 = dict(device=dev, dtype=torch.float32)
for name, path in (("tf32", L.GEMM_TC_3XTF32), ("f16", L.GEMM_TC_2XF16)):
    e = KbotStep(hidden_size=H, gemm_path=path)
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
    ctas = 148
    for layer in (0, 1):
        tr = torch.zeros((ctas * 8 + 256 + 2 * ctas,), dtype=torch.int64, device=dev)
        e.lib.kbs_debug_tc_trace_attach(e._h, tr.data_ptr(), 10, layer)
        for rep in range(3):
            e.rollout(io, N)
        torch.cuda.synchronize()
        full_t = tr.cpu().numpy().astype(np.float64); t = full_t[:ctas * 8].reshape(ctas, 8); t2 = full_t[ctas * 8:ctas * 8 + 256].reshape(4, 64); t3 = full_t[ctas * 8 + 256:].reshape(ctas, 2)
        dd = {"setup": t[:, 1] - t[:, 0], "first stage wait": t[:, 2] - t[:, 1], "K loop (issue)": t[:, 3] - t[:, 2],
              "MMA drain->epi start": t[:, 4] - t[:, 3], "epilogue (warp 5)": t[:, 7] - t[:, 4], "total": t[:, 5] - t[:, 0]}
        print(f"--- {name} layer {layer} in rollout: n={N} ctas={ctas} (cycles, median / p90 / max)")
        for k, v in dd.items():
            print(f"   {k:22s} {np.median(v):9.0f} {np.percentile(v, 90):9.0f} {v.max():9.0f}")
        g0 = t3[:, 0].min()
        print(f"   globaltimer (ns from first CTA start): CTA start median {np.median(t3[:,0]-g0):.0f} max {np.max(t3[:,0]-g0):.0f}; "
              f"CTA end median {np.median(t3[:,1]-g0):.0f} max {np.max(t3[:,1]-g0):.0f}")
        if False:
            base = t2[2, 0]
            print("   CTA 0 per-stage stamps relative to first full (prod_empty_ok, prod_issued, mma_full_ok, mma_issued):")
            for g in range(0, 40):
                print(f"     g={g:2d}  {t2[0,g]-base:8.0f} {t2[1,g]-base:8.0f} {t2[2,g]-base:8.0f} {t2[3,g]-base:8.0f}")
        sm = t[:, 6].astype(int)
        span = (t[:, 5].max() - t[:, 0].min())
        print(f"   persistent CTAs: {len(np.unique(sm))} SMs; stamps 2-4,7 are for each CTA's FIRST item; total = whole CTA lifetime")
    e.lib.kbs_debug_tc_trace_attach(e._h, None, -1, 0)
    e.close()
