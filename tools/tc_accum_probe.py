"""Diagnostic (GPU): how does tcgen05 kind::tf32 accumulate?  Inputs are made TF32-representable (lo planes = 0) so
every product is exact in fp32 and the only error source is the accumulation inside the tensor core."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np, torch
import kbot_joystick_b200
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep

def tf32(x):
    b = x.astype(np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return b.view(np.float32)

dev = torch.device("cuda:0")
H = 256
rng = np.random.default_rng(1)
w = synth.make_weights(3, 65, 40, H, 2)
for lw in w["layers"]:
    lw["w_ih"] = tf32(lw["w_ih"]); lw["w_hh"] = tf32(lw["w_hh"]); lw["b"] = np.zeros_like(lw["b"])
e = KbotStep(hidden_size=H, gemm_path=L.GEMM_TC_3XTF32)
e.pack_weights(L.NET_ACTOR, synth.weights_to_device(w, dev))
n = 256
def run(x, h):
    out = torch.empty((n, 4 * H), device=dev)
    xd, hd = torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev)     # keep alive across the call
    L.check(e.lib.kbs_debug_tc_gates(e._h, 0, 0, L.ptr(xd), L.ptr(hd),
                                     L.ptr(out), n, torch.cuda.current_stream().cuda_stream), "dbg")
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)
lw = w["layers"][0]
for keff in (8, 16, 64, 128, 256, 512):
    x = np.zeros((n, H), np.float32); h = np.zeros((n, H), np.float32)
    kx = min(keff, H); kh = max(keff - H, 0)
    x[:, :kx] = tf32(rng.standard_normal((n, kx)).astype(np.float32))
    if kh: h[:, :kh] = tf32(rng.standard_normal((n, kh)).astype(np.float32))
    ref = x.astype(np.float64) @ lw["w_ih"].astype(np.float64).T + h.astype(np.float64) @ lw["w_hh"].astype(np.float64).T
    g = run(x, h)
    err = g - ref
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    toward_zero = np.mean(np.sign(err) == -np.sign(ref))
    print(f"K={keff:4d}: max|err|={np.abs(err).max():.3e} mean|err|={np.abs(err).mean():.3e} "
          f"mean err/ulp={np.mean(np.abs(err)/ulp):.2f} max={np.max(np.abs(err)/ulp):.1f}  frac toward zero={toward_zero:.3f} "
          f"rms(ref)={ref.std():.3f}")
# positive-only data: partial sums grow monotonically -> pure truncation bias visible
x = tf32(np.abs(rng.standard_normal((n, H))).astype(np.float32)); h = np.zeros((n, H), np.float32)
wpos = {k: (np.abs(v) if k != "layers" else [{kk: np.abs(vv) for kk, vv in l.items()} for l in v]) for k, v in w.items()}
e.pack_weights(L.NET_ACTOR, synth.weights_to_device(wpos, dev))
ref = x.astype(np.float64) @ wpos["layers"][0]["w_ih"].astype(np.float64).T
g = run(x, h); err = g - ref
ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
print(f"positive data K=256: mean err/ulp = {np.mean(err/ulp):.2f} (negative = truncation), max|err/ulp| = {np.max(np.abs(err/ulp)):.1f}")
ref32 = (x @ wpos["layers"][0]["w_ih"].T).astype(np.float64)
print(f"  numpy fp32 same data: mean err/ulp = {np.mean((ref32-ref)/ulp):.2f}, max = {np.max(np.abs(ref32-ref)/ulp):.1f}")
