"""tcgen05 datapath on its own: the 3xTF32 LSTM-layer GEMM (kbs_debug_tc_gates) against fp64 NumPy, with one-hot probes
that localise an operand-layout error (which k / which row / which gate column went where) if the product is wrong."""

import numpy as np
import pytest
import torch

import harness as Hn
from kbot_joystick_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _gates(e, net, layer, x, h, dev):
    n, H = x.shape
    out = torch.full((n, 4 * H), float("nan"), device=dev)
    xd, hd = torch.from_numpy(x).to(dev), torch.from_numpy(h).to(dev)
    L.check(e.lib.kbs_debug_tc_gates(e._h, net, layer, L.ptr(xd), L.ptr(hd), L.ptr(out), n,
                                     torch.cuda.current_stream().cuda_stream), "kbs_debug_tc_gates")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _ref(w, layer, x, h):
    lw = w["layers"][layer]
    f = np.float64
    return x.astype(f) @ lw["w_ih"].astype(f).T + h.astype(f) @ lw["w_hh"].astype(f).T + lw["b"].astype(f)


@pytest.mark.parametrize("path", [pytest.param(L.GEMM_TC_3XTF32, id="tc3xtf32"), pytest.param(L.GEMM_TC_2XF16, id="tc2xf16")])
@pytest.mark.parametrize("hidden", [256, 128])
def test_tc_layer_gemm_probes_and_accuracy(cuda_device, lib_built, hidden, path):
    dev = cuda_device
    e, wa, wc = Hn.make_engine(hidden=hidden, gemm_path=path, device=dev)
    H = hidden
    rng = np.random.default_rng(0)
    # 1. one-hot probes: x = e_k in row r  =>  gates[r] - b = W_ih[:, k]; anything else pinpoints a layout error
    n = 130
    x = np.zeros((n, H), np.float32)
    h = np.zeros((n, H), np.float32)
    ks = rng.integers(0, H, size=n)
    x[np.arange(n), ks] = 1.0
    g = _gates(e, L.NET_ACTOR, 0, x, h, dev)
    ref = _ref(wa, 0, x, h)
    err = np.abs(g - ref)
    if not (err.max() < 1e-6):
        w = wa["layers"][0]["w_ih"].astype(np.float64)
        b = wa["layers"][0]["b"].astype(np.float64)
        lines = [f"one-hot probe failed: max err {err.max():.3e}, nan count {np.isnan(g).sum()}"]
        for r in (0, 1, 8, 33, 127, 128, 129):
            d = g[r].astype(np.float64) - b
            # which column of W_ih does row r's output match best?
            cand = np.argmin(np.abs(w - d[:, None]).sum(0)) if np.isfinite(d).all() else -1
            lines.append(f"  row {r}: fed k={ks[r]}, output matches W_ih[:, {cand}]; |d|max={np.nanmax(np.abs(d)):.3e}")
        pytest.fail("\n".join(lines))
    # 2. same through the recurrent half
    g = _gates(e, L.NET_ACTOR, 1, h, x, dev)
    assert np.abs(g - _ref(wa, 1, h, x)).max() < 1e-6
    # 3. accuracy on dense random data, critic weights, ragged n: 3xTF32 must sit at fp32 rounding level
    for n in (1, 77, 128, 1000):
        x = rng.standard_normal((n, H)).astype(np.float32)
        h = np.tanh(rng.standard_normal((n, H))).astype(np.float32)
        g = _gates(e, L.NET_CRITIC, 1, x, h, dev)
        ref = _ref(wc, 1, x, h)
        ref32 = (x @ wc["layers"][1]["w_ih"].T + h @ wc["layers"][1]["w_hh"].T + wc["layers"][1]["b"]).astype(np.float64)
        e_tc, e_32 = np.abs(g - ref).max(), np.abs(ref32 - ref).max()
        m_tc = np.abs(g - ref).mean()
        print(f"path={path} H={H} n={n}: split-MMA max/mean abs err vs fp64 {e_tc:.3e}/{m_tc:.3e}; NumPy fp32 max {e_32:.3e}")
        assert np.isfinite(g).all()
        # tcgen05 adds every K=8 partial sum with truncation (tools/tc_accum_probe.py): K/8 = 64 truncations of the
        # O(1) hi.hi accumulator bound the error at ~64 * 0.6 ulp; the hi.lo/lo.hi terms sit in their own accumulator.
        assert e_tc < 8e-6 and m_tc < 1.5e-6, (e_tc, m_tc, e_32)
    e.close()
