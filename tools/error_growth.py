"""Error growth of the tcgen05 FP16-split (and 3xTF32) datapath through the recurrence: the fused
rollout at the reference's rollout length and beyond, against the NumPy oracle (fp32) on the same seeded inputs, as PURE
relative errors |gpu - oracle| / |oracle| per output and time bucket (no absolute floor; elements with |oracle| < 1e-3 of the
output's scale are reported separately as `near_zero`, where a relative error is not meaningful).
Writes one JSON (stdout) -> profiles/r02_error_growth.md is generated from it by tools/error_growth.py --render <json>."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "oracle"))
BUCKETS = ((0, 10), (10, 25), (25, 50), (50, 100), (100, 180), (180, 256))


def rel_stats(gpu, ref, t_axis_len):
    import numpy as np

    gpu, ref = np.asarray(gpu, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() + 1e-300
    out = {}
    for lo, hi in BUCKETS:
        if lo >= t_axis_len:
            break
        g, r = gpu[lo:min(hi, t_axis_len)], ref[lo:min(hi, t_axis_len)]
        big = np.abs(r) >= 1e-3 * scale
        rel = np.abs(g - r)[big] / np.abs(r)[big]
        small = np.abs(g - r)[~big]
        out[f"{lo}-{min(hi, t_axis_len)}"] = {
            "max_rel": float(rel.max()) if rel.size else 0.0, "p999_rel": float(np.quantile(rel, 0.999)) if rel.size else 0.0,
            "median_rel": float(np.median(rel)) if rel.size else 0.0, "n": int(rel.size),
            "near_zero_max_abs_over_scale": float(small.max() / scale) if small.size else 0.0}
    return out


def run_case(T, N, path_name):
    import numpy as np
    import torch
    import harness as Hn
    from kbot_joystick_b200 import _lib as L, synth

    dev = torch.device("cuda:0")
    path = {"tc2xf16": L.GEMM_TC_2XF16, "tc3xtf32": L.GEMM_TC_3XTF32, "simt": L.GEMM_SIMT_FP32}[path_name]
    b = Hn.Batch(4000 + N + T, T, N, dev)
    eng, wa, wc = Hn.make_engine(hidden=256, depth=2, gemm_path=path, device=dev)
    io = Hn.rollout_buffers(b, 256, 2, True)
    eng.rollout(io, N)
    torch.cuda.synchronize()
    assert eng.device_status() == 0
    ref = Hn.oracle_rollout(b, wa, wc, 256, 2, True)
    s = synth.from_soa
    res = {"T": T, "N": N, "path": path_name, "done_fraction": float(ref["done"].mean()), "outputs": {}}
    for name, g, r in (("log_prob", s(io["log_prob"], N), ref["log_prob"]), ("value", s(io["value"], N), ref["value"]),
                       ("action", s(io["action"], N, (20,)), ref["action"]), ("ctrl", s(io["ctrl"], N, (20,)), ref["ctrl"]),
                       ("actor_obs", s(io["actor_obs"], N, (65,)), ref["actor_obs"])):
        res["outputs"][name] = rel_stats(g, r, T)
    # the pass/fail criterion of the parity tests (1e-5 relative + the stated absolute floor), for reference
    res["scaled_errors_of_the_parity_test"] = {k: float(v) for k, v in Hn.compare_rollout(io, ref, N, True).items()}
    # get_ppo_variables on the stored trajectory (T = 100 only: train.py:1766, 1776)
    if T <= 100 and path_name != "simt":
        import kbot_oracle as O
        p = O.OracleParams()
        ac, cc = torch.zeros((2, 2, N, 256), device=dev), torch.zeros((2, 2, N, 256), device=dev)
        lpf = torch.zeros((20, b.ld), device=dev)
        eng2, _, _ = Hn.make_engine(hidden=256, depth=2, gemm_path=path, device=dev)
        cobs = torch.zeros((T, 475, b.ld), device=dev)
        for t in range(T):          # the critic observations the rollout consumed (kbs_observations per step)
            eng2.observations(b.state_at(t), io["command"][t], noise=None, episode=b.episode, critic_obs=cobs[t], n_envs=N)
        out = eng2.ppo_variables(io["actor_obs"], io["action"], io["done"], ac, lpf, cobs, cc, n_envs=N)
        torch.cuda.synchronize()
        res["outputs"]["ppo_variables.log_probs vs rollout log_prob (oracle)"] = rel_stats(s(out["log_probs"], N), ref["log_prob"], T)
        res["outputs"]["ppo_variables.values vs rollout value (oracle)"] = rel_stats(s(out["values"], N), ref["value"], T)
        eng2.close()
    eng.close()
    return res


def render(js):
    d = json.loads(Path(js).read_text())
    lines = ["# r02 — error growth of the tensor-core datapaths through the recurrence (VERDICT r01 item 2)", "",
             "Fused rollout (`kbs_rollout`) against the NumPy oracle (fp32, `oracle/kbot_oracle.py`) on the same seeded inputs; PURE relative",
             "error `|gpu - oracle| / |oracle|` per output and time bucket (elements with `|oracle| < 1e-3 x max|oracle|` are excluded from the",
             "relative statistics and reported as `near-zero: max |err| / scale`).  Produced by `tools/error_growth.py` on one B200.",
             "north_star asks 1e-5 relative for log-probs / values / torques / actions; the parity tests use `1e-5 |ref| + atol` (atol per",
             "quantity, tests/harness.py) -- the last column block restates their scaled errors (<= 1 passes).", ""]
    for c in d["cases"]:
        lines += [f"## T = {c['T']}, N = {c['N']}, datapath {c['path']} (done fraction {c['done_fraction']:.3f})", "",
                  "| output | steps | max rel | 99.9 pct rel | median rel | near-zero: max abs / scale |", "|---|---|---|---|---|---|"]
        for name, st in c["outputs"].items():
            for bk, v in st.items():
                lines.append(f"| {name} | {bk} | {v['max_rel']:.2e} | {v['p999_rel']:.2e} | {v['median_rel']:.2e} | {v['near_zero_max_abs_over_scale']:.2e} |")
        lines += ["", "parity-test scaled errors (<= 1 passes): " + ", ".join(f"{k} {v:.3f}" for k, v in c["scaled_errors_of_the_parity_test"].items()), ""]
    lines += ["Reading: see DESIGN.md section 2 (tolerances)."]
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--render":
        sys.stdout.write(render(sys.argv[2]))
        sys.exit(0)
    cases = []
    for T, N, path in ((100, 4096, "tc2xf16"), (256, 512, "tc2xf16"), (256, 512, "tc3xtf32"), (100, 1024, "simt")):
        cases.append(run_case(T, N, path))
        print(f"[error_growth] T={T} N={N} {path}: done", file=sys.stderr, flush=True)
    print(json.dumps({"cases": cases}))
