"""BASELINE.json configs[2] and configs[4] on one GPU (the default bench.py line is configs[1]):
  configs[2]  full rollout + rewards + GAE, 16 384 envs x 256 steps (env-sharded across GPUs = the same per-GPU work)
  configs[4]  actor-only inference batch sweep 2^10 .. 2^20 envs: the deployed policy step of convert.py:84-119
Prints one JSON line per measurement (CUDA events, 3 warm-up + 5 timed iterations, inputs larger than L2 or rotated)."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import kbot_joystick_b200  # noqa: F401
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep

from kbot_joystick_b200.sharding import max_over_ranks
# one process per GPU under torch.distributed.run: every rank runs the same per-GPU work on its own envs (environments are
# independent: no data-path collective); times are the MAX over ranks, throughput = per-GPU units x world / that time
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
H = 256
eng = KbotStep(hidden_size=H, depth=2, gemm_path=L.GEMM_TC_2XF16)
eng.pack_weights(L.NET_ACTOR, synth.weights_to_device(synth.make_weights(77, 65, 40, H, 2), dev))
eng.pack_weights(L.NET_CRITIC, synth.weights_to_device(synth.make_weights(78, 475, 1, H, 2), dev))
f32 = dict(device=dev, dtype=torch.float32)


def timed(fn, warm=3, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return max_over_ranks(e0.elapsed_time(e1) / it, dev)


def emit(rec):
    rec["n_gpus"] = world
    for k in list(rec):
        if k.startswith("env_steps_per_s"):
            rec[k] *= world
    if world > 1:
        rec["n_envs_per_gpu"] = rec.pop("n_envs")
    if rank == 0:
        print(json.dumps(rec), flush=True)


def rollout_case(N, T, label="configs[2] full rollout + rewards + GAE", it=5, graph=False):
    ld = (N + 3) // 4 * 4
    d = synth.make_batch_device(1234 + 3, T, N, dev)
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
    rc = {"t_single": torch.zeros(ld, **f32), "airtime": torch.zeros((2, ld), **f32),
          "prev_contact": torch.ones((2, ld), device=dev, dtype=torch.uint8)}
    total, adv, tgt = (torch.zeros((T, ld), **f32) for _ in range(3))

    def step():
        eng.rollout(io, N)
        eng.rewards(io["state"], io["command"][:T], io["ctrl"], io["done"], rc, total=total, n_envs=N)
        eng.gae(io["value"], total, io["done"], io["success"], adv=adv, targets=tgt, n_envs=N)
        io["command"][0].copy_(io["command"][T])

    ms = timed(step, it=it)
    assert eng.device_status() == 0 and torch.isfinite(adv).all()
    rec = {"config": label, "n_envs": N, "T": T, "ms_per_rollout": ms, "env_steps_per_s": N * T / (ms * 1e-3), "n_gpus": 1}
    if graph:   # the same step replayed as one CUDA graph (the library launches only on the caller's stream)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            msg = timed(g.replay, it=it)
            rec["ms_per_rollout_cuda_graph"] = msg
            rec["env_steps_per_s_cuda_graph"] = N * T / (msg * 1e-3)
        except Exception as ex:  # noqa: BLE001
            rec["cuda_graph_error"] = repr(ex)[:200]
    emit(rec)


def policy_sweep():
    for p in range(10, 21):
        N = 1 << p
        g = torch.Generator(device=dev).manual_seed(p)
        ja = torch.randn((N, 20), generator=g, **f32) * 0.3
        jv = torch.randn((N, 20), generator=g, **f32)
        pg = torch.randn((N, 3), generator=g, **f32)
        gy = torch.randn((N, 3), generator=g, **f32)
        cmd = torch.randn((N, 16), generator=g, **f32) * 0.3
        carry = [torch.zeros((N, 2 * 2 * H + 20), **f32)]

        def step():
            _, carry[0] = eng.policy_step(ja, jv, pg, gy, cmd, carry[0])

        ms = timed(step)
        rec = {"config": "configs[4] actor-only policy step (convert.py step_fn)", "n_envs": N, "ms_per_step": ms,
               "env_steps_per_s": N / (ms * 1e-3), "n_gpus": 1}
        if p in (12, 20):   # per-kernel CUDA-event pass (events around every launch of one step)
            eng.profile(True)
            step()
            torch.cuda.synchronize()
            prof = eng.profile_read()
            eng.profile(False)
            rec["kernel_breakdown_ms"] = {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in prof.items() if k != "_overflow"}
        emit(rec)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "rollout"):
        rollout_case(16384, 256)
    if which in ("all", "online"):
        # the call ksim's engine loop makes with MJX between control steps: one kbs_rollout per control step (T = 1)
        rollout_case(4096, 1, "configs[1] online form: one kbs_rollout + rewards + GAE call per control step (T = 1)", it=200, graph=True)
    if which in ("all", "policy"):
        policy_sweep()
    eng.close()
