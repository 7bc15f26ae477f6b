"""BASELINE.json configs[3]: PPO minibatch update with the NCCL gradient all-reduce.  One process per GPU
(`python -m torch.distributed.run --nproc-per-node N tools/bench_ppo_update.py`, or plain `python` for N = 1): every rank
owns `--envs` stored trajectories of `--T` steps (train.py:1764-1766: batch_size 512, rollout 100 steps), computes the
gradients of the PPO loss (kbs_ppo_grad), all-reduces the 2.25 M-float gradient over NVLink / NVSwitch (NCCL) and applies
AdamW with the global-norm clip (kbs_grad_norm + kbs_adamw_step).  Prints one JSON line from rank 0 (CUDA events, max over ranks)."""
import argparse, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
import kbot_joystick_b200  # noqa: F401
from kbot_joystick_b200 import _lib as L, synth
from kbot_joystick_b200.engine import KbotStep
from kbot_joystick_b200.ppo import PpoUpdater

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=512)
ap.add_argument("--T", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--graph", type=int, default=1, help="replay kbs_ppo_grad as one CUDA graph (default) or launch eagerly")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
H, N, T = 256, a.envs, a.T
ld = (N + 3) // 4 * 4
eng = KbotStep(hidden_size=H, depth=2, gemm_path=L.GEMM_TC_2XF16)
wa, wc = synth.make_weights(77, 65, 40, H, 2), synth.make_weights(78, 475, 1, H, 2)      # same weights on every rank
up = PpoUpdater(eng, wa, wc)
g = torch.Generator(device=dev).manual_seed(100 + rank)
f32 = dict(device=dev, dtype=torch.float32)
rn = lambda *s, sc=1.0: torch.randn(s, generator=g, **f32) * sc
batch = {"actor_obs": rn(T, 65, ld, sc=0.7), "critic_obs": rn(T, 475, ld, sc=0.7), "action": rn(T, 20, ld, sc=0.3),
         "done": (torch.rand((T, ld), generator=g, device=dev) < 0.01).to(torch.uint8),
         "old_log_probs": rn(T, ld) - 20.0, "advantages": rn(T, ld), "value_targets": rn(T, ld, sc=0.5), "old_values": rn(T, ld, sc=0.5)}
# old log-probs / values near the current policy, as in a real update: taken from a forward-only pass (kbs_ppo_variables)
fwd = eng.ppo_variables(batch["actor_obs"], batch["action"], batch["done"], torch.zeros((2, 2, N, H), **f32), torch.zeros((20, ld), **f32),
                        batch["critic_obs"], torch.zeros((2, 2, N, H), **f32), want_std=False, n_envs=N)
batch["old_log_probs"], batch["old_values"] = fwd["log_probs"] + rn(T, ld, sc=0.1), fwd["values"] + rn(T, ld, sc=0.2)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(a.warmup):
    up.update(batch, N)
if a.graph:
    up.capture(batch, N)                         # the update is launch-bound on the host otherwise
    up.update(batch, N)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    st = up.update(batch, N)
e1.record()
sync()
ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    chk = up.param.double().sum().reshape(1).clone()          # replicas must stay identical after the all-reduced update
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert float(hi - lo) == 0.0, "replicas diverged"
assert eng.device_status() == 0, "device health word set during the update"
if rank == 0:
    msf = float(ms.item())
    print(json.dumps({"config": "configs[3] PPO minibatch update (kbs_ppo_grad + NCCL all-reduce + kbs_grad_norm + kbs_adamw_step)", "n_gpus": world,
                      "trajectories_per_gpu": N, "T": T, "ms_per_update": msf, "env_steps_per_s": world * N * T / (msf * 1e-3),
                      "grad_floats": int(up.grad.numel()), "loss": float(st["stats"][0]), "datapath": "persistent tcgen05 forward (SAVE) + BPTT kernels, split-K tcgen05 weight-gradient GEMMs (2xFP16-split)", "launch": "cuda-graph replay" if a.graph else "eager"}),
          flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
