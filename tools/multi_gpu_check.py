"""torchrun --nproc-per-node N tools/multi_gpu_check.py [--model toy128] [--envs 16] [--steps 5] [--discrete]
Env-sharded rollout (env i -> rank i % N, custom_eval_callback.py:385) must give bit-identical action tokens to
the single-GPU run of the same envs. Rank 0 also replays the whole batch alone and compares."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.rollout import gather_env_results, shard_env_ids
from lram_b200.synth import make_state_dict, make_stream

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="toy128")
ap.add_argument("--envs", type=int, default=16)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--discrete", action="store_true")
args = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = preset(args.model)
sd = make_state_dict(cfg, seed=0)
flags = L.XL_FLAG_DISCRETE if args.discrete else 0


def run(env_ids):
    eng = XLSTMEngine(cfg, sd, max_batch=len(env_ids))
    cache = eng.new_state(len(env_ids))
    states, rtg, _ = make_stream(cfg, env_ids, args.steps, domains="mixed")
    toks = []
    for t in range(args.steps):
        o = eng.policy_step(cache, torch.from_numpy(states[t]).cuda(), torch.from_numpy(rtg[t]).cuda(), flags=flags | L.XL_FLAG_GRAPH)
        toks.append(o["action_tokens"].clone())
    torch.cuda.synchronize()
    eng.close()
    return torch.stack(toks, dim=1)          # [B_local, steps, A]


local = run(shard_env_ids(args.envs, rank, world))
full = gather_env_results(local, args.envs, rank, world)
ok = True
if rank == 0:
    ref = run(list(range(args.envs)))
    ok = bool(torch.equal(full, ref))
    print(f"multi-gpu check: world={world} model={args.model} envs={args.envs} steps={args.steps} discrete={args.discrete} "
          f"-> {'BIT-EXACT' if ok else 'MISMATCH'}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
