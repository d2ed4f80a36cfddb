"""torchrun --nproc-per-node N tools/multi_gpu_check.py [--model toy128] [--envs 16] [--steps 6] [--discrete]
                                                        [--oracle-rows 8] [--every 4]
Env-sharded rollout (env i -> rank i % N, custom_eval_callback.py:385): every rank steps its shard from a CUDA graph,
the action tokens travel through the library's token ring (written by the argmax kernel inside the graph) and ONE
all-gather per `every` steps. Rank 0 checks the gathered tokens, put back in env order, against
  (a) its own single-GPU run of ALL envs (bit-exact: sharding must not change a single token), and
  (b) the fp32 CPU oracle on a subsample of env rows (bit-exact action tokens / argmax actions).
Exit code 0 only if both hold. BASELINE.json configs[4]: --model 110M --envs 256 --discrete at N = 1, 2, 4, 8."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lram_b200 import _lib as L
from lram_b200.config import preset
from lram_b200.engine import XLSTMEngine
from lram_b200.rollout import OverlappedTokenGather, shard_env_ids
from lram_b200.synth import make_state_dict, make_stream

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="toy128")
ap.add_argument("--envs", type=int, default=16)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--every", type=int, default=4)
ap.add_argument("--oracle-rows", type=int, default=8)
ap.add_argument("--discrete", action="store_true")
args = ap.parse_args()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
assert args.envs % world == 0
cfg = preset(args.model)
sd = make_state_dict(cfg, seed=0)
flags = (L.XL_FLAG_DISCRETE if args.discrete else 0) | L.XL_FLAG_GRAPH


def inputs(env_ids):
    states, rtg, _ = make_stream(cfg, env_ids, args.steps, domains="mixed")
    if args.discrete:     # Atari-style envs: frames -> embeddings; synthetic post-ReLU state embeddings per env
        emb = torch.stack([torch.randn(args.steps, cfg.d, generator=torch.Generator().manual_seed(1000 + e)).clamp_min(0)
                           for e in env_ids], dim=1)            # [steps, B, d]
        return emb, rtg
    return torch.from_numpy(states), rtg


def run(env_ids, gather):
    B = len(env_ids)
    eng = XLSTMEngine(cfg, sd, max_batch=B)
    cache = eng.new_state(B)
    x, rtg = inputs(env_ids)
    g = OverlappedTokenGather(B, cfg.act_dim, world, dev, every=args.every, keep=True, engine=eng) if gather else None
    s_dev = torch.empty(B, x.shape[-1], device=dev)
    r_dev = torch.empty(B, device=dev)
    out, toks = None, []
    for t in range(args.steps):
        s_dev.copy_(x[t])
        r_dev.copy_(torch.from_numpy(rtg[t]))
        out = eng.policy_step(cache, s_dev, r_dev, flags=flags, out=out, state_embeds=args.discrete)
        if g is not None:
            g.submit(out["action_tokens"])
        else:
            toks.append(out["action_tokens"].clone())
    if g is not None:
        g.finish()
        torch.cuda.synchronize()
        res = torch.cat([g.global_view(r) for r in g.results], dim=0)      # [steps, n_envs, A]
    else:
        torch.cuda.synchronize()
        res = torch.stack(toks)
    eng.close()
    return res


if world > 1:
    full = run(shard_env_ids(args.envs, rank, world), gather=True)
else:
    full = run(list(range(args.envs)), gather=False)
ok_single, ok_oracle = True, True
if rank == 0:
    A = 1 if args.discrete else cfg.act_dim
    if world > 1:
        ref = run(list(range(args.envs)), gather=False)
        ok_single = bool(torch.equal(full[..., :A], ref[..., :A]))
    if args.oracle_rows > 0:
        from oracle.xlstm_oracle import OraclePolicy            # checker only
        n = min(args.oracle_rows, args.envs)
        rows = sorted({int(round(i * (args.envs - 1) / max(n - 1, 1))) for i in range(n)})
        x, rtg = inputs(rows)
        ora, pkv = OraclePolicy(cfg, sd), None
        for t in range(args.steps):
            o = ora.step(x[t], torch.from_numpy(rtg[t]), past_key_values=pkv, discrete=args.discrete,
                         state_embeds=args.discrete)
            pkv = o["past_key_values"]
            got = full[t].cpu()[rows][:, :A].long()
            if not torch.equal(got, o["action_tokens"].reshape(len(rows), -1)):
                ok_oracle = False
    print(f"multi-gpu check: world={world} model={args.model} envs={args.envs} ({args.envs // world}/GPU) steps={args.steps} "
          f"discrete={args.discrete} gather_every={args.every} (token ring) -> vs single GPU: "
          f"{'BIT-EXACT' if ok_single else 'MISMATCH'}; vs fp32 oracle on {args.oracle_rows} env rows: "
          f"{'BIT-EXACT' if ok_oracle else 'MISMATCH'}", flush=True)
ok = ok_single and ok_oracle
if world > 1:
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    ok = flag.item() == 1
sys.exit(0 if ok else 1)
