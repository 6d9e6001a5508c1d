"""Multi-learner exchange check (run under torch.distributed.run, one rank per GPU):
the fused P2P reduce + Adam kernel vs NCCL all-reduce(mean) + the single-learner Adam kernel, on real LoRA
gradients of a small model (each rank scores its own shard).  Every rank must end with identical parameters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from distrl_llm_b200 import ops
from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
from distrl_llm_b200.p2p import P2PGroup, owned_slice
from distrl_llm_b200.policy import LMConfig, Policy
from distrl_llm_b200.trainer_prep import synthetic_candidates

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = LMConfig(vocab=4096, hidden=512, inter=1024, n_layers=3, n_q_heads=8, n_kv_heads=2, head_dim=64, lora_r=16, lora_alpha=16)
P, T, B, N = 24, 40, 4, 8
group, kw = P2PGroup.from_torch_distributed(cfg, B, P, T, dev)
pol = Policy.random_init(cfg, dev, B, P, T, seed=7, **kw)      # identical base + LoRA on every learner
group.attach(pol)
ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-3})
ln.enable_p2p(group)
ok = True
p_ref = pol.lora_flat.clone(); m_ref = torch.zeros_like(p_ref); v_ref = torch.zeros_like(p_ref)
for step in range(1, 4):
    _, (prompts, answers, adv) = synthetic_candidates(cfg.vocab, N, P, T, 4, seed=100 * step + rank)   # different shard per rank
    loss = ln.compute_loss(prompts, answers, adv)
    g_local = pol.lora_grad.clone()
    # reference exchange: NCCL mean + single-learner Adam on a private copy
    g_mean = g_local.clone()
    dist.all_reduce(g_mean)
    g_mean /= world
    ops.adamw_step(p_ref, m_ref, v_ref, g_mean.clone(), step, 1e-3, zero_grad=False)
    ln.apply_merged_gradients()                                   # fused P2P reduce + Adam + write-back
    torch.cuda.synchronize()
    diff = (pol.lora_flat - p_ref).abs().max().item()
    # all ranks identical?
    mx, mn = pol.lora_flat.clone(), pol.lora_flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    same = bool((mx == mn).all())
    gz = bool((pol.lora_grad == 0).all())
    if rank == 0:
        print(f"step {step}: loss {loss:+.5f}  max|p2p - nccl| = {diff:.3e}  identical_across_ranks={same}  grads_zeroed={gz}", flush=True)
    # NCCL sums in a different order than the rank-ordered P2P sum: allow fp32 reassociation noise on the mean
    ok = ok and diff < 5e-6 and same and gz
# timing of the exchange alone on the real 7B-sized buffer (40.37 M fp32)
if rank == 0:
    print("P2P_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
