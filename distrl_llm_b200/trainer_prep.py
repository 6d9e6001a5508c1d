"""Trainer-side preparation of the learner batch: group-relative advantages + top-k subselect through
the G9 CUDA kernel (reference Trainer.train, distributed_trainer.py:262-294), the multi-learner split
(:308-322), and synthetic candidates for benchmarks."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def apply_advantages_and_topk(candidates, learner_type, topk, device):
    """In-place equivalent of distributed_trainer.py:262-294 on the reference's `candidates` payload (list
    of dicts with 'rewards': list of np [C,2], 'answers'/'problem': [n_prob][C]).  Numerics run in the fp64
    G9 kernel; only the string re-ordering stays on the host."""
    for cand in candidates:
        rw = np.stack([np.asarray(r, dtype=np.float64) for r in cand["rewards"]])      # [n_prob, C, 2]
        vals, base, idx, val = ops.group_advantage_topk(torch.from_numpy(rw).to(device), topk, learner_type == "grpo")
        idx_h, val_h, base_h = idx.cpu().numpy(), val.cpu().numpy(), base.cpu().numpy()
        if learner_type != "grpo":
            cand["baselines"] = [b for b in base_h]                                       # :278
        cand["answers"] = [[cand["answers"][j][i] for i in idx_h[j]] for j in range(len(idx_h))]   # :289
        cand["rewards"] = [val_h[j] for j in range(len(idx_h))]                           # :290
        cand["problem"] = [cand["problem"][j][:topk] for j in range(len(idx_h))]          # :291
    return candidates


def merge_candidates(candidates):
    """Trainer.merge_candidates (distributed_trainer.py:221-230) — note it drops 'baselines' (quirk Q5)."""
    problems, answers, rewards = [], [], []
    for cand in candidates:
        for a, p, r in zip(cand["answers"], cand["problem"], cand["rewards"]):
            problems.extend(p)
            answers.extend(a)
            rewards.extend(r)
    return problems, answers, rewards


def split_for_learners(problems, answers, rewards, n_learners):
    """Even split, remainder to the first learners (distributed_trainer.py:312-322)."""
    sizes = [len(problems) // n_learners] * n_learners
    for i in range(len(problems) % n_learners):
        sizes[i] += 1
    chunks, start = [], 0
    for s in sizes:
        chunks.append((problems[start:start + s], answers[start:start + s], rewards[start:start + s]))
        start += s
    return chunks


def synthetic_candidates(vocab, n_seq, P, T, group_size, seed=1234, device=None, ragged=False):
    """Synthetic learner batch: token ids ~ U[1, V), full-length prompts and completions,
    rewards = format in {0,.1,.2} w.p. (.5,.3,.2) + accuracy ~ Bernoulli(.25) per candidate (degenerate groups
    redrawn), advantages by the GRPO rule.  Returns (candidates payload, (prompts, answers, advantages))."""
    rng = np.random.default_rng(seed)
    n_prob = n_seq // group_size
    cand = {"answers": [], "problem": [], "rewards": []}
    for _ in range(n_prob):
        # ragged variant: prompt_len ~ U[P/2, P], completion_len ~ U[T/4, T]
        prompt = rng.integers(1, vocab, size=int(rng.integers(P // 2, P + 1)) if ragged else P).tolist()
        cand["problem"].append([prompt] * group_size)
        cand["answers"].append([rng.integers(1, vocab, size=int(rng.integers(T // 4, T + 1)) if ragged else T).tolist()
                                for _ in range(group_size)])
        while True:
            fmt = rng.choice([0.0, 0.1, 0.2], size=group_size, p=[0.5, 0.3, 0.2])
            acc = (rng.random(group_size) < 0.25).astype(np.float64)
            s = fmt + acc
            adv = (s - np.mean(s)) / (np.std(s) + 1e-8)   # distributed_trainer.py:273
            # redraw degenerate groups AND groups with an exactly-zero advantage: the reference skips every
            # micro-batch that contains one (quirk Q1, distributed_actor.py:459), which would silently
            # remove work from a benchmark step
            if np.std(s) > 0 and np.all(adv != 0):
                break
        cand["rewards"].append(adv)
    problems, answers, rewards = merge_candidates([cand])
    return [cand], (problems, answers, np.asarray(rewards, dtype=np.float64))
