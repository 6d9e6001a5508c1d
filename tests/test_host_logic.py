"""CPU tests of the host-side logic around the hot path: tokenise/pad rules, micro-batch skip predicate,
candidate merge / split, slice ownership of the P2P reduce, and the N>1 semantics on 2 gloo ranks."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import learner_oracle as lo
from tests.golden_utils import load_cfg1


def test_id_tokenizer_matches_reference_padding_rules():
    from distrl_llm_b200.learner import IdTokenizer
    tok = IdTokenizer()
    P, T = 6, 5
    prompts = [[1, 2, 3], [4, 5, 6, 7, 8, 9, 10, 11], []]
    answers = [[21, 22], [23, 24, 25, 26, 27, 28], [29]]
    a = tok.batch_encode_plus(prompts, padding="max_length", padding_side="left", max_length=P, truncation=True)
    b = tok.batch_encode_plus(answers, padding="max_length", padding_side="right", max_length=T, truncation=True)
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)   # oracle = restatement of distributed_actor.py:217-239
    assert torch.equal(torch.cat([a["input_ids"], b["input_ids"]], 1).long(), ids)
    assert torch.equal(torch.cat([a["attention_mask"], b["attention_mask"]], 1).long(), am)
    assert a["input_ids"][1].tolist() == [4, 5, 6, 7, 8, 9]          # truncation keeps the first tokens
    assert a["attention_mask"][0].tolist() == [0, 0, 0, 1, 1, 1]     # left padding
    assert b["attention_mask"][0].tolist() == [1, 1, 0, 0, 0]        # right padding


@pytest.mark.parametrize("r,skip", [([1.0, 0.0, 3.0], True), ([-1.0, 0.5, 1e-9], False), ([0.0, 0.0], True), ([2.0], False)])
def test_q1_skip_predicate(r, skip):
    """reference `if batch_rewards.all() == 0: continue` (distributed_actor.py:367/:459) vs the learner's numpy form."""
    ref = bool(torch.tensor(r, dtype=torch.float64).all() == 0)
    ours = not bool(np.all(np.asarray(r) != 0))
    assert ref == ours == skip


def test_merge_and_split_like_trainer():
    from distrl_llm_b200 import trainer_prep as tp
    cands = [{"answers": [["a0", "a1"], ["b0", "b1"]], "problem": [["p", "p"], ["q", "q"]], "rewards": [np.array([1., 2.]), np.array([3., 4.])],
              "baselines": [1.5, 3.5]},
             {"answers": [["c0", "c1"]], "problem": [["r", "r"]], "rewards": [np.array([5., 6.])]}]
    p, a, r = tp.merge_candidates(cands)           # drops 'baselines' like distributed_trainer.py:221-230 (quirk Q5)
    assert a == ["a0", "a1", "b0", "b1", "c0", "c1"] and p == ["p", "p", "q", "q", "r", "r"] and r == [1, 2, 3, 4, 5, 6]
    chunks = tp.split_for_learners(p, a, r, 4)
    assert [len(c[0]) for c in chunks] == [2, 2, 1, 1]
    assert [(s, n) for s, n in lo.split_evenly(6, 4)] == [(0, 2), (2, 2), (4, 1), (5, 1)]
    assert sum((c[1] for c in chunks), []) == a


def test_synthetic_candidates_never_trigger_q1():
    from distrl_llm_b200 import trainer_prep as tp
    for seed in range(20):
        cands, (p, a, r) = tp.synthetic_candidates(1000, 64, 4, 6, 8, seed=seed)
        assert len(p) == len(a) == len(r) == 64
        assert np.all(r != 0)
        for g in range(8):
            grp = r[g * 8:(g + 1) * 8]
            assert abs(grp.mean()) < 1e-6 and abs(grp.std() - 1) < 1e-6   # GRPO-normalised per group


@pytest.mark.parametrize("n,world", [(40370176, 2), (40370176, 8), (1004, 3), (8, 8), (4, 2)])
def test_owned_slices_partition_the_buffer(n, world):
    from distrl_llm_b200.p2p import owned_slice
    cover = 0
    prev_hi = 0
    for r in range(world):
        lo_, hi = owned_slice(n, world, r)
        assert lo_ == prev_hi and lo_ % 4 == 0 and hi % 4 == 0 and hi >= lo_
        prev_hi = hi
        cover += hi - lo_
    assert cover == n and prev_hi == n


# ---- N > 1 path on 2 CPU ranks (gloo): shard -> per-learner gradients -> mean -> Adam == reference golden ----
def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    z, cfg, params, nf4, prompts, answers = load_cfg1()
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    r = z["rewards"] - z["baseline"]
    start, size = lo.split_evenly(len(prompts), world)[rank]        # distributed_trainer.py:312-322
    ids, am, ansm = lo.pad_batch(prompts[start:start + size], answers[start:start + size], P, T)
    grads, _ = lo.compute_gradients(params, cfg, ids, am, ansm, r[start:start + size], P, B, "pg")
    names = lo.lora_names(cfg)
    flat = torch.cat([grads[n].flatten() for n in names])
    dist.all_reduce(flat)                      # what the P2P reduce kernel computes: sum over learners ...
    flat /= world                              # ... divided by the number of learners (distributed_actor.py:323)
    merged, off = {}, 0
    for n in names:
        k = grads[n].numel()
        merged[n] = flat[off:off + k].view_as(grads[n])
        off += k
    lo.adam_step(params, merged, {}, lr=2e-5)  # EVERY rank steps (unlike the reference's learner-0-only step, quirk Q4)
    err = max((params[n].data - torch.from_numpy(z[f"fp32.merged_step.{n}"])).abs().max().item() for n in names)
    out[rank] = err
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_and_step_matches_reference_golden():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        assert out[rank] < 2e-7, f"rank {rank}: {out[rank]}"   # every learner ends with the reference's merged step
