"""CPU tests: the oracle restatement (oracle/learner_oracle.py) is pinned against golden vectors
produced by the reference's own code (oracle/make_golden.py), so that GPU parity tests can trust it."""
import os

import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from tests.golden_utils import GOLDEN, load_cfg1


@pytest.fixture(scope="module")
def g1():
    return load_cfg1()


def _grads(z, mode, kind, cfg):
    return {n: torch.from_numpy(z[f"{mode}.{kind}.grad.{n}"]) for n in lo.lora_names(cfg)}


@pytest.mark.parametrize("kind", ["pg", "grpo"])
def test_compute_gradients_matches_reference_fp32(g1, kind):
    z, cfg, params, nf4, prompts, answers = g1
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    r = z["rewards"] - z["baseline"] if kind == "pg" else z["grpo_adv"]
    grads, loss = lo.compute_gradients(params, cfg, ids, am, ansm, r, P, B, kind)
    ref = _grads(z, "fp32", kind, cfg)
    assert abs(loss - float(z[f"fp32.{kind}.loss"])) < 1e-6 * max(1.0, abs(loss)) + 1e-7
    num = sum((grads[n] - ref[n]).double().pow(2).sum() for n in ref).sqrt()
    den = sum(ref[n].double().pow(2).sum() for n in ref).sqrt()
    assert den > 0
    assert (num / den).item() < 2e-5, f"global grad rel-L2 {num / den}"
    for n in ref:  # per tensor as well
        d = (grads[n] - ref[n]).norm() / (ref[n].norm() + 1e-12)
        assert d < 5e-4, (n, d.item())


def test_pg_and_grpo_share_the_gradient(g1):
    """the GRPO surrogate exp(lp - lp.detach()) has value -mean(A) and the PG gradient."""
    z, cfg, *_ = g1
    # identical rewards would give identical grads; here rewards differ (baseline vs normalised), so check the
    # loss-value identity only: GRPO loss = sum over micro-batches of -mean(adv)
    adv = z["grpo_adv"]
    B = int(z["train_batch_size"])
    exp = sum(-adv[i:i + B].mean() for i in range(0, len(adv), B))
    assert abs(float(z["fp32.grpo.loss"]) - exp) < 1e-6


def test_logprobs_match_reference(g1):
    z, cfg, params, nf4, prompts, answers = g1
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ids, am, ansm = lo.pad_batch(prompts[:B], answers[:B], P, T)
    with torch.no_grad():
        lp = lo.compute_current_policy_probs(params, cfg, ids, am, P)
    assert np.array_equal(ansm.numpy(), z["answer_mask_mb0"])
    m = ansm.bool()
    assert (lp[m] - torch.from_numpy(z["fp32.logp_mb0"])[m]).abs().max() < 2e-5


def test_quirk_q1_any_zero_reward_skips_microbatch(g1):
    z, cfg, params, nf4, prompts, answers = g1
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    grads, loss = lo.compute_gradients(params, cfg, ids, am, ansm, z["fp32.q1.rewards"], P, B, "grpo")
    assert abs(loss - float(z["fp32.q1.loss"])) < 1e-7
    ref = torch.from_numpy(z["fp32.q1.grad.l0.q.B"])
    assert (grads["l0.q.B"] - ref).norm() / ref.norm() < 5e-4
    # without the quirk the first micro-batch (rewards [0.5, 0.0]) is trained too -> different loss
    _, loss_fixed = lo.compute_gradients(params, cfg, ids, am, ansm, z["fp32.q1.rewards"], P, B, "grpo",
                                         reference_quirks=False)
    assert abs(loss_fixed - loss) > 1e-3


def test_merge_and_adam_step_match_reference(g1):
    """Two learners' gradients -> reference apply_merged_gradients (mean, then Adam) vs the oracle's."""
    z, cfg, params, nf4, prompts, answers = g1
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    r = z["rewards"] - z["baseline"]
    params = {k: (v.detach().clone().requires_grad_(v.requires_grad)) for k, v in params.items()}
    gds = []
    for sl in (slice(0, 2), slice(2, 4)):
        ids, am, ansm = lo.pad_batch(prompts[sl], answers[sl], P, T)
        g, _ = lo.compute_gradients(params, cfg, ids, am, ansm, r[sl], P, B, "pg")
        gds.append(g)
    merged = lo.merge_gradients(gds)
    lo.adam_step(params, merged, {}, lr=2e-5)
    for n in lo.lora_names(cfg):
        ref = torch.from_numpy(z[f"fp32.merged_step.{n}"])
        assert torch.allclose(params[n].data, ref, rtol=0, atol=2e-7), n


def test_bf16_autocast_golden_is_close_to_fp32(g1):
    """Bounds the precision gap the GPU (bf16) parity tolerances have to absorb."""
    z, cfg, *_ = g1
    for kind in ("pg", "grpo"):
        a, b = _grads(z, "bf16", kind, cfg), _grads(z, "fp32", kind, cfg)
        va = torch.cat([a[n].flatten() for n in a]).double()
        vb = torch.cat([b[n].flatten() for n in b]).double()
        cos = (va @ vb) / (va.norm() * vb.norm())
        assert cos > 0.999


# ---- Trainer advantage / top-k block (executed from the reference's source by make_golden.py) ----
def test_advantages_topk_match_reference_block():
    z = np.load(os.path.join(GOLDEN, "trainer_advantages.npz"))
    for case in range(4):
        rewards = z[f"c{case}.rewards"]
        topk = int(z[f"c{case}.topk"])
        for lt in ("grpo", "pg"):
            ref_vals = z[f"c{case}.{lt}.filtered_rewards"]
            ref_ans = z[f"c{case}.{lt}.filtered_answers"]
            for j in range(rewards.shape[0]):
                vals, base = lo.group_advantages(rewards[j], lt)
                idx = lo.topk_filter(vals, topk)
                assert np.array_equal(vals[idx], ref_vals[j]), "bit-exact float64"
                assert [f"a{j}_{c}" for c in idx] == list(ref_ans[j])
                if lt == "pg":
                    assert base == z[f"c{case}.pg.baselines"][j]


def test_split_evenly_like_trainer():
    assert lo.split_evenly(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert lo.split_evenly(8, 2) == [(0, 4), (4, 4)]
    assert lo.split_evenly(3, 4) == [(0, 1), (1, 1), (2, 1), (3, 0)]


def test_nf4_restatement_roundtrip():
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((64, 128)) * 0.02).astype(np.float32)
    packed, absmax = lo.nf4_quantize(w)
    deq = lo.nf4_dequantize(packed, absmax, w.shape).float().numpy()
    # every value maps to its nearest level; error bounded by half the largest level gap * absmax (+ bf16 rounding)
    gap = np.diff(lo.NF4_LEVELS).max() / 2
    am = np.repeat(absmax, 64).reshape(w.shape)
    assert (np.abs(deq - w) <= gap * am + np.abs(deq) * 2.0 ** -8 + 1e-7).all()  # + bf16 rounding of the output
    # exact levels survive
    lv = (lo.NF4_LEVELS[None, :].repeat(4, 0).reshape(-1) * 0.5).astype(np.float32)
    p2, a2 = lo.nf4_quantize(lv)
    assert np.array_equal(lo.nf4_dequantize(p2, a2, lv.shape).float().numpy(),
                          torch.from_numpy(lv).to(torch.bfloat16).float().numpy())
