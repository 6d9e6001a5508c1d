"""Parity at the shapes bench.py measures (BASELINE config 2): full-width Qwen2.5-7B layers (hidden 3584, inter 18944,
28 q / 4 kv heads of 128, V = 152064, rank-16 LoRA), P = 350, T = 512, micro-batch 8, 16 sequences in 2 groups —
through the exact bench path: packed shared-prompt rows, two reference micro-batches fused into one model pass, resident
bf16 weight cache, grouped dW, SwiGLU in the GEMM epilogues, GEMM tail split.  Only the layer COUNT is reduced (1 and 2
instead of 28) so that the checker — oracle/learner_oracle.py (the restatement of distributed_actor.py:215-261, :440-493
that tests/test_oracle.py pins to the reference's own outputs), run in fp32 on the same GPU — finishes in seconds.

Tolerances (bf16 tensor-core path vs fp32 oracle): loss |d| <= 2e-2*|loss| + 2e-3 (GRPO loss value is -mean(adv): exact);
LoRA gradients: global cosine >= 0.999, global rel-L2 <= 3e-2, per-tensor cosine >= 0.99;
per-token log-prob max |d| <= 8e-2, mean |d| <= 2e-2.  The log-prob bound is wider than at the toy shapes of
test_gpu_learner.py (4e-2 / 6e-3) for a stated reason: the logits are a bf16 tensor here exactly as under the reference's
autocast (distributed_actor.py:241-243, :462), and with hidden 3584 they reach |z| ~ 4-8, where one bf16 ulp is
1.6e-2 - 3.1e-2; lp = z_y - logsumexp(z) inherits the rounding of z_y (measured on B200: max 0.041 / 0.057,
mean 0.010 / 0.013 for the two cases below).  The gradients, which average over 152064 logits per token, stay at 3e-2.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import learner_oracle as lo  # noqa: E402  (checker only)
from tests.policy_bridge import flat_grad_from_oracle, oracle_config, oracle_params_from_policy  # noqa: E402

P, T, B, N, GROUP = 350, 512, 8, 16, 8


def _build(cuda, n_layers, kind, seed):
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer, Learner
    from distrl_llm_b200.policy import LMConfig, Policy
    cfg = LMConfig.qwen25_7b(lora_r=16)
    cfg.n_layers = n_layers
    pol = Policy.random_init(cfg, cuda, 2 * B, P, T, seed=seed)          # capacity for 2 fused micro-batches, like bench.py
    cls = GRPOLearner if kind == "grpo" else Learner
    ln = cls(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 2e-5})
    # the configuration bench.py runs
    assert ln.share_prompts and ln.ragged_rows and ln.fuse_microbatches == 2 and pol.weight_cache is not None
    return cfg, pol, ln


def _check_grads(pol, ref_flat, cos_min=0.999, rel_max=3e-2):
    got = pol.lora_grad.double()
    for (i, m, ab), (off, shp) in pol.offsets.items():
        n = shp[0] * shp[1]
        a, b = got[off:off + n], ref_flat[off:off + n]
        if b.norm() > 0:
            c = (a @ b) / (a.norm() * b.norm() + 1e-300)
            assert c > 0.99, f"l{i}.{m}.{ab}: cosine {c.item():.5f}"
    cos = ((got @ ref_flat) / (got.norm() * ref_flat.norm())).item()
    rel = ((got - ref_flat).norm() / ref_flat.norm()).item()
    assert cos >= cos_min and rel <= rel_max, f"global cosine {cos:.6f}, rel-L2 {rel:.4f}"
    return cos, rel


@pytest.mark.parametrize("n_layers,ragged,kind", [(1, True, "pg"), (2, False, "grpo")],
                         ids=["1layer_ragged_pg", "2layer_full_grpo"])
def test_cfg2_shapes_learner_vs_oracle(cuda, n_layers, ragged, kind):
    from distrl_llm_b200.trainer_prep import synthetic_candidates
    cfg, pol, ln = _build(cuda, n_layers, kind, seed=7 + n_layers)
    _, (prompts, answers, adv) = synthetic_candidates(cfg.vocab, N, P, T, GROUP, seed=99, ragged=ragged)
    rewards = np.asarray(adv, dtype=np.float64)
    if kind == "pg":   # PG consumes (reward - baseline); any non-zero per-sequence weights exercise the same path
        rewards = rewards * 0.37 + 0.05
    _, loss = ln._compute_gradients(prompts, answers, list(rewards), export=False)
    lp, mask = ln.compute_current_policy_probs(pol, prompts[:B], answers[:B])
    torch.cuda.synchronize()

    ocfg = oracle_config(cfg)
    params = oracle_params_from_policy(pol)
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    ids, am, ansm = ids.to(cuda), am.to(cuda), ansm.to(cuda)
    ref_grads, ref_loss = lo.compute_gradients(params, ocfg, ids, am, ansm, rewards, P, B, kind)
    loss_tol = 4e-2 * sum(np.abs(rewards[i:i + B]).mean() for i in range(0, N, B)) + 2e-3 if kind == "pg" else 1e-9
    assert abs(loss - ref_loss) <= loss_tol, (loss, ref_loss)
    _check_grads(pol, flat_grad_from_oracle(pol, ref_grads))
    with torch.no_grad():
        lp_ref = lo.compute_current_policy_probs(params, ocfg, ids[:B], am[:B], P)
    m = mask.bool()
    d = (lp[m] - lp_ref[m]).abs()
    assert d.max().item() < 8e-2 and d.mean().item() < 2e-2, (d.max().item(), d.mean().item())


def test_cfg2_shapes_classic_layout_matches_packed(cuda):
    """Same full-width layer through the classic [B, P+T] layout (bench.py --no_share_prompts) and one micro-batch per
    pass: gradients equal the packed / fused path up to bf16 reassociation."""
    from distrl_llm_b200.trainer_prep import synthetic_candidates
    cfg, pol, ln = _build(cuda, 1, "grpo", seed=21)
    _, (prompts, answers, adv) = synthetic_candidates(cfg.vocab, N, P, T, GROUP, seed=5, ragged=True)
    ln._compute_gradients(prompts, answers, list(adv), export=False)
    g_packed = pol.lora_grad.double().clone()
    ln.share_prompts = False
    ln.fuse_microbatches = 1
    ln._compute_gradients(prompts, answers, list(adv), export=False)
    _check_grads(pol, g_packed, cos_min=0.9995, rel_max=3e-2)
