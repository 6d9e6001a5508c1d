"""End-to-end parity (GPU): the libb200rl learner (through the reference-shaped Learner / GRPOLearner
classes and the C ABI) against
  (a) the committed golden vectors produced by the REFERENCE's own code (tests/golden/cfg1_learner.npz,
      BASELINE config 1), and
  (b) the pinned oracle (oracle/learner_oracle.py) on larger seeded synthetic batches.

Tolerances (floating point; the CUDA path computes in bf16 with fp32 accumulation like the reference's
autocast, the oracle/golden in fp32):
  per-token log-prob  |d| <= 4e-2 ;  loss |d| <= 2e-2*|loss| + 2e-3 ;
  LoRA grads: global cosine >= 0.999 and global rel-L2 <= 5e-2, per-tensor cosine >= 0.99.
The bf16-autocast golden (also generated from the reference) differs from its own fp32 golden by a
similar amount (tests/test_oracle.py::test_bf16_autocast_golden_is_close_to_fp32).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import learner_oracle as lo  # noqa: E402  (checker only)
from tests.golden_utils import load_cfg1  # noqa: E402


def _mk_cfg(ocfg):
    from distrl_llm_b200.policy import LMConfig
    return LMConfig(vocab=ocfg.vocab, hidden=ocfg.hidden, inter=ocfg.inter, n_layers=ocfg.n_layers,
                    n_q_heads=ocfg.n_q_heads, n_kv_heads=ocfg.n_kv_heads, head_dim=ocfg.head_dim,
                    lora_r=ocfg.lora_r, lora_alpha=ocfg.lora_alpha, rms_eps=ocfg.rms_eps, rope_theta=ocfg.rope_theta)


def _mk_learner(kind, ocfg, params, nf4, P, T, B, device, lr=2e-5, quirks=True):
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer, Learner
    from distrl_llm_b200.policy import Policy
    pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, device, max_batch=B, P=P, T=T)
    config = {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": lr}
    cls = Learner if kind == "pg" else GRPOLearner
    return cls(pol, IdTokenizer(), config, reference_quirks=quirks)


def _compare_grads(got: dict, ref: dict, ocfg, pol, cos_min=0.999, rel_max=5e-2):
    va, vb = [], []
    for i in range(ocfg.n_layers):
        for m in lo.LORA_MODULES:
            for ab in ("A", "B"):
                a = got[pol.peft_name(i, m, ab)].double().flatten()
                b = ref[f"l{i}.{m}.{ab}"].double().flatten().cpu()
                if b.norm() > 0:
                    c = (a @ b) / (a.norm() * b.norm() + 1e-300)
                    assert c > 0.99, f"l{i}.{m}.{ab} cosine {c.item():.5f}"
                va.append(a)
                vb.append(b)
    va, vb = torch.cat(va), torch.cat(vb)
    cos = ((va @ vb) / (va.norm() * vb.norm())).item()
    rel = ((va - vb).norm() / vb.norm()).item()
    assert cos >= cos_min and rel <= rel_max, f"global cosine {cos:.6f}, rel-L2 {rel:.4f}"
    return cos, rel


# ---------------------------------------------------------------------------------------------------
# (a) golden vectors from the reference's own code — BASELINE config 1
# ---------------------------------------------------------------------------------------------------
def test_cfg1_logprobs_vs_reference_golden(cuda):
    z, ocfg, params, nf4, prompts, answers = load_cfg1()
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ln = _mk_learner("pg", ocfg, params, nf4, P, T, B, cuda)
    lp, mask = ln.compute_current_policy_probs(ln.policy, prompts[:B], answers[:B])
    assert np.array_equal(mask.cpu().numpy(), z["answer_mask_mb0"])
    m = mask.bool().cpu()
    d = (lp.cpu()[m] - torch.from_numpy(z["fp32.logp_mb0"])[m]).abs().max().item()
    assert d < 4e-2, f"max |dlogp| {d}"


@pytest.mark.parametrize("kind", ["pg", "grpo"])
def test_cfg1_gradients_vs_reference_golden(cuda, kind):
    z, ocfg, params, nf4, prompts, answers = load_cfg1()
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ln = _mk_learner(kind, ocfg, params, nf4, P, T, B, cuda)
    r = z["rewards"] - z["baseline"] if kind == "pg" else z["grpo_adv"]
    grads, loss = ln._compute_gradients(prompts, answers, list(r))
    # PG loss = -mean(lp_seq * r) with sum(r) = 0 is a small difference of O(6) numbers: its error bound is
    # max|dlp| * mean|r| per micro-batch (the bf16 golden itself is 6e-3 away from the fp32 golden)
    loss_tol = 4e-2 * sum(np.abs(r[i:i + B]).mean() for i in range(0, len(r), B)) + 2e-3
    for mode in ("fp32", "bf16"):
        ref_loss = float(z[f"{mode}.{kind}.loss"])
        assert abs(loss - ref_loss) <= loss_tol, (mode, loss, ref_loss)
        ref = {n: torch.from_numpy(z[f"{mode}.{kind}.grad.{n}"]) for n in lo.lora_names(ocfg)}
        _compare_grads(grads, ref, ocfg, ln.policy)


def test_cfg1_quirk_q1_skip(cuda):
    z, ocfg, params, nf4, prompts, answers = load_cfg1()
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ln = _mk_learner("grpo", ocfg, params, nf4, P, T, B, cuda)
    grads, loss = ln._compute_gradients(prompts, answers, list(z["fp32.q1.rewards"]))
    assert abs(loss - float(z["fp32.q1.loss"])) < 1e-9  # GRPO loss value = -mean(adv) of the trained micro-batch: exact
    ref = torch.from_numpy(z["fp32.q1.grad.l0.q.B"]).double().flatten()
    got = grads[ln.policy.peft_name(0, "q", "B")].double().flatten()
    assert (got @ ref) / (got.norm() * ref.norm()) > 0.999


def test_cfg1_merge_and_step_vs_reference_golden(cuda):
    """Two learners' gradient dicts through apply_merged_gradients (reference :302-333 semantics)."""
    z, ocfg, params, nf4, prompts, answers = load_cfg1()
    P, T, B = int(z["P"]), int(z["T"]), int(z["train_batch_size"])
    ln = _mk_learner("pg", ocfg, params, nf4, P, T, B, cuda)
    r = list(z["rewards"] - z["baseline"])
    g1, _ = ln._compute_gradients(prompts[:2], answers[:2], r[:2])
    g2, _ = ln._compute_gradients(prompts[2:], answers[2:], r[2:])
    before = ln.policy.lora_state_dict()
    ln.apply_merged_gradients([g1, g2])
    after = ln.policy.lora_state_dict()
    # Adam's first step moves every weight by ~lr*sign(g): compare the UPDATE direction with the reference's
    num = den = 0.0
    for i in range(ocfg.n_layers):
        for m in lo.LORA_MODULES:
            for ab in ("A", "B"):
                k = ln.policy.peft_name(i, m, ab)
                ref_delta = torch.from_numpy(z[f"fp32.merged_step.l{i}.{m}.{ab}"]) - torch.from_numpy(z[f"param.l{i}.{m}.{ab}"])
                d = (after[k] - before[k]).double()
                num += (d.flatten() @ ref_delta.double().flatten()).item()
                den += d.norm().item() ** 2
                assert (d.abs() <= 2e-5 * 1.001).all()  # |step| <= lr on step 1
    assert num / den > 0.97, "merged Adam update must point the same way as the reference's"
    assert (ln.policy.lora_grad == 0).all()


# ---------------------------------------------------------------------------------------------------
# (b) oracle on larger seeded batches (ragged lengths, several micro-batches, head_dim 64 and 128)
# ---------------------------------------------------------------------------------------------------
CASES = [
    # name, cfg kwargs, N, P, T, B, learner
    ("hd64_grpo", dict(vocab=8192, hidden=512, inter=1024, n_layers=4, n_q_heads=8, n_kv_heads=2, head_dim=64, lora_r=16, lora_alpha=16), 6, 24, 40, 4, "grpo"),
    ("hd128_pg", dict(vocab=4096, hidden=512, inter=1536, n_layers=3, n_q_heads=4, n_kv_heads=2, head_dim=128, lora_r=16, lora_alpha=32), 5, 70, 130, 2, "pg"),
    ("r32_grpo", dict(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=4, head_dim=64, lora_r=32, lora_alpha=16), 8, 16, 48, 8, "grpo"),
]


@pytest.mark.parametrize("name,ckw,N,P,T,B,kind", CASES, ids=[c[0] for c in CASES])
def test_learner_vs_oracle(cuda, name, ckw, N, P, T, B, kind):
    ocfg = lo.OracleConfig(**ckw)
    params, nf4 = lo.make_params(ocfg, seed=11)
    prompts, answers, rewards = lo.make_batch(ocfg, N, P, T, seed=3, ragged=True, group_size=N, learner=kind)
    # oracle in fp32 on the GPU (checker only; same code that is pinned on CPU against the goldens)
    dparams = {k: (v.detach().to(cuda).requires_grad_(v.requires_grad)) for k, v in params.items()}
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    ref_grads, ref_loss = lo.compute_gradients(dparams, ocfg, ids.to(cuda), am.to(cuda), ansm.to(cuda), rewards, P, B, kind)
    ln = _mk_learner(kind, ocfg, params, nf4, P, T, B, cuda)
    grads, loss = ln._compute_gradients(prompts, answers, list(rewards))
    loss_tol = 4e-2 * sum(np.abs(rewards[i:i + B]).mean() for i in range(0, N, B)) + 2e-3
    assert abs(loss - ref_loss) <= loss_tol, (loss, ref_loss)
    _compare_grads(grads, ref_grads, ocfg, ln.policy)
    # scoring-only API agrees with the oracle's per-token log-probs
    with torch.no_grad():
        lp_ref = lo.compute_current_policy_probs(dparams, ocfg, ids[:B].to(cuda), am[:B].to(cuda), P)
    lp, mask = ln.compute_current_policy_probs(ln.policy, prompts[:B], answers[:B])
    m = mask.bool()
    assert (lp[m] - lp_ref[m]).abs().max().item() < 4e-2


def test_train_step_updates_like_adam(cuda):
    """GRPOLearner.train (reference :495-514): after the step, params == Adam(params, our grads) and the
    bf16 operand copies are refreshed (a second scoring call sees the new adapter)."""
    ocfg = lo.OracleConfig(vocab=1024, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=5)
    P, T, B = 8, 24, 4
    prompts, answers, rewards = lo.make_batch(ocfg, 4, P, T, seed=9, ragged=True, group_size=4, learner="grpo")
    ln = _mk_learner("grpo", ocfg, params, nf4, P, T, B, cuda, lr=1e-3)
    lp0, _ = ln.compute_current_policy_probs(ln.policy, prompts, answers)
    g, _ = ln._compute_gradients(prompts, answers, list(rewards), export=False)
    gflat = ln.policy.lora_grad.clone()
    p0 = ln.policy.lora_flat.clone()
    cand = [{"answers": [answers], "problem": [prompts], "rewards": [rewards]}]
    ln.train(cand)
    refp = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([refp], lr=1e-3, foreach=False, fused=False)
    refp.grad = gflat
    opt.step()
    assert torch.allclose(ln.policy.lora_flat, refp.data, rtol=1e-6, atol=1e-9)
    lp1, mask = ln.compute_current_policy_probs(ln.policy, prompts, answers)
    assert (lp1 - lp0)[mask.bool()].abs().max() > 1e-4  # adapter change is visible to the next forward


def test_kl_term_vs_oracle(cuda):
    """Optional KL(pi || pi_ref) term (north_star; absent from the reference => oracle restates OUR definition,
    'parity unpinned'): pi_ref = adapter-disabled forward, k3 estimator, beta = 0.2."""
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=32)
    params, nf4 = lo.make_params(ocfg, seed=21, lora_b_std=0.05)   # larger B so that pi differs visibly from pi_ref
    N, P, T, B, beta = 6, 12, 36, 3, 0.2
    prompts, answers, rewards = lo.make_batch(ocfg, N, P, T, seed=5, ragged=True, group_size=N, learner="grpo")
    dparams = {k: (v.detach().to(cuda).requires_grad_(v.requires_grad)) for k, v in params.items()}
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    ref_grads, ref_loss = lo.compute_gradients(dparams, ocfg, ids.to(cuda), am.to(cuda), ansm.to(cuda), rewards, P, B,
                                               "grpo", kl_beta=beta)
    g0, loss0 = lo.compute_gradients(dparams, ocfg, ids.to(cuda), am.to(cuda), ansm.to(cuda), rewards, P, B, "grpo")
    assert abs(ref_loss - loss0) > 1e-5, "the KL term must be active in this test"
    ln = _mk_learner("grpo", ocfg, params, nf4, P, T, B, cuda)
    ln.kl_beta = beta
    grads, loss = ln._compute_gradients(prompts, answers, list(rewards))
    assert abs(loss - ref_loss) <= 5e-2 * abs(ref_loss - loss0) + 2e-3, (loss, ref_loss, loss0)
    _compare_grads(grads, ref_grads, ocfg, ln.policy)
    # and beta = 0 falls back to the reference loss exactly (GRPO loss value = -mean(adv))
    ln.kl_beta = 0.0
    _, loss_b0 = ln._compute_gradients(prompts, answers, list(rewards))
    assert abs(loss_b0 - loss0) < 1e-9


@pytest.mark.parametrize("kind,beta", [("grpo", 0.0), ("pg", 0.0), ("grpo", 0.2)])
def test_shared_prompt_packing_vs_oracle(cuda, kind, beta):
    """Packed shared-prompt layout (packing.py): 3 problems x 4 completions with identical prompts inside a group,
    micro-batches of 6 (so one group straddles two micro-batches);
    the oracle computes every (prompt, completion) pair separately like the reference."""
    ocfg = lo.OracleConfig(vocab=4096, hidden=512, inter=1024, n_layers=3, n_q_heads=4, n_kv_heads=2, head_dim=128,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=31, lora_b_std=0.05 if beta else 0.01)
    P, T, B, G, n = 40, 72, 6, 3, 4
    rng = np.random.default_rng(17)
    prompts, answers = [], []
    for g in range(G):
        pr = rng.integers(1, ocfg.vocab, size=int(rng.integers(P // 2, P + 1))).tolist()
        for _ in range(n):
            prompts.append(pr)
            answers.append(rng.integers(1, ocfg.vocab, size=int(rng.integers(T // 4, T + 1))).tolist())
    _, _, rewards = lo.make_batch(ocfg, G * n, P, T, seed=5, group_size=n, learner=kind)
    dparams = {k: (v.detach().to(cuda).requires_grad_(v.requires_grad)) for k, v in params.items()}
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    kw = {"kl_beta": beta} if beta else {}
    ref_grads, ref_loss = lo.compute_gradients(dparams, ocfg, ids.to(cuda), am.to(cuda), ansm.to(cuda), rewards, P, B, kind, **kw)
    ln = _mk_learner(kind, ocfg, params, nf4, P, T, B, cuda)
    ln.kl_beta = beta
    assert ln.share_prompts
    grads, loss = ln._compute_gradients(prompts, answers, list(rewards))
    loss_tol = 4e-2 * sum(np.abs(rewards[i:i + B]).mean() for i in range(0, len(rewards), B)) + 2e-3
    assert abs(loss - ref_loss) <= loss_tol, (loss, ref_loss)
    _compare_grads(grads, ref_grads, ocfg, ln.policy)
    # and the unshared (classic) layout of the same learner gives the same gradients up to bf16 reassociation
    ln.share_prompts = False
    grads2, loss2 = ln._compute_gradients(prompts, answers, list(rewards))
    _compare_grads(grads, {f"l{i}.{m}.{ab}": grads2[ln.policy.peft_name(i, m, ab)] for i in range(ocfg.n_layers)
                           for m in lo.LORA_MODULES for ab in ("A", "B")}, ocfg, ln.policy, cos_min=0.9995, rel_max=3e-2)


def test_weight_cache_is_bit_identical(cuda):
    """Resident bf16 copy of the dequantised base (b200rl_model_set_weight_cache) vs per-layer dequant scratch:
    same dequant kernel, same GEMM operands -> identical gradients and log-probs, bit for bit."""
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 12, 36, 4
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    for cache in (False, True):
        pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T, cache_weights=cache)
        assert (pol.weight_cache is not None) == cache
        ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
        for _ in range(2):   # second pass reads the already-filled cache
            ln._compute_gradients(prompts, answers, list(rewards), export=False)
        out.append((pol.lora_grad.clone(), float(pol.loss_accum.item())))
    assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


def test_nf4_in_mainloop_is_bit_identical(cuda):
    """No weight cache: NF4 dequantised inside the GEMM mainloop (Policy(cache_weights="inkernel"), north_star) vs
    dequantised into a scratch before each GEMM — same bf16 operand values, same K order -> identical gradients."""
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 12, 36, 4          # 192 rows per micro-batch > 128: the CTA-pair kernels are used
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    try:
        _capi.lib().b200rl_gemm_set_tail_split(0)      # the two paths may pick different tile widths; keep the K order equal
        for mode in (False, "inkernel"):
            pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T, cache_weights=mode)
            assert pol.weight_cache is None
            ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
            ln._compute_gradients(prompts, answers, list(rewards), export=False)
            out.append((pol.lora_grad.clone(), float(pol.loss_accum.item())))
    finally:
        _capi.lib().b200rl_gemm_set_tail_split(1)
    assert out[0][0].abs().max() > 0
    assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


def test_compact_scored_rows_is_bit_identical(cuda):
    """Head on the loss-carrying completion tokens only (compact_scored_rows, packed layout) vs on all B*T positions like
    the reference (distributed_actor.py:245-260 scores everything, masks afterwards): rows are independent in the final
    norm, lm_head and log-softmax and the dropped rows have coefficient 0 -> identical loss, gradients and log-probs."""
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=128,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 24, 72, 4
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    try:
        _capi.lib().b200rl_gemm_set_tail_split(0)      # the row count changes the tile count; keep the K order equal
        for compact in (True, False):
            pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T)
            ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5,
                                                  "compact_scored_rows": compact})
            assert ln.share_prompts
            pk = ln._pack(*ln._encode(prompts[:B], answers[:B])[:2])
            n_live = int(pk.answer_mask.sum().item())
            assert pk.n_score == (n_live if compact else B * T) and 0 < n_live < B * T
            lp, mask = ln.compute_current_policy_probs(pol, prompts[:B], answers[:B])
            ln._compute_gradients(prompts, answers, list(rewards), export=False)
            out.append((pol.lora_grad.clone(), float(pol.loss_accum.item()), lp.clone(), mask.clone()))
    finally:
        _capi.lib().b200rl_gemm_set_tail_split(1)
    assert out[0][0].abs().max() > 0
    assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
    live = out[0][3] != 0
    assert torch.equal(out[0][3], out[1][3]) and torch.equal(out[0][2][live], out[1][2][live])
    assert (out[0][2][~live] == 0).all()               # positions left out report 0


def test_swiglu_fusion_is_bit_identical(cuda):
    """b200rl_model_set_fusion: SwiGLU inside the GEMM epilogues vs separate row kernels -> identical gradients."""
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 12, 36, 4          # 192 rows per micro-batch > 128: the CTA-pair kernels are used
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    for fuse in (1, 0):
        pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T)
        _capi.check(_capi.lib().b200rl_model_set_fusion(pol.handle, fuse), "set_fusion")
        ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
        ln._compute_gradients(prompts, answers, list(rewards), export=False)
        out.append((pol.lora_grad.clone(), float(pol.loss_accum.item())))
    assert out[0][0].abs().max() > 0
    assert torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]


def test_grouped_dw_matches_per_group_path(cuda, monkeypatch):
    """One grouped dW launch per layer (default) vs eight separate dW GEMMs: same K-ranges are NOT guaranteed, so the
    comparison is numerical (fp32 slabs summed in a different split): rel-L2 < 1e-5 on the flat gradient."""
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 12, 36, 4
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    try:
        _capi.lib().b200rl_gemm_set_ext(0)   # isolate the dW path: the in-kernel LoRA intermediates exist only next to it
        for env in ("1", "0"):
            monkeypatch.setenv("B200RL_GROUPED_DW", env)   # read at model creation
            pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T)
            ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
            ln._compute_gradients(prompts, answers, list(rewards), export=False)
            out.append(pol.lora_grad.clone())
    finally:
        _capi.lib().b200rl_gemm_set_ext(1)
    assert out[1].abs().max() > 0
    rel = ((out[0] - out[1]).norm() / out[1].norm()).item()
    assert rel < 1e-5, rel


def test_inkernel_lora_intermediates_match_separate_gemms(cuda):
    """LoRA intermediates u = s x A^T / du = s dY B produced by the ext units of the big GEMMs (default) vs the separate
    split-K skinny GEMMs: the same products summed in a different fp32 order and rounded to bf16 once, so the gradients
    agree to bf16 reassociation (measured 2.3e-4 rel-L2 on B200; bound 2e-3)."""
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=2)
    P, T, B = 12, 36, 4          # 192 rows per micro-batch > 128: the CTA-pair kernels (and with them the ext units) run
    prompts, answers, rewards = lo.make_batch(ocfg, 8, P, T, seed=4, ragged=True, group_size=4, learner="grpo")
    out = []
    try:
        for ext in (1, 0):
            _capi.lib().b200rl_gemm_set_ext(ext)
            pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=B, P=P, T=T)
            ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
            ln._compute_gradients(prompts, answers, list(rewards), export=False)
            out.append((pol.lora_grad.clone(), float(pol.loss_accum.item())))
    finally:
        _capi.lib().b200rl_gemm_set_ext(1)
    assert out[0][0].abs().max() > 0 and not torch.equal(out[0][0], out[1][0])   # the two paths really differ
    rel = ((out[0][0] - out[1][0]).norm() / out[1][0].norm()).item()
    assert rel < 2e-3, rel
    assert abs(out[0][1] - out[1][1]) < 1e-9


def test_hf_state_dict_loader_vs_transformers(cuda):
    """Policy.from_hf_state_dict (HF Qwen2 names, NF4 quantisation at load like load_in_4bit) against
    transformers' own Qwen2ForCausalLM run on the NF4-dequantised weights: per-token log-probs of the scoring API
    (reference distributed_actor.py:241-260) agree within bf16 error.  LoRA B = 0 (PEFT init) => adapter is a no-op."""
    transformers = pytest.importorskip("transformers")
    from distrl_llm_b200 import ops
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import LMConfig, Policy
    hfc = transformers.Qwen2Config(vocab_size=1024, hidden_size=256, intermediate_size=512, num_hidden_layers=2,
                                   num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=256,
                                   tie_word_embeddings=False, attn_implementation="eager")
    torch.manual_seed(0)
    hf = transformers.Qwen2ForCausalLM(hfc).to(cuda).float().eval()
    with torch.no_grad():   # make the biases / norms non-trivial
        for n, p in hf.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.02)
            if "layernorm" in n or n == "model.norm.weight":
                p.uniform_(0.8, 1.2)
    cfg = LMConfig.from_hf_config(hfc, lora_r=16, lora_alpha=16)
    assert (cfg.head_dim, cfg.n_kv_heads, cfg.inter) == (64, 2, 512) and cfg.rope_theta == 10000.0
    P, T, B = 10, 30, 4
    pol = Policy.from_hf_state_dict(cfg, hf.state_dict(), cuda, max_batch=B, P=P, T=T)
    # give HF exactly the weights the learner holds: NF4 round trip of every linear, bf16 round trip of the rest
    with torch.no_grad():
        for i, L in enumerate(pol.layers):
            lay = hf.model.layers[i]
            qkv = ops.nf4_dequant(L["qkv_p"], L["qkv_a"], cfg.qd + 2 * cfg.kd, cfg.hidden).float()
            lay.self_attn.q_proj.weight.copy_(qkv[:cfg.qd]); lay.self_attn.k_proj.weight.copy_(qkv[cfg.qd:cfg.qd + cfg.kd])
            lay.self_attn.v_proj.weight.copy_(qkv[cfg.qd + cfg.kd:])
            lay.self_attn.o_proj.weight.copy_(ops.nf4_dequant(L["o_p"], L["o_a"], cfg.hidden, cfg.qd).float())
            gu = ops.nf4_dequant(L["gu_p"], L["gu_a"], 2 * cfg.inter, cfg.hidden).float()
            lay.mlp.gate_proj.weight.copy_(gu[:cfg.inter]); lay.mlp.up_proj.weight.copy_(gu[cfg.inter:])
            lay.mlp.down_proj.weight.copy_(ops.nf4_dequant(L["down_p"], L["down_a"], cfg.hidden, cfg.inter).float())
        for p in hf.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5})
    rng = np.random.default_rng(0)
    prompts = [rng.integers(1, 1024, size=int(rng.integers(4, P + 1))).tolist() for _ in range(B)]
    answers = [rng.integers(1, 1024, size=int(rng.integers(8, T + 1))).tolist() for _ in range(B)]
    lp, mask = ln.compute_current_policy_probs(pol, prompts, answers)
    ids, am, _ = lo.pad_batch(prompts, answers, P, T)
    ids, am = ids.to(cuda), am.to(cuda)
    with torch.no_grad():   # reference :241-260 verbatim in spirit: logits -> shift -> log_softmax -> gather
        logits = hf(input_ids=ids, attention_mask=am).logits[:, P - 1:-1, :]
        ref = torch.log_softmax(logits.float(), -1).gather(-1, ids[:, P:].unsqueeze(-1)).squeeze(-1)
    m = mask.bool()
    assert (lp[m] - ref[m]).abs().max().item() < 4e-2
    assert (lp[m] - ref[m]).abs().mean().item() < 6e-3


@pytest.mark.parametrize("kind,beta,hd", [("grpo", 0.0, 64), ("pg", 0.0, 128), ("grpo", 0.3, 128)])
def test_fused_passes_match_per_microbatch_oracle(cuda, kind, beta, hd):
    """Pass fusion (learner.fuse_microbatches): k reference micro-batches per model pass with advantages and the KL
    weight scaled by k.  The oracle loops over micro-batches of train_batch_size like the reference; includes a skipped
    micro-batch (quirk Q1), a ragged last micro-batch (never fused) and prompts shared inside groups."""
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer, Learner
    from distrl_llm_b200.policy import Policy
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=4 if hd == 64 else 2, n_kv_heads=2 if hd == 64 else 1,
                           head_dim=hd, lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=8, lora_b_std=0.05 if beta else 0.01)
    P, T, B, N = 16, 40, 4, 22            # 6 micro-batches: 5 full + one of 2
    rng = np.random.default_rng(5)
    prompts, answers = [], []
    for g in range((N + 3) // 4):
        pr = rng.integers(1, ocfg.vocab, size=int(rng.integers(P // 2, P + 1))).tolist()
        for _ in range(4):
            prompts.append(pr)
            answers.append(rng.integers(1, ocfg.vocab, size=int(rng.integers(T // 4, T + 1))).tolist())
    prompts, answers = prompts[:N], answers[:N]
    rewards = rng.normal(size=N)
    rewards[5] = 0.0                        # micro-batch 1 is skipped by the reference's predicate
    dparams = {k: (v.detach().to(cuda).requires_grad_(v.requires_grad)) for k, v in params.items()}
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    kw = {"kl_beta": beta} if beta else {}
    ref_grads, ref_loss = lo.compute_gradients(dparams, ocfg, ids.to(cuda), am.to(cuda), ansm.to(cuda), rewards, P, B, kind, **kw)
    pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=3 * B, P=P, T=T)
    cls = Learner if kind == "pg" else GRPOLearner
    ln = cls(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5, "kl_beta": beta})
    assert ln.fuse_microbatches == 3
    grads, loss = ln._compute_gradients(prompts, answers, list(rewards))
    tol = 4e-2 * sum(np.abs(rewards[i:i + B]).mean() for i in range(0, N, B)) + 2e-3 + (5e-2 * abs(ref_loss) if beta else 0)
    assert abs(loss - ref_loss) <= tol, (loss, ref_loss)
    _compare_grads(grads, ref_grads, ocfg, pol)
    # one pass per micro-batch on the same policy: same gradient up to bf16 reassociation
    ln.fuse_microbatches = 1
    grads1, loss1 = ln._compute_gradients(prompts, answers, list(rewards))
    assert abs(loss1 - loss) <= 1e-2 * max(1.0, abs(loss))
    _compare_grads(grads, {f"l{i}.{m}.{ab}": grads1[pol.peft_name(i, m, ab)] for i in range(ocfg.n_layers)
                           for m in lo.LORA_MODULES for ab in ("A", "B")}, ocfg, pol, cos_min=0.9995, rel_max=3e-2)


def test_adapter_save_and_resume(cuda, tmp_path):
    """save_checkpoint -> train step -> load_checkpoint restores the adapter (flat fp32 master AND the bf16 operand
    copies the GEMMs read): scoring returns to the saved values bit for bit."""
    ocfg = lo.OracleConfig(vocab=1024, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64,
                           lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=5)
    P, T, B = 8, 24, 4
    prompts, answers, rewards = lo.make_batch(ocfg, 4, P, T, seed=9, ragged=True, group_size=4, learner="grpo")
    ln = _mk_learner("grpo", ocfg, params, nf4, P, T, B, cuda, lr=1e-2)
    lp0, mask = ln.compute_current_policy_probs(ln.policy, prompts, answers)
    flat0 = ln.policy.lora_flat.clone()
    ln.save_checkpoint(str(tmp_path))
    ln.train([{"answers": [answers], "problem": [prompts], "rewards": [rewards]}])
    lp1, _ = ln.compute_current_policy_probs(ln.policy, prompts, answers)
    assert (lp1 - lp0)[mask.bool()].abs().max() > 1e-4
    ln.load_checkpoint(str(tmp_path))
    assert torch.equal(ln.policy.lora_flat, flat0)
    lp2, _ = ln.compute_current_policy_probs(ln.policy, prompts, answers)
    assert torch.equal(lp2, lp0)


# ---------------------------------------------------------------------------------------------------
# clipped-ratio surrogate, inner epochs, optimizer state in checkpoints.  The reference has none of
# these (ratio identically 1, one step per batch, no resume): the oracle restates THIS repo's definition — parity unpinned.
# ---------------------------------------------------------------------------------------------------
def _n4_setup(cuda, hd=128, lr=5e-3, **cfgkw):
    ocfg = lo.OracleConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=2 if hd == 128 else 4,
                           n_kv_heads=1 if hd == 128 else 2, head_dim=hd, lora_r=16, lora_alpha=16)
    params, nf4 = lo.make_params(ocfg, seed=13, lora_b_std=0.05)
    P, T, B, N = 12, 36, 4, 8
    prompts, answers, rewards = lo.make_batch(ocfg, N, P, T, seed=6, ragged=True, group_size=4, learner="grpo")
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import Policy
    pol = Policy.from_params(_mk_cfg(ocfg), params, nf4, cuda, max_batch=2 * B, P=P, T=T)
    ln = GRPOLearner(pol, IdTokenizer(), dict({"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": lr}, **cfgkw))
    return ocfg, params, ln, prompts, answers, rewards, (P, T, B, N)


@pytest.mark.parametrize("hd", [64, 128])
def test_clipped_ratio_vs_oracle(cuda, hd):
    eps = 0.1
    ocfg, params, ln, prompts, answers, rewards, (P, T, B, N) = _n4_setup(cuda, hd=hd, clip_eps=eps)
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    ids, am, ansm = ids.to(cuda), am.to(cuda), ansm.to(cuda)
    # "old policy" log-probs: the current ones shifted by a per-token perturbation large enough to leave the trust region
    lp_a, mask_a = ln.compute_current_policy_probs(ln.policy, prompts[:B], answers[:B])
    lp_b, mask_b = ln.compute_current_policy_probs(ln.policy, prompts[B:], answers[B:])
    lp0 = torch.cat([lp_a, lp_b]) * torch.cat([mask_a, mask_b]).float()
    g = torch.Generator(device=cuda).manual_seed(0)
    old = (lp0 + 0.25 * torch.randn(lp0.shape, generator=g, device=cuda)) * ansm.float()
    loss = ln.compute_loss(prompts, answers, list(rewards), old_lp=old)
    grads = ln.export_gradients()
    dparams = {k: (v.detach().to(cuda).requires_grad_(v.requires_grad)) for k, v in params.items()}
    ref_grads, ref_loss = lo.compute_gradients(dparams, ocfg, ids, am, ansm, rewards, P, B, "grpo", old_lp=old, clip_eps=eps)
    g_plain, loss_plain = lo.compute_gradients(dparams, ocfg, ids, am, ansm, rewards, P, B, "grpo")
    assert abs(ref_loss - loss_plain) > 1e-3, "the clipping must be active in this test"
    assert abs(loss - ref_loss) <= 5e-2 * abs(ref_loss) + 5e-3, (loss, ref_loss)
    # clipping zeroes the gradient of whole tokens: a token whose ratio sits within bf16 error of the trust-region edge
    # may fall on either side, so the global tolerance is looser than for the smooth losses
    _compare_grads(grads, ref_grads, ocfg, ln.policy, cos_min=0.995, rel_max=1e-1)
    # old_lp given but clip_eps = 0 -> exactly the reference loss
    ln.clip_eps = 0.0
    assert abs(ln.compute_loss(prompts, answers, list(rewards), old_lp=old) - loss_plain) < 1e-9


def test_inner_epochs_equal_manual_steps(cuda):
    """train() with inner_epochs = 2: epoch 0 is the reference's step and records the log-probs, epoch 1 is a clipped step
    against them — bit-identical to the same two steps issued by hand."""
    cand = lambda a, p, r: [{"answers": [a], "problem": [p], "rewards": [r]}]
    _, _, ln, prompts, answers, rewards, (P, T, B, N) = _n4_setup(cuda, clip_eps=0.2, inner_epochs=2)
    ln.train(cand(answers, prompts, rewards))
    assert ln.policy.opt_step == 2 and len(ln.last_epoch_losses) == 2
    _, _, m, _, _, _, _ = _n4_setup(cuda, clip_eps=0.2, inner_epochs=1)
    old = torch.zeros(N, T, device=cuda)
    m.compute_loss(prompts, answers, list(rewards), lp_capture=old)
    m.policy.optimizer_step(m.lr)
    m.compute_loss(prompts, answers, list(rewards), old_lp=old)
    m.policy.optimizer_step(m.lr)
    assert torch.equal(ln.policy.lora_flat, m.policy.lora_flat)
    # and with the defaults (clip 0, 1 epoch) train() is the reference's single step
    _, _, r1, _, _, _, _ = _n4_setup(cuda)
    _, _, r2, _, _, _, _ = _n4_setup(cuda, inner_epochs=1, clip_eps=0.0)
    r1.train(cand(answers, prompts, rewards))
    r2.compute_loss(prompts, answers, list(rewards))
    r2.policy.optimizer_step(r2.lr)
    assert torch.equal(r1.policy.lora_flat, r2.policy.lora_flat) and r1.policy.opt_step == 1


def test_checkpoint_carries_optimizer_state(cuda, tmp_path):
    """save -> step -> load restores adapter, Adam moments and step counter; training on from the checkpoint equals
    uninterrupted training bit for bit.  A PEFT-only directory (no optimizer_state.pt) resets the optimizer."""
    import os
    cand = lambda a, p, r: [{"answers": [a], "problem": [p], "rewards": [r]}]
    _, _, ln, prompts, answers, rewards, _ = _n4_setup(cuda, lr=1e-3)
    ln.train(cand(answers, prompts, rewards))
    ln.train(cand(answers, prompts, rewards))
    ln.save_checkpoint(str(tmp_path))
    m0, v0, p0 = ln.policy.adam_m.clone(), ln.policy.adam_v.clone(), ln.policy.lora_flat.clone()
    ln.train(cand(answers, prompts, rewards))
    after3 = ln.policy.lora_flat.clone()
    ln.load_checkpoint(str(tmp_path))
    assert ln.policy.opt_step == 2 and torch.equal(ln.policy.adam_m, m0) and torch.equal(ln.policy.adam_v, v0)
    assert torch.equal(ln.policy.lora_flat, p0)
    ln.train(cand(answers, prompts, rewards))
    assert torch.equal(ln.policy.lora_flat, after3), "resumed training must continue exactly where it stopped"
    os.remove(os.path.join(str(tmp_path), "optimizer_state.pt"))
    ln.load_checkpoint(str(tmp_path))
    assert ln.policy.opt_step == 0 and float(ln.policy.adam_m.abs().max()) == 0.0 and float(ln.policy.adam_v.abs().max()) == 0.0
    import json
    conf = json.load(open(os.path.join(str(tmp_path), "adapter_config.json")))
    conf["lora_alpha"] = 64.0
    json.dump(conf, open(os.path.join(str(tmp_path), "adapter_config.json"), "w"))
    with pytest.raises(ValueError, match="lora_alpha"):
        ln.load_checkpoint(str(tmp_path))
