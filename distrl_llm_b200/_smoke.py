"""smoke(): one tiny GRPO learner step on cuda:0 through the C ABI, checked against the oracle."""
from __future__ import annotations

import torch


def run():
    if not torch.cuda.is_available():
        raise RuntimeError("smoke() needs a CUDA device (B200)")
    from oracle import learner_oracle as lo  # checker only
    from .learner import GRPOLearner, IdTokenizer
    from .policy import LMConfig, Policy

    dev = torch.device("cuda:0")
    ocfg = lo.OracleConfig(vocab=512, hidden=128, inter=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=32,
                           lora_r=16, lora_alpha=16)
    P, T, B, N = 16, 32, 2, 4
    params, nf4 = lo.make_params(ocfg, seed=3)
    prompts, answers, adv = lo.make_batch(ocfg, N, P, T, seed=4, ragged=True, group_size=N, learner="grpo")
    cfg = LMConfig(vocab=512, hidden=128, inter=256, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=32)
    pol = Policy.from_params(cfg, params, nf4, dev, max_batch=B, P=P, T=T)
    ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 2e-5})
    grads, loss = ln._compute_gradients(prompts, answers, list(adv))
    ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
    ref, ref_loss = lo.compute_gradients(params, ocfg, ids, am, ansm, adv, P, B, "grpo")
    va = torch.cat([grads[pol.peft_name(i, m, ab)].double().flatten() for i in range(2) for m in lo.LORA_MODULES for ab in "AB"])
    vb = torch.cat([ref[f"l{i}.{m}.{ab}"].double().flatten() for i in range(2) for m in lo.LORA_MODULES for ab in "AB"])
    cos = float((va @ vb) / (va.norm() * vb.norm()))
    assert abs(loss - ref_loss) < 1e-6, (loss, ref_loss)
    assert cos > 0.999, cos
    ln.policy.optimizer_step(ln.lr)
    torch.cuda.synchronize()
    print(f"smoke ok: GRPO step on {torch.cuda.get_device_name(0)}, loss {loss:.6f} (oracle {ref_loss:.6f}), "
          f"LoRA-grad cosine vs oracle {cos:.6f}")
