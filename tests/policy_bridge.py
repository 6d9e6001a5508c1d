"""Test helper: view a device-resident `Policy` (NF4 base + flat LoRA buffer) as the oracle's named-tensor dict, so
that models too large to build through the oracle's CPU/numpy path (Qwen2.5-7B-wide layers, V = 152064) can still be
checked against oracle/learner_oracle.py.  The base matrices are dequantised by the library's own NF4 kernel, which
tests/test_gpu_kernels.py::test_nf4_roundtrip pins bit-for-bit to the oracle's `bf16(level * absmax)` definition."""
import torch

from oracle import learner_oracle as lo


def oracle_config(cfg):
    return lo.OracleConfig(vocab=cfg.vocab, hidden=cfg.hidden, inter=cfg.inter, n_layers=cfg.n_layers,
                           n_q_heads=cfg.n_q_heads, n_kv_heads=cfg.n_kv_heads, head_dim=cfg.head_dim,
                           lora_r=cfg.lora_r, lora_alpha=cfg.lora_alpha, rms_eps=cfg.rms_eps, rope_theta=cfg.rope_theta)


def oracle_params_from_policy(pol, dtype=torch.float32):
    """{oracle name: tensor on pol.device}; LoRA tensors are leaves with requires_grad (fp32 copies of the flat master)."""
    from distrl_llm_b200 import ops
    cfg = pol.cfg
    H, I, qd, kd = cfg.hidden, cfg.inter, cfg.qd, cfg.kd
    p = {"embed": pol.embed.to(dtype), "lm_head": pol.lm_head.to(dtype), "final_norm": pol.final_norm.to(dtype)}
    for i, L in enumerate(pol.layers):
        qkv = ops.nf4_dequant(L["qkv_p"], L["qkv_a"], qd + 2 * kd, H).to(dtype)
        p[f"l{i}.wq"], p[f"l{i}.wk"], p[f"l{i}.wv"] = qkv[:qd], qkv[qd:qd + kd], qkv[qd + kd:]
        p[f"l{i}.wo"] = ops.nf4_dequant(L["o_p"], L["o_a"], H, qd).to(dtype)
        gu = ops.nf4_dequant(L["gu_p"], L["gu_a"], 2 * I, H).to(dtype)
        p[f"l{i}.wg"], p[f"l{i}.wu"] = gu[:I], gu[I:]
        p[f"l{i}.wd"] = ops.nf4_dequant(L["down_p"], L["down_a"], H, I).to(dtype)
        b = L["qkv_bias"].to(dtype)
        p[f"l{i}.bq"], p[f"l{i}.bk"], p[f"l{i}.bv"] = b[:qd], b[qd:qd + kd], b[qd + kd:]
        p[f"l{i}.ln1"], p[f"l{i}.ln2"] = L["ln1"].to(dtype), L["ln2"].to(dtype)
    for (i, m, ab), (off, shp) in pol.offsets.items():
        n = shp[0] * shp[1]
        p[f"l{i}.{m}.{ab}"] = pol.lora_flat[off:off + n].detach().clone().view(*shp).to(dtype).requires_grad_(True)
    return p


def flat_grad_from_oracle(pol, grads):
    """Oracle gradient dict -> flat fp64 vector in the policy's flat-buffer order."""
    out = torch.zeros(pol.lora_numel, dtype=torch.float64, device=pol.device)
    for (i, m, ab), (off, shp) in pol.offsets.items():
        n = shp[0] * shp[1]
        out[off:off + n] = grads[f"l{i}.{m}.{ab}"].detach().double().reshape(-1)
    return out
