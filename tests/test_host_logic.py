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


# ---------------------------------------------------------------------------------------------------
# packed shared-prompt / ragged layout bookkeeping (distrl_llm_b200/packing.py) — pure integer logic
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("compact", [True, False])
@pytest.mark.parametrize("ragged", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_packing_descriptors_are_consistent(ragged, seed, compact):
    from distrl_llm_b200 import packing
    rng = np.random.default_rng(seed)
    P, T, groups, per = 150, 300, 3, 4
    B = groups * per
    ids = np.zeros((B, P + T), np.int32)
    am = np.zeros_like(ids)
    for g in range(groups):
        plen = int(rng.integers(1, P + 1))
        pr = rng.integers(1, 5000, size=plen)
        for j in range(per):
            i = g * per + j
            ids[i, P - plen:P], am[i, P - plen:P] = pr, 1
            n = int(rng.integers(0, T + 1)) if (i % 5) else 0          # some empty completions
            ids[i, P:P + n], am[i, P:P + n] = rng.integers(1, 5000, size=n), 1
    h = packing.pack_microbatch(ids, am, P, T, ragged=ragged, compact_scored=compact)
    a = h.arrays
    assert h.n_groups == groups
    # scored rows: every (sequence, t) position, or only those that carry loss (score_slot = their i*T+t, ascending)
    slot = a["score_slot"] if a["score_slot"].size else np.arange(B * T)
    assert compact or a["score_slot"].size == 0
    assert a["score_src"].size == slot.size and a["targets"].size == slot.size
    if compact and a["score_slot"].size:
        assert np.array_equal(slot, np.flatnonzero(a["answer_mask"])) or (a["answer_mask"].sum() == 0 and slot.tolist() == [0])
    row_of_slot = {int(sl): k for k, sl in enumerate(slot)}
    real = int(am[:, P:].sum() + sum(am[g * per, :P].sum() for g in range(groups)))
    assert h.rows == (max(real, 1) if ragged else groups * P + B * T)
    assert int(a["key_mask"].sum()) == real
    # every real token is stored once, with its ORIGINAL position (index in the padded row, reference :217-239)
    for i in range(B):
        g = int(h.seq_group[i])
        st, n = h.prompt_ext[g]
        r0 = int(h.prompt_row0[g])
        assert np.array_equal(a["ids"][r0:r0 + n], ids[i, st:st + n]) and np.array_equal(a["pos"][r0:r0 + n], st + np.arange(n))
        c0, cn = int(h.comp_row0[i]), h.comp_len[i]
        assert np.array_equal(a["ids"][c0:c0 + cn], ids[i, P:P + cn]) and np.array_equal(a["pos"][c0:c0 + cn], P + np.arange(cn))
        # scored positions: the logit at padded position P-1+t predicts completion token t (reference :245-249)
        for t in range(T):
            if a["answer_mask"][i * T + t]:
                k = row_of_slot[i * T + t]
                src = int(a["score_src"][k])
                assert a["pos"][src] == P - 1 + t
                assert a["targets"][k] == ids[i, P + t]
                assert (r0 <= src < r0 + n) if t == 0 else (c0 <= src < c0 + cn)
    assert np.array_equal(a["answer_mask"].reshape(B, T) != 0, (am[:, P:] != 0) & (am[:, :P].sum(1, keepdims=True) > 0))
    # scatter CSR = inverse of score_src over the live positions
    live = []
    for r in range(h.rows):
        for k in a["sc_list"][a["sc_start"][r]:a["sc_start"][r + 1]]:
            assert a["score_src"][k] == r
            live.append(int(k))
    assert len(live) == len(set(live)) and set(np.flatnonzero(a["answer_mask"])) <= set(int(slot[k]) for k in live)
    # query blocks tile every stored row exactly once; key blocks' partial slabs are each reduced into exactly one row
    qb = a["qblocks"].reshape(-1, packing.QB_FIELDS)
    cover = np.zeros(h.rows + 128, np.int32)
    for q_row0, q_rows, q_local0, own_row0, own_len, pre_row0, pre_len, stat0 in qb:
        cover[q_row0:q_row0 + q_rows] += 1
        assert q_row0 == own_row0 + q_local0 and q_local0 + q_rows <= own_len and stat0 == q_row0
    stored = real if ragged else h.rows
    assert (cover[:stored] == 1).all() and cover[stored:].sum() == 0
    kbk = a["kblocks"].reshape(-1, packing.KB_FIELDS)
    assert sum(int(k[1]) for k in kbk) == (h.part_rows if stored else 0)
    used = np.zeros(h.part_rows, np.int32)
    for r in range(h.rows):
        for prow in a["red_list"][a["red_start"][r]:a["red_start"][r + 1]]:
            used[prow] += 1
    assert (used[:sum(int(k[1]) for k in kbk)] == 1).all()
    for k_row0, k_rows, k_local0, q_row0, q_len, causal, stat0, out_row0 in kbk:
        assert q_len > 0 and k_rows > 0
        for rr in (0, k_rows - 1):   # the slab rows of this block are reduced into the block's own key rows
            assert out_row0 + rr in a["red_list"][a["red_start"][k_row0 + rr]:a["red_start"][k_row0 + rr + 1]]


# ---------------------------------------------------------------------------------------------------
# pass planning of compute_loss (learner.py): which sequences go into which model pass, with which scaling
# ---------------------------------------------------------------------------------------------------
class _RecordingPolicy:
    """Stands in for Policy on a machine without a GPU: records every microbatch call of compute_loss."""

    class _Cfg:
        head_dim = 64

    def __init__(self, max_batch):
        self.cfg, self.max_batch, self.device = self._Cfg(), max_batch, torch.device("cpu")
        self.loss_accum = torch.zeros(1, dtype=torch.float64)
        self.calls = []

    def zero_grad(self):
        self.calls.append(("zero_grad",))

    def microbatch(self, ids, am, ansm, adv, P, T, nb, grpo, backward, lp_out=None, lora_off=False, ref_lp=None, kl_beta=0.0,
                   old_lp=None, clip_eps=0.0):
        self.calls.append(("mb", ids.clone(), None if adv is None else adv.clone(), nb, grpo, backward, lora_off, kl_beta))


@pytest.mark.parametrize("k", [1, 2, 3])
def test_compute_loss_pass_planning(k):
    """reference loop: distributed_actor.py:354-389 / :452-487 (micro-batches of train_batch_size, skip predicate :367,
    1/nb scaling with skipped batches still counted).  Fused passes carry k full micro-batches with advantages x k."""
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    B, P, T, n = 4, 6, 8, 22                       # 6 micro-batches: 5 full + one of 2
    rng = np.random.default_rng(0)
    prompts = [rng.integers(1, 50, size=int(rng.integers(1, P + 1))).tolist() for _ in range(n)]
    answers = [rng.integers(1, 50, size=int(rng.integers(1, T + 1))).tolist() for _ in range(n)]
    rewards = rng.normal(size=n)
    rewards[9] = 0.0                                # micro-batch 2 (sequences 8..11) is skipped (quirk Q1)
    pol = _RecordingPolicy(max_batch=k * B)
    ln = GRPOLearner(pol, IdTokenizer(), {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-5,
                                           "share_prompts": False, "kl_beta": 0.5})
    ln._h2d = lambda t: t.contiguous()              # no pinned memory without CUDA
    assert ln.fuse_microbatches == k
    ln.compute_loss(prompts, answers, list(rewards))
    assert pol.calls[0] == ("zero_grad",)
    ref_ids, _, _ = lo.pad_batch(prompts, answers, P, T)
    seen, nb = [], 6
    train = [c for c in pol.calls[1:] if c[5]]      # backward passes
    score = [c for c in pol.calls[1:] if not c[5]]  # KL reference passes (adapter off), one per training pass
    assert len(train) == len(score) and all(c[6] for c in score) and not any(c[6] for c in train)
    for c in train:
        _, ids, adv, nb_c, grpo, backward, lora_off, beta = c
        rows = ids.shape[0]
        kk = rows // B if rows % B == 0 else 1
        assert nb_c == nb and grpo and kk <= k
        # which sequences: match the padded rows back to the batch
        idx = [int(np.flatnonzero((ref_ids.numpy() == ids[r].numpy()).all(1))[0]) for r in range(rows)]
        assert idx == sorted(idx) and not (set(idx) & set(range(8, 12)))
        assert all(len({j // B for j in idx[g * B:(g + 1) * B]}) == 1 for g in range(kk))   # whole micro-batches only
        np.testing.assert_allclose(adv.numpy(), rewards[idx] * kk, rtol=0, atol=0)
        assert beta == 0.5 * kk
        seen += idx
    assert sorted(seen) == [j for j in range(n) if not 8 <= j < 12]
    # the ragged last micro-batch (2 sequences) is never fused
    assert any(c[1].shape[0] == 2 for c in train)


def test_from_pretrained_reads_config_and_safetensors(tmp_path, monkeypatch):
    """Policy.from_pretrained: config.json -> LMConfig, all *.safetensors shards merged, handed to from_hf_state_dict
    (the GPU part — NF4 quantisation + model creation — is covered by test_hf_state_dict_loader_vs_transformers)."""
    import json
    from safetensors.torch import save_file
    from distrl_llm_b200.policy import LMConfig, Policy
    hf = {"vocab_size": 512, "hidden_size": 128, "intermediate_size": 256, "num_hidden_layers": 2,
          "num_attention_heads": 2, "num_key_value_heads": 1, "rms_norm_eps": 1e-5,
          "rope_parameters": {"rope_theta": 123456.0, "rope_type": "default"}}
    (tmp_path / "config.json").write_text(json.dumps(hf))
    save_file({"model.embed_tokens.weight": torch.zeros(512, 128), "model.norm.weight": torch.ones(128)},
              str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({"model.layers.0.self_attn.q_proj.weight": torch.zeros(128, 128)},
              str(tmp_path / "model-00002-of-00002.safetensors"))
    got = {}

    def fake(cls, cfg, sd, device, max_batch, P, T, **kw):
        got.update(cfg=cfg, keys=sorted(sd), device=device, mb=max_batch, P=P, T=T, kw=kw)
        return "policy"

    monkeypatch.setattr(Policy, "from_hf_state_dict", classmethod(fake))
    pol, cfg = Policy.from_pretrained(str(tmp_path), "cuda:0", 4, 16, 32, lora_r=8, lora_alpha=32.0, lora_seed=3)
    assert pol == "policy" and cfg is got["cfg"]
    assert cfg == LMConfig(vocab=512, hidden=128, inter=256, n_layers=2, n_q_heads=2, n_kv_heads=1, head_dim=64,
                           lora_r=8, lora_alpha=32.0, rms_eps=1e-5, rope_theta=123456.0)
    assert got["keys"] == ["model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight", "model.norm.weight"]
    assert (got["mb"], got["P"], got["T"], got["kw"]) == (4, 16, 32, {"lora_seed": 3})
    with pytest.raises(FileNotFoundError):
        (tmp_path / "empty").mkdir()
        (tmp_path / "empty" / "config.json").write_text(json.dumps(hf))
        Policy.from_pretrained(str(tmp_path / "empty"), "cuda:0", 4, 16, 32)


def test_adapter_directory_is_peft_format(tmp_path):
    """save_checkpoint / save_adapter (reference distributed_actor.py:84-86, :263-264): PEFT adapter directory."""
    import json
    from safetensors.torch import load_file
    from distrl_llm_b200.learner import write_peft_adapter
    from distrl_llm_b200.policy import LMConfig, Policy
    cfg = LMConfig(vocab=64, hidden=128, inter=256, n_layers=2, n_q_heads=2, n_kv_heads=1, head_dim=64, lora_r=8, lora_alpha=32.0)
    shapes = cfg.module_shapes()
    sd = {}
    for i in range(cfg.n_layers):
        for m, (fin, fout) in shapes.items():
            sd[Policy.peft_name(i, m, "A")] = torch.randn(cfg.lora_r, fin)
            sd[Policy.peft_name(i, m, "B")] = torch.randn(fout, cfg.lora_r)
    write_peft_adapter(str(tmp_path), sd, cfg, base_model="unsloth/Qwen2.5-7B-Instruct")
    conf = json.load(open(tmp_path / "adapter_config.json"))
    assert (conf["peft_type"], conf["r"], conf["lora_alpha"], conf["bias"]) == ("LORA", 8, 32.0, "none")
    assert sorted(conf["target_modules"]) == sorted(["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])
    got = load_file(str(tmp_path / "adapter_model.safetensors"))
    assert len(got) == 2 * 7 * cfg.n_layers
    k = "base_model.model.model.layers.1.mlp.down_proj.lora_B.weight"
    assert k in got and torch.equal(got[k], sd[Policy.peft_name(1, "down", "B")])
    assert got["base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight"].shape == (8, 128)
    # round trip back to the in-memory names (load_checkpoint)
    from distrl_llm_b200.learner import read_peft_adapter
    back, conf2 = read_peft_adapter(str(tmp_path))
    assert conf2["r"] == 8 and set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
