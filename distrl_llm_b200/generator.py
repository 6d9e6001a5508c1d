"""Stub generator: stands in for the reference's vLLM `Generator` actor (distributed_actor.py:20-193) where no checkpoint
or vLLM engine is available (north_star: "actors still generate completions (vLLM or a stub)").  It returns the SAME
payload `BaseActor.vllm_generate` builds (:147-172) — answers / token_lengths per problem, solution and problem repeated
per candidate — with random token ids, and it consumes the learner's adapter through the in-memory hand-off
(adapter_sync.py) exactly where the reference calls `load_lora(self.policy, self.lora_save_path)` (:150).

Generation itself is out of scope; what this class exercises is everything around it: chunk payloads,
ragged completion lengths, the token-id passthrough to the learner (the ids the generator sampled are
what the learner scores — no retokenisation, quirk Q6), and the adapter version every batch was generated with."""
from __future__ import annotations

import time

import numpy as np


class StubGenerator:
    def __init__(self, vocab, num_candidates, max_new_tokens, seed=0, min_len_frac=0.25, tokens_per_second=None,
                 adapter_subscriber=None):
        self.vocab, self.num_candidates, self.max_new_tokens = vocab, num_candidates, max_new_tokens
        self.rng = np.random.default_rng(seed)
        self.min_len = max(1, int(max_new_tokens * min_len_frac))
        self.tokens_per_second = tokens_per_second     # None = as fast as possible
        self.adapter = adapter_subscriber
        self.adapter_version = 0
        self.generated_tokens = 0

    def attach_adapter(self, subscriber):
        self.adapter = subscriber

    def stats(self):
        return {"adapter_version": self.adapter_version, "generated_tokens": self.generated_tokens,
                "adapter_checksum": None if self.adapter is None else float(self.adapter.flat.double().sum().item())}

    def generate(self, task, sampling_params=None):
        """task: {"problem": [prompt, ...], "solution": [...], <other dataset columns>} (a chunk made by
        Trainer.split_dict_lists).  Returns the task dict extended like distributed_actor.py:165-172."""
        if self.adapter is not None:      # reference: load_lora(...) from disk on every generate (:150)
            self.adapter_version = self.adapter.pull()
        n = getattr(sampling_params, "n", None) or self.num_candidates
        t0 = time.perf_counter()
        answers, lengths = [], []
        for _ in task["problem"]:
            lens = self.rng.integers(self.min_len, self.max_new_tokens + 1, size=n)
            answers.append([self.rng.integers(1, self.vocab, size=int(k)).tolist() for k in lens])
            lengths.append([int(k) for k in lens])
            self.generated_tokens += int(lens.sum())
        if self.tokens_per_second:
            budget = sum(map(sum, lengths)) / self.tokens_per_second - (time.perf_counter() - t0)
            if budget > 0:
                time.sleep(budget)
        task = dict(task)
        task["answers"] = answers
        task["token_lengths"] = lengths
        task["solution"] = [[s for _ in range(n)] for s in task["solution"]]
        task["problem"] = [[p for _ in range(n)] for p in task["problem"]]
        task["adapter_version"] = self.adapter_version
        return task


def synthetic_reward_function(completions, solutions):
    """Synthetic stand-in for reward_functions.reward_function (reference reward_functions.py:44-49) on token-id
    completions: column 0 = "format" in {0, .1, .2}, column 1 = "accuracy" in {0, 1}, both deterministic functions of the
    token ids (so repeated runs agree), with the marginals of SURVEY.md 8d."""
    out = np.zeros((len(completions), 2), dtype=np.float64)
    for i, c in enumerate(completions):
        h = (sum(c[:8]) * 2654435761 + len(c) * 40503) & 0xFFFFFFFF
        u, v = (h & 0xFFFF) / 65536.0, (h >> 16) / 65536.0
        out[i, 0] = 0.0 if u < 0.5 else (0.1 if u < 0.8 else 0.2)
        out[i, 1] = 1.0 if v < 0.25 else 0.0
    return out
