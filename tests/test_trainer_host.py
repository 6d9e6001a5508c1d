"""CPU tests of the trainer-side host logic (distrl_llm_b200/trainer.py, train_distributed.py, generator.py) against golden
vectors produced by the REFERENCE's own code (oracle/make_golden_trainer.py -> tests/golden/trainer_chunks.json):
batch chunking, the CLI surface, and the order of the actor calls in one trainer step (with CPU stand-ins for the
learners, so no GPU is needed)."""
import json
import os

import numpy as np
import pytest

from oracle import learner_oracle as lo

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trainer_chunks.json")))


@pytest.mark.parametrize("case", GOLD["chunk_cases"], ids=lambda c: f"{c['batch_size']}_{c['num_actors']}_{c['num_learners']}_{c['learner_chunk_size']}")
def test_calculate_chunk_sizes_matches_reference(case):
    from distrl_llm_b200.trainer import Trainer
    got = Trainer.calculate_chunk_sizes(case["batch_size"], case["num_actors"], case["num_learners"], case["learner_chunk_size"])
    assert got == case["chunks"]


def test_calculate_chunk_sizes_rejects_what_the_reference_rejects():
    from distrl_llm_b200.trainer import Trainer
    for bad in ((0, 1, 1, 1), (4, 1, 0, 1), (4, -1, 1, 1)):
        with pytest.raises(ValueError):
            Trainer.calculate_chunk_sizes(*bad)


def test_split_dict_lists_matches_reference():
    from distrl_llm_b200.trainer import Trainer
    for case in GOLD["split_cases"]:
        if case["error"]:
            with pytest.raises(ValueError):
                Trainer.split_dict_lists(GOLD["split_data"], case["sizes"])
        else:
            assert Trainer.split_dict_lists(GOLD["split_data"], case["sizes"]) == case["out"]


def test_cli_has_every_reference_flag_with_the_same_default():
    from distrl_llm_b200.train_distributed import build_parser, config_from_args
    ap = build_parser()
    actions = {a.option_strings[0]: a for a in ap._actions if a.option_strings}
    types = {"str": str, "int": int, "float": float}
    for f in GOLD["cli_flags"]:
        a = actions[f["flag"]]
        assert a.default == f["default"], f["flag"]
        assert a.type is types[f["type"]], f["flag"]
        if f["choices"]:
            assert list(a.choices) == f["choices"]
    cfg = config_from_args(ap.parse_args([]))
    # the keys the reference's config dict carries (train_distributed.py:51-81)
    for k in ("run_name", "project_name", "lora_save_path", "lr", "max_prompt_tokens", "max_new_tokens", "episodes",
              "num_candidates", "batch_size", "train_batch_size", "temperature", "save_every", "eval_every", "model", "dataset",
              "number_of_actors", "number_of_learners", "learner", "use_vllm", "max_lora_rank", "topk", "learner_chunk_size",
              "actor_gpu_usage", "learner_gpu_usage", "lora_alpha", "lora_dropout"):
        assert k in cfg
    assert cfg["use_vllm"] is True


def test_stub_generator_payload_has_the_reference_shape():
    from distrl_llm_b200.generator import StubGenerator, synthetic_reward_function
    g = StubGenerator(vocab=1000, num_candidates=4, max_new_tokens=20, seed=1)
    task = {"problem": [[1, 2, 3], [4, 5]], "solution": [7, 8], "extra": ["a", "b"]}
    out = g.generate(task)
    # distributed_actor.py:165-172: answers / token_lengths [n_prob][n_cand]; solution and problem repeated per candidate
    assert len(out["answers"]) == 2 and all(len(a) == 4 for a in out["answers"])
    assert out["token_lengths"] == [[len(x) for x in a] for a in out["answers"]]
    assert out["problem"][0] == [[1, 2, 3]] * 4 and out["solution"][1] == [8] * 4 and out["extra"] == ["a", "b"]
    assert task["problem"] == [[1, 2, 3], [4, 5]], "the caller's chunk is not modified"
    r = synthetic_reward_function(out["answers"][0], out["solution"][0])
    assert r.shape == (4, 2) and set(np.unique(r[:, 1])) <= {0.0, 1.0} and set(np.unique(r[:, 0])) <= {0.0, 0.1, 0.2}
    n8 = g.generate(task, type("SP", (), {"n": 8})())
    assert all(len(a) == 8 for a in n8["answers"])


class _FakeLearner:
    """CPU stand-in with the learner's actor surface; records the order of the calls."""

    def __init__(self, log, name, p2p=False):
        self.log, self.name, self.p2p = log, name, p2p

    def generate(self, task, sampling_params=None):
        self.log.append((self.name, "generate", len(task["problem"])))
        return dict(task, answers=[], token_lengths=[])

    def train(self, candidates):
        self.log.append((self.name, "train", sum(len(a) for c in candidates for a in c["answers"])))
        return 0.5

    def compute_gradients(self, chunk):
        self.log.append((self.name, "compute_gradients", len(chunk[0])))
        return ({} if self.p2p else {"w": np.ones(2)}), 1.0 + len(chunk[0])

    def apply_merged_gradients(self, grads):
        self.log.append((self.name, "apply_merged_gradients", None if grads is None else len(grads)))

    def save_adapter(self):
        self.log.append((self.name, "save_adapter", None))

    def save_checkpoint(self, path):
        self.log.append((self.name, "save_checkpoint", path))

    def save_optimizer_state(self, path):
        self.log.append((self.name, "save_optimizer_state", path))


def _oracle_prep(candidates, learner_type, topk, device):
    """trainer_prep.apply_advantages_and_topk restated with the oracle (the product runs this block in the G9 CUDA kernel)."""
    for cand in candidates:
        vals, bases = zip(*[lo.group_advantages(np.asarray(r), learner_type) for r in cand["rewards"]])
        idx = [lo.topk_filter(v, topk) for v in vals]
        if learner_type != "grpo":
            cand["baselines"] = list(bases)
        cand["answers"] = [[cand["answers"][j][i] for i in idx[j]] for j in range(len(idx))]
        cand["rewards"] = [vals[j][idx[j]] for j in range(len(idx))]
        cand["problem"] = [cand["problem"][j][:topk] for j in range(len(idx))]
    return candidates


@pytest.mark.parametrize("n_learners,p2p,overlap", [(1, False, False), (2, False, False), (2, True, False), (2, True, True)])
def test_trainer_step_calls_the_actors_in_the_reference_order(monkeypatch, n_learners, p2p, overlap):
    from distrl_llm_b200 import local_rpc, trainer as tr, trainer_prep
    from distrl_llm_b200.generator import StubGenerator, synthetic_reward_function
    monkeypatch.setattr(trainer_prep, "apply_advantages_and_topk", _oracle_prep)
    log = []
    actors = [local_rpc.ActorHandle(lambda i=i: StubGenerator(500, 4, 12, seed=i)) for i in range(2)]
    learners = [local_rpc.ActorHandle(lambda i=i: _FakeLearner(log, f"L{i}", p2p)) for i in range(n_learners)]
    config = dict(episodes=1, batch_size=6, learner_chunk_size=0, num_candidates=4, save_every=100, eval_every=0, topk=3,
                  learner="grpo", run_name="t", number_of_actors=2, number_of_learners=n_learners, max_steps=2,
                  overlap_generation=overlap)
    ds = tr.SyntheticDataset(12, 500, 10, seed=0)
    t = tr.Trainer(ds, tr.SyntheticDataset(2, 500, 10, seed=1), synthetic_reward_function, config, actors=actors, learners=learners)
    steps, _ = t.train()
    assert steps == 2 and len(t.history) == 2
    per_step = [e for e in log if e[1] != "generate"]
    if n_learners == 1:
        # 6 problems x top-3 of 4 candidates = 18 sequences to learners[0].train, then save_adapter (:306-307, :346)
        assert per_step == [("L0", "train", 18), ("L0", "save_adapter", None)] * 2
    else:
        one = per_step[:len(per_step) // 2]
        grads = [e for e in one if e[1] == "compute_gradients"]
        assert sorted(grads) == [("L0", "compute_gradients", 9), ("L1", "compute_gradients", 9)]   # even split (:312-322)
        applies = [e for e in one if e[1] == "apply_merged_gradients"]
        if p2p:   # fused exchange: every learner takes part
            assert sorted(applies) == [("L0", "apply_merged_gradients", None), ("L1", "apply_merged_gradients", None)]
        else:     # reference exchange: the list of gradient dicts goes to learner 0 only (:342)
            assert applies == [("L0", "apply_merged_gradients", 2)]
        assert one[-1] == ("L0", "save_adapter", None)
        assert max(i for i, e in enumerate(one) if e[1] == "compute_gradients") < min(i for i, e in enumerate(one) if e[1] == "apply_merged_gradients")
    # learner_chunk_size = 0: the learners are still asked to generate (an empty chunk), like the reference's Trainer
    assert sorted([e for e in log if e[1] == "generate"][:n_learners]) == [(f"L{i}", "generate", 0) for i in range(n_learners)]
    m = t.history[0]
    for k in ("loss", "mean_format_reward", "mean_accuracy_reward", "min_accuracy_reward", "max_accuracy_reward", "mean_token_length",
              "episode", "total_batch_steps", "total_samples_processed", "timing/update_duration", "timing/reward_duration",
              "timing/generation_duration"):
        assert k in m, k     # the reference's wandb keys (:348-366)
    for h in actors + learners:
        h.shutdown()


def test_bench_reference_arm_bookkeeping_on_a_tiny_model():
    """bench.py --impl reference: one bounded sample per step, measured ms_per_step, extrapolated value (tiny debug shape)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B200RL_CPU_ARM_SHAPE="1024,128,256,2,1", B200RL_CPU_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "1",
                          "--prompt_len", "8", "--new_tokens", "16"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["ms_per_step"] < line["extrapolated_ms_per_full_step"]
    assert line["cpu_baseline"]["cores"] == 2 and line["steps"] == 4
