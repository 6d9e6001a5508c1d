"""GPU tests of the pieces around the learner that BASELINE config 5 (full pipeline) exercises: the in-memory adapter
hand-off (reference: save_lora to disk :84-86 + load_lora from disk :150), and the trainer loop
(distributed_trainer.py:232-382) driving real learners and stub generators through train_distributed's CLI surface."""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny_policy(cuda, seed=3):
    from distrl_llm_b200.policy import LMConfig, Policy
    cfg = LMConfig(vocab=1024, hidden=256, inter=512, n_layers=2, n_q_heads=2, n_kv_heads=1, head_dim=128)
    return cfg, Policy.random_init(cfg, cuda, 4, 8, 24, seed=seed)


def test_adapter_publish_pull(cuda):
    from distrl_llm_b200.adapter_sync import AdapterPublisher, AdapterSubscriber
    from distrl_llm_b200.policy import LoraLayout
    cfg, pol = _tiny_policy(cuda)
    pub = AdapterPublisher(pol)
    sub = AdapterSubscriber(pub.describe(), cuda, same_process=True)
    assert sub.pull() == 0 and float(sub.flat.abs().max()) == 0.0          # nothing published yet
    v1 = pub.publish()
    assert sub.pull() == v1 == 2 and torch.equal(sub.flat, pol.lora_flat)
    pol.lora_flat.mul_(1.5)                                                 # an optimizer step changes the adapter ...
    assert sub.pull() == v1 and not torch.equal(sub.flat, pol.lora_flat)    # ... invisible until published
    v2 = pub.publish()
    assert sub.pull() == v2 == 4 and torch.equal(sub.flat, pol.lora_flat)
    # the pulled buffer under PEFT names == the learner's own named views (what vLLM's in-memory LoRA loading consumes)
    layout = LoraLayout(cfg)
    assert layout.numel == pol.lora_numel
    mine, theirs = pol.named_views(pol.lora_flat), sub.as_peft_tensors(layout)
    assert mine.keys() == theirs.keys() and len(mine) == 2 * 7 * cfg.n_layers
    for k in mine:
        assert torch.equal(mine[k], theirs[k])
    sub.close()
    pub.close()


def _run_cli(argv, capsys):
    from distrl_llm_b200 import train_distributed
    trainer = train_distributed.main(argv)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    return trainer, line


BASE = ["--model", "random:tiny", "--learner", "grpo", "--number_of_actors", "2", "--batch_size", "4", "--learner_chunk_size", "0",
        "--num_candidates", "4", "--topk", "3", "--max_new_tokens", "24", "--max_prompt_tokens", "12", "--train_batch_size", "4",
        "--max_lora_rank", "16", "--episodes", "1", "--eval_every", "0", "--lr", "1e-3", "--bench"]


def test_trainer_single_learner_end_to_end(cuda, capsys):
    from distrl_llm_b200 import local_rpc
    trainer, line = _run_cli(BASE + ["--number_of_learners", "1", "--max_steps", "3"], capsys)
    assert line["steps"] == 3 and line["value"] > 0 and line["generators"] == "stub"
    assert len(trainer.history) == 3 and all(abs(m["loss"]) < 10 for m in trainer.history)
    # the generators pulled the adapter the learner published after each step: 1 initial publish + 3 steps = version 8,
    # the last generate() ran before the last publish
    stats = local_rpc.get([a.stats.remote() for a in trainer.actors])
    assert all(s["adapter_version"] == 6 for s in stats), stats
    for h in trainer.actors + trainer.learners:
        h.shutdown()


@pytest.mark.parametrize("overlap", [False, True])
def test_trainer_two_learners_on_one_gpu(cuda, capsys, overlap):
    """Two learners as two threads on ONE device (own streams), fused P2P reduce + Adam between them through plain
    pointers: after every step both hold identical adapters (fixes quirk Q4)."""
    from distrl_llm_b200 import local_rpc
    from distrl_llm_b200 import train_distributed
    argv = BASE + ["--number_of_learners", "2", "--max_steps", "2"] + (["--overlap_generation"] if overlap else [])
    args = train_distributed.build_parser().parse_args(argv)
    config = train_distributed.config_from_args(args)
    config["learner_gpus"] = [0, 0]
    config["stub_generators_share_gpus"] = True
    from distrl_llm_b200.actors import create_actor_and_learner
    from distrl_llm_b200.generator import synthetic_reward_function
    from distrl_llm_b200.trainer import SyntheticDataset, Trainer
    actors, learners = create_actor_and_learner(2, 2, args.model, None, config)
    t = Trainer(SyntheticDataset(8, 2048, 12, seed=0), SyntheticDataset(2, 2048, 12, seed=1), synthetic_reward_function, config,
                actors=actors, learners=learners)
    steps, _ = t.train()
    assert steps == 2
    flats = local_rpc.get([l.export_flat.remote() for l in learners])
    assert torch.equal(flats[0], flats[1]) and float(flats[0].abs().sum()) > 0
    for h in actors + learners:
        h.shutdown()
