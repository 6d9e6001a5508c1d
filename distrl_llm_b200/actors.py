"""create_actor_and_learner: the reference's process topology (distributed_actor.py:517-585) for the B200 learner.

Reference: one Ray actor per generator on GPUs [0, A) and per learner on GPUs [A, A+L) of ONE node (`STRICT_PACK`),
learner class picked by config["learner"] in {"pg", "grpo"} (:571-576).  Same here, with the actor handles coming from
local_rpc (one thread per actor in this process) when Ray is absent; under Ray the same classes are wrapped with
`ray.remote(num_gpus=1)`.  `model` names a local Hugging Face checkpoint directory (config.json +
*.safetensors, quantised to NF4 at load like `load_in_4bit=True`, :58-66) or a synthetic model:
  "random:qwen2.5-7b"            Qwen2.5-7B shape, random-init (BASELINE configs)
  "random:qwen2.5-7b:layers=4"   same width, fewer layers (tests / smoke runs)
  "random:tiny"                  2-layer, hidden 256 (CPU-cheap checker shapes)
Generators are StubGenerator instances unless config["generator_factory"] provides something else (e.g. a vLLM wrapper).
"""
from __future__ import annotations

import torch

from . import local_rpc
from .adapter_sync import AdapterSubscriber
from .generator import StubGenerator
from .learner import GRPOLearner, IdTokenizer, Learner
from .p2p import P2PGroup, lora_numel
from .policy import LMConfig, Policy


def model_config(name, lora_r, lora_alpha):
    if name.startswith("random:"):
        parts = name.split(":")[1:]
        if parts[0] == "tiny":
            cfg = LMConfig(vocab=2048, hidden=256, inter=512, n_layers=2, n_q_heads=2, n_kv_heads=1, head_dim=128,
                           lora_r=lora_r, lora_alpha=lora_alpha)
        elif parts[0] == "qwen2.5-7b":
            cfg = LMConfig.qwen25_7b(lora_r=lora_r, lora_alpha=lora_alpha)
        else:
            raise ValueError(f"unknown synthetic model '{name}'")
        for kv in parts[1:]:
            k, v = kv.split("=")
            if k != "layers":
                raise ValueError(f"unknown model option '{kv}'")
            cfg.n_layers = int(v)
        return cfg
    return None   # a checkpoint directory: the config comes from its config.json


def _build_policy(name, cfg, device, max_batch, P, T, lora_r, lora_alpha, bufs):
    if cfg is not None:
        return Policy.random_init(cfg, device, max_batch, P, T, seed=1234, **bufs)
    pol, _ = Policy.from_pretrained(name, device, max_batch, P, T, lora_r=lora_r, lora_alpha=lora_alpha, **bufs)
    return pol


def create_actor_and_learner(number_of_actors=1, number_of_learners=1, model_name="random:tiny", gen_config=None, config=None):
    n_gpu = torch.cuda.device_count()
    need = number_of_actors + number_of_learners
    share = bool(config.get("stub_generators_share_gpus", n_gpu < need))
    if not share and n_gpu < need:
        raise RuntimeError(f"Not enough GPUs available. Available: {n_gpu}, Required: {need}")
    if n_gpu < number_of_learners and not config.get("learner_gpus"):
        raise RuntimeError(f"Not enough GPUs for the learners. Available: {n_gpu}, Required: {number_of_learners}")
    # reference: actors on the first GPUs, learners on the next ones (:533-537); stub generators only need a GPU for the
    # adapter pull, so on a small box they share the learners' GPUs
    learner_gpus = list(range(number_of_learners)) if share else list(range(number_of_actors, need))
    if config.get("learner_gpus"):   # explicit placement (tests: several learners on one GPU, each on its own stream)
        learner_gpus = list(config["learner_gpus"])
        assert len(learner_gpus) == number_of_learners
    actor_gpus = [learner_gpus[i % len(learner_gpus)] for i in range(number_of_actors)] if share else list(range(number_of_actors))
    shared_dev = len(set(learner_gpus)) < len(learner_gpus)
    assert config["learner"] in ("grpo", "pg"), "Learner can be only 'pg' or 'grpo'!"
    cls = GRPOLearner if config["learner"] == "grpo" else Learner
    r, alpha = config["max_lora_rank"], config["lora_alpha"]
    P, T, B = config["max_prompt_tokens"], config["max_new_tokens"], config["train_batch_size"]
    cfg = model_config(model_name, r, alpha)
    fuse = max(1, int(config.get("fuse_microbatches", 2)))

    def learner_factory(rank):
        dev = torch.device("cuda", learner_gpus[rank])

        def make():
            build = lambda **bufs: _build_policy(model_name, cfg, dev, fuse * B, P, T, r, alpha, bufs)
            if number_of_learners > 1:
                if cfg is None:
                    raise NotImplementedError("multi-learner P2P with a checkpoint: pass its LMConfig (lora_numel is needed before the load)")
                return cls.create_for_p2p(rank, number_of_learners, dev, build, IdTokenizer(), config, lora_numel(cfg, fuse * B, P, T))
            return cls(build(), IdTokenizer(), config, gpu_id=learner_gpus[rank])
        return make, dev

    learners = [local_rpc.ActorHandle(*learner_factory(i), own_stream=shared_dev) for i in range(number_of_learners)]
    if number_of_learners > 1:   # same-process peers: plain pointers + peer access instead of IPC handles
        if shared_dev:           # learners sharing a device: every kernel must be loaded before a barrier can spin (warmup())
            for l in learners:
                local_rpc.get(l.warmup.remote())
        groups = local_rpc.get([l.p2p_local_group.remote() for l in learners])
        P2PGroup.wire_same_process(groups, threaded=True)
        local_rpc.get([l.p2p_attach_local.remote() for l in learners])
    desc = local_rpc.get(learners[0].adapter_publisher.remote())
    vocab = cfg.vocab if cfg is not None else local_rpc.get(learners[0].vocab_size.remote())
    gen_factory = config.get("generator_factory")

    def actor_factory(i):
        dev = torch.device("cuda", actor_gpus[i])

        def make():
            sub = AdapterSubscriber(desc, dev, same_process=True)
            if gen_factory is not None:
                return gen_factory(i, dev, sub)
            return StubGenerator(vocab, config["num_candidates"], T, seed=1000 + i, adapter_subscriber=sub,
                                 tokens_per_second=config.get("stub_tokens_per_second"))
        return make, dev

    actors = [local_rpc.ActorHandle(*actor_factory(i)) for i in range(number_of_actors)]
    return actors, learners
