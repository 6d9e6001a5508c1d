"""Multi-learner exchange on ONE device ("fake multi-GPU on one device"): the fused one-shot
reduce + Adam kernel and the flag barrier (csrc/optim.cu) with world = 2 / 4 / 8 buffer sets on a single GPU, one CUDA
stream per "rank", against  mean-of-gradients + torch.optim.Adam  (reference semantics: distributed_actor.py:311-323
mean over learners, :331-333 step; distributed_trainer.py:325-342).  Plus the cross-PROCESS path (CUDA IPC handles, the
Ray flow of INTEGRATION.md) with two processes sharing the GPU.

Tolerances: parameters bit-identical across ranks; vs torch.optim.Adam on the rank-order fp32 mean: <= 2 ulp
(rtol 3e-7), as for the single-learner kernel (test_adam_matches_torch)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


class _FlatPolicy:
    """The attributes P2PGroup.reduce_adam_step touches, without a model behind them."""

    def __init__(self, flat, grad):
        self.lora_flat, self.lora_grad = flat, grad
        self.adam_m, self.adam_v = torch.zeros_like(flat), torch.zeros_like(flat)
        self.opt_step = 0
        self.synced = 0

    def sync_lora(self):
        self.synced += 1


def _make_world(world, n, dev):
    from distrl_llm_b200.p2p import P2PGroup
    groups, pols = [], []
    for r in range(world):
        g = P2PGroup(r, world, dev)
        _, kw = g.alloc_local(n)
        groups.append(g)
        pols.append(_FlatPolicy(kw["lora_flat"], kw["lora_grad"]))
    P2PGroup.wire_same_process(groups)
    return groups, pols


@pytest.mark.parametrize("world", [2, 4, 8, 3])
def test_fake_world_reduce_adam_matches_mean_plus_adam(cuda, world):
    from distrl_llm_b200.p2p import owned_slice
    n = 4 * 100_003                       # not a multiple of world * 4: uneven last slices
    groups, pols = _make_world(world, n, cuda)
    gen = torch.Generator(device=cuda).manual_seed(world)
    p0 = torch.randn(n, device=cuda, generator=gen) * 0.1
    for pol in pols:
        pol.lora_flat.copy_(p0)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3, foreach=False, fused=False)
    streams = [torch.cuda.Stream(device=cuda) for _ in range(world)]
    for pol in pols:            # first use of torch's fill kernel happens here, not while a barrier kernel is spinning
        pol.lora_grad.zero_()   # (CUDA loads a kernel's code at its first launch, which synchronises with the device)
    groups[0].reset_status()    # the status word is process-wide and sticky: start clean whatever ran before
    torch.cuda.synchronize()
    covered = torch.zeros(n, dtype=torch.int32)
    for r in range(world):
        lo, hi = owned_slice(n, world, r)
        covered[lo:hi] += 1
    assert (covered == 1).all()           # the slices partition the buffer
    for step in range(1, 4):
        grads = [torch.randn(n, device=cuda, generator=gen) * 10.0 ** (step - 2) for _ in range(world)]
        mean = grads[0].clone()
        for g in grads[1:]:
            mean += g                      # rank order, fp32: same summation order as the kernel
        mean *= 1.0 / world
        ref_p.grad = mean
        opt.step()
        for pol, g in zip(pols, grads):
            pol.lora_grad.copy_(g)
        torch.cuda.synchronize()
        for r in range(world):            # every "learner" on its own stream; the barrier kernels spin until all arrived
            with torch.cuda.stream(streams[r]):
                groups[r].reduce_adam_step(pols[r], 1e-3, timing=(r == 0))
        torch.cuda.synchronize()
        groups[0].check(reset=True)     # raises if a barrier gave up (and clears the sticky status for the next test)
        for r in range(world):
            assert torch.equal(pols[r].lora_flat, pols[0].lora_flat), f"rank {r} differs from rank 0 at step {step}"
            assert (pols[r].lora_grad == 0).all() and pols[r].synced == step
        assert torch.allclose(pols[0].lora_flat, ref_p.data, rtol=3e-7, atol=1e-9), (pols[0].lora_flat - ref_p.data).abs().max()
        wait_ms, reduce_ms, refresh_ms = groups[0].exchange_ms()
        assert reduce_ms > 0 and wait_ms >= 0 and refresh_ms >= 0
    for g in groups:
        g.close()


def test_barrier_timeout_reports_instead_of_trapping(cuda):
    """A learner that never arrives (the old kernel called __trap after ~10 s and killed every waiting
    context): the waiter gives up after the timeout, the reduce kernel leaves the buffers untouched, check() raises, and
    the context is still usable."""
    n = 4096
    groups, pols = _make_world(2, n, cuda)
    pols[0].lora_flat.fill_(1.0)
    pols[0].lora_grad.fill_(0.5)
    before = pols[0].lora_flat.clone()
    groups[0].barrier(timeout_s=0.2)       # rank 1 never launches its side
    pols[0].opt_step = 1
    groups[0].reduce_adam_step(pols[0], 1e-3)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="timed out waiting for learner 1"):
        groups[0].check(reset=True)
    assert torch.equal(pols[0].lora_flat, before), "a failed exchange must not touch the parameters"
    groups[0].check()                      # status cleared
    assert float((pols[0].lora_flat * 2).sum()) == 2.0 * n   # context alive
    for g in groups:
        g.close()


# ---------------------------------------------------------------------------------------------------
# two PROCESSES on one GPU: create_for_p2p -> export handles -> open -> step (the Ray flow of INTEGRATION.md)
# ---------------------------------------------------------------------------------------------------
def _ipc_worker(rank, world, q_in, q_out):
    import numpy as np
    os.environ["B200RL_P2P_TIMEOUT_S"] = "120"
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.p2p import lora_numel
    from distrl_llm_b200.policy import LMConfig, Policy
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    cfg = LMConfig(vocab=1024, hidden=256, inter=512, n_layers=2, n_q_heads=4, n_kv_heads=2, head_dim=64)
    P, T, B = 8, 24, 4
    config = {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 1e-3}
    ln = GRPOLearner.create_for_p2p(rank, world, dev, lambda **b: Policy.random_init(cfg, dev, B, P, T, seed=5, **b),
                                    IdTokenizer(), config, lora_numel(cfg, B, P, T))
    q_out.put((rank, ln.p2p_export_handles()))
    all_handles = q_in.get()
    ln.p2p_open_handles(all_handles)
    rng = np.random.default_rng(100 + rank)      # each learner its own shard (distributed_trainer.py:312-322)
    prompts = [rng.integers(1, 1024, size=P).tolist() for _ in range(B)]
    answers = [rng.integers(1, 1024, size=int(rng.integers(6, T + 1))).tolist() for _ in range(B)]
    adv = rng.normal(size=B)
    p0 = ln.policy.lora_flat.detach().cpu().clone()
    grads, loss = ln.compute_gradients((prompts, answers, list(adv)))
    assert grads == {}                            # P2P mode: gradients stay on the device
    g_local = ln.policy.lora_grad.detach().cpu().clone()
    ln.apply_merged_gradients()
    torch.cuda.synchronize()
    ln.p2p.check()
    q_out.put((rank, p0, g_local, ln.policy.lora_flat.detach().cpu().clone()))
    q_in.get()                                    # keep the allocations alive until the parent has compared
    ln.p2p.close()


def test_two_process_ipc_exchange(cuda):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    world = 2
    q_out = ctx.Queue()
    q_ins = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, q_ins[r], q_out)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        handles = dict(q_out.get(timeout=600) for _ in range(world))
        for q in q_ins:
            q.put([handles[r] for r in range(world)])
        res = {}
        for _ in range(world):
            r, p0, g, p1 = q_out.get(timeout=600)
            res[r] = (p0, g, p1)
        assert torch.equal(res[0][0], res[1][0]), "same seed: both learners start from identical adapters"
        assert torch.equal(res[0][2], res[1][2]), "after the exchange every learner holds identical parameters (fixes quirk Q4)"
        assert not torch.equal(res[0][1], res[1][1]), "the two learners computed gradients on different shards"
        ref = torch.nn.Parameter(res[0][0].clone())
        opt = torch.optim.Adam([ref], lr=1e-3, foreach=False, fused=False)
        ref.grad = (res[0][1] + res[1][1]) * 0.5
        opt.step()
        assert torch.allclose(res[0][2], ref.data, rtol=3e-7, atol=1e-9)
    finally:
        for q in q_ins:
            q.put(None)
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
