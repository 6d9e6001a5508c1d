#!/usr/bin/env python
"""bench.py — learner tokens processed/sec (GRPO step, Qwen2.5-7B LoRA) on B200.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU).

Workload (BASELINE.json configs[1]): GRPO learner, Qwen2.5-7B-shaped random-init NF4 base + rank-16
LoRA, group_size 8, 64 completions of length 512 (P=350 prompt tokens, micro-batch 8 -> 8 micro-batches),
per GPU.  A "step" = one learner update: zero_grad, the 8 reference micro-batches (forward, fused log-prob/loss,
backward into LoRA grads) as 4 model passes of 2 micro-batches each (--fuse_microbatches; gradient accumulation is
linear, identical result) in the packed shared-prompt layout, (P2P reduce +) Adam on the LoRA parameters, refresh of
the bf16 LoRA operands.
  value  : completion tokens scored+updated / s with the batch already resident in HBM
  e2e    : same through the reference-shaped public API GRPOLearner.train(candidates) with HOST
           token-id lists (CPU padding + pinned H2D copies + D2H of the loss inside the timed region)
  N > 1  : weak scaling (each learner owns 64 sequences), gradient mean + Adam through the one-shot
           P2P reduce kernel over NVLink; time = max over ranks.
--impl reference: the reference learner's CPU path (the pinned torch oracle port of its code, HF-style fp32
math on all host cores) timed on a bounded sample of the same workload and extrapolated (see cpu_sample()).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CATS = ["gemm", "gemm_lora_skinny", "gemm_dw", "nf4_dequant", "attn_fwd", "attn_bwd", "row_kernels", "logprob", "misc"]


def workload(args):
    return dict(n_seq=args.seqs, P=args.prompt_len, T=args.new_tokens, B=args.micro_batch, group=args.group_size,
                rank=args.lora_rank)


# ---------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc = None
        self.lines = []
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# algorithmic work
# ---------------------------------------------------------------------------------------------------
def flops_per_sequence(P, T, r=16):
    L = P + T
    p_layers, p_lm = 6_525_288_448, 544_997_376
    p_lora = r * 90_112 * 28
    return L * (4 * p_layers + 6 * p_lora) + T * 4 * p_lm + 3 * 28 * 2 * L * L * 3584


# ---------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference learner) on a bounded sample
# ---------------------------------------------------------------------------------------------------
def cpu_sample(w, threads=None, n_rows=None):
    """Times the reference learner's CPU math (oracle/learner_oracle.py = restatement of
    distributed_actor.py:215-261, :440-493) on ONE micro-batch of the workload through 1-layer and 2-layer
    full-width, full-vocab slices of the model, fits t = head + layers * per_layer and extrapolates to
    28 layers x n_microbatches.  Returns (tokens/s extrapolated, description)."""
    from oracle import learner_oracle as lo
    if threads:
        torch.set_num_threads(threads)
    B = n_rows or w["B"]
    P, T = w["P"], w["T"]
    times = {}
    for nl in (1, 2):
        cfg = lo.OracleConfig(vocab=152064, hidden=3584, inter=18944, n_layers=nl, n_q_heads=28, n_kv_heads=4,
                              head_dim=128, lora_r=w["rank"], lora_alpha=16)
        params, _ = lo.make_params(cfg, seed=1, quantize_base=False)
        prompts, answers, rewards = lo.make_batch(cfg, B, P, T, seed=2, ragged=False, group_size=B, learner="grpo")
        ids, am, ansm = lo.pad_batch(prompts, answers, P, T)
        t0 = time.perf_counter()
        lo.compute_gradients(params, cfg, ids, am, ansm, rewards, P, B, "grpo")
        times[nl] = time.perf_counter() - t0
        del params
    per_layer = max(times[2] - times[1], 1e-9)
    head = max(times[1] - per_layer, 0.0)
    t_mb = head + 28 * per_layer
    nb = (w["n_seq"] + w["B"] - 1) // w["B"]
    t_step = t_mb * nb * (w["B"] / B)
    tok_s = w["n_seq"] * T / t_step
    desc = (f"oracle port of the reference learner, fp32 torch on {torch.get_num_threads()} host threads: one micro-batch "
            f"({B}x{P + T} tokens) fwd+bwd through 1- and 2-layer full-width/full-vocab slices "
            f"({times[1]:.1f}s, {times[2]:.1f}s), extrapolated to 28 layers x {nb} micro-batches")
    return tok_s, desc, t_step


def host_threads():
    """All the host cores the CPU arm can use: torchrun exports OMP_NUM_THREADS=1, which would silently turn the
    reference arm into a single-thread run (measured 3.3 instead of 12.9 tok/s).  One thread per physical core."""
    forced = int(os.environ.get("B200RL_CPU_THREADS", "0"))
    return forced if forced > 0 else max(1, (os.cpu_count() or 2) // 2)


def run_reference(args, rank, world):
    w = workload(args)
    if rank != 0:
        return
    vals = []
    threads = host_threads()
    for i in range(args.warmup + args.steps):
        tok_s, desc, t_step = cpu_sample(w, threads=threads, n_rows=args.cpu_rows)
        if i >= args.warmup:
            vals.append((tok_s, t_step))
    tok = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": "learner tokens processed/sec (GRPO step, Qwen2.5-7B LoRA)", "value": tok,
            "unit": "completion tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args),
            "cpu_baseline": {"value": tok, "unit": "completion tokens/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": tok, "unit": "completion tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args):
    return {"workload": f"GRPO learner step, Qwen2.5-7B-shaped random-init NF4 base + rank-{args.lora_rank} LoRA, "
                        f"group_size={args.group_size}, {args.seqs} completions len={args.new_tokens} per GPU "
                        f"(P={args.prompt_len}, micro-batch {args.micro_batch})"
                        + (", RAGGED lengths (prompt ~ U[P/2,P], completion ~ U[T/4,T]); value counts real completion tokens" if getattr(args, "ragged", False) else ""),
            "global_batch": args.seqs * args.gpus, "seq_len": args.prompt_len + args.new_tokens,
            "parallelism": f"dp{args.gpus}" if args.gpus > 1 else "single learner",
            "passes": f"{getattr(args, 'fuse_microbatches', 1)} reference micro-batches of 8 per model pass (gradient accumulation is linear: identical result)",
            "layout": "classic [B, P+T] rows" if getattr(args, "no_share_prompts", False) else
                      "packed shared-prompt rows (each group's prompt processed once; identical gradients)",
            "l2": "per-step activations and weights (>30 GB) far exceed the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seqs", type=int, default=64)
    ap.add_argument("--new_tokens", type=int, default=512)
    ap.add_argument("--prompt_len", type=int, default=350)
    ap.add_argument("--micro_batch", type=int, default=8)
    ap.add_argument("--group_size", type=int, default=8)
    ap.add_argument("--lora_rank", type=int, default=16)
    ap.add_argument("--layers", type=int, default=28, help="debug only: anything but 28 is not the benchmark")
    ap.add_argument("--cpu_rows", type=int, default=2, help="sequences in the CPU sample micro-batch")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--fuse_microbatches", type=int, default=2, help="reference micro-batches per model pass (identical gradients; 1 = one pass per micro-batch like the reference)")
    ap.add_argument("--ragged", action="store_true", help="ragged synthetic lengths (prompt ~ U[P/2,P], completion ~ U[T/4,T]); value counts REAL completion tokens")
    ap.add_argument("--no_share_prompts", action="store_true", help="classic [B, P+T] layout (every prompt recomputed per completion)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    import torch.distributed as dist
    from distrl_llm_b200 import _capi
    from distrl_llm_b200.learner import GRPOLearner, IdTokenizer
    from distrl_llm_b200.policy import LMConfig, Policy
    from distrl_llm_b200.p2p import P2PGroup

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to stdout when the first communicator
        # is created (NCCL_DEBUG=VERSION/WARN in this image), so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        _saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(_saved_fd, 1)
            os.close(_saved_fd)
    w = workload(args)
    cfg = LMConfig.qwen25_7b(lora_r=args.lora_rank)
    cfg.n_layers = args.layers
    P, T, B, N = w["P"], w["T"], w["B"], w["n_seq"]
    group = None
    kw = {}
    FK = max(1, args.fuse_microbatches)
    try:   # long-sequence configs (T = 2048: 38k rows per fused pass) may not leave room for 2 micro-batches per pass
        from distrl_llm_b200.policy import largest_pass_that_fits
        free_b, _ = torch.cuda.mem_get_info(dev)
        FK = largest_pass_that_fits(cfg, B, P, T, FK, free_b, reserve_bytes=24 << 30)   # NF4 + bf16 cache + lm_head/embed
        args.fuse_microbatches = FK
    except Exception as e:  # sizing is best-effort; the allocation itself still fails loudly
        print(f"[bench] pass sizing skipped: {e}", file=sys.stderr)
    if world > 1:
        group, kw = P2PGroup.from_torch_distributed(cfg, FK * B, P, T, dev)
    pol = Policy.random_init(cfg, dev, FK * B, P, T, seed=1234, **kw)   # same base + LoRA on every learner
    if group is not None:
        group.attach(pol)
    config = {"train_batch_size": B, "max_new_tokens": T, "max_prompt_tokens": P, "lr": 2e-5}
    learner = GRPOLearner(pol, IdTokenizer(), config, gpu_id=local)
    if group is not None:
        learner.enable_p2p(group)

    # synthetic batch: ids ~ U[1,V), full-length prompts/completions, rewards -> group advantages
    from distrl_llm_b200.trainer_prep import synthetic_candidates
    cands, flat = synthetic_candidates(cfg.vocab, N, P, T, args.group_size, seed=1234 + rank, ragged=args.ragged)
    prompts, answers, adv = flat
    assert np.all(np.asarray(adv) != 0), "synthetic advantages must be non-zero (quirk Q1 would skip work)"
    nb = (N + B - 1) // B
    # device-resident copy for the `value` measurement
    ids_h, am_h, ansm_h = learner._encode(prompts, answers)
    d_ids, d_am, d_ansm = ids_h.to(dev), am_h.to(dev), ansm_h.to(dev)
    d_adv = torch.from_numpy(np.asarray(adv, dtype=np.float64)).to(dev)

    share = learner.share_prompts and not args.no_share_prompts
    learner.share_prompts = share
    assert learner.fuse_microbatches == FK
    # passes of FK reference micro-batches each (learner.compute_loss does the same planning on the e2e path)
    bounds = [(i * B, min((i + 1) * B, N)) for i in range(nb)]
    passes, cur = [], []
    for (s_, e_) in bounds:
        if e_ - s_ == B and FK > 1:
            cur.append((s_, e_))
            if len(cur) == FK:
                passes.append(cur); cur = []
        else:
            passes.append([(s_, e_)])
    if cur:
        passes.append(cur)
    pass_rng = [(g[0][0], g[-1][1], len(g)) for g in passes]      # contiguous because only full micro-batches fuse
    d_adv_k = [d_adv[s_:e_] * float(k) for (s_, e_, k) in pass_rng]
    packed = []
    if share:  # packed shared-prompt layout, device-resident for the `value` measurement
        from distrl_llm_b200 import packing
        for (s_, e_, k) in pass_rng:
            packed.append(packing.PackedDevice(packing.pack_microbatch(ids_h[s_:e_].numpy(), am_h[s_:e_].numpy(), P, T), dev))
        torch.cuda.synchronize()

    def device_step():
        pol.zero_grad()
        pol.loss_accum.zero_()
        for i, (s, e, k) in enumerate(pass_rng):
            if share:
                pol.microbatch_packed(packed[i], d_adv_k[i], nb, True, backward=True)
            else:
                pol.microbatch(d_ids[s:e], d_am[s:e], d_ansm[s:e], d_adv_k[i], P, T, nb, True, backward=True)
        if group is not None:
            group.reduce_adam_step(pol, learner.lr, 0.0)
        else:
            pol.optimizer_step(learner.lr)

    def e2e_step():
        if group is not None:
            loss = learner.compute_loss(prompts, answers, adv)
            learner.apply_merged_gradients()
        else:
            loss = learner.train(cands)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _capi.lib().b200rl_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = _capi.lib().b200rl_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, clocks, launches

    for _ in range(args.warmup):
        device_step()
    if os.environ.get("B200RL_PROFILE_ONE_STEP"):
        # ncu --profile-from-start off: exactly one learner step between cudaProfilerStart/Stop
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        device_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    ms_dev, clocks, launches = timed(device_step, args.steps, sample_clocks=True)
    for _ in range(2):
        e2e_step()
    ms_e2e, _, _ = timed(e2e_step, args.steps)
    ms_dev2, _, _ = timed(device_step, args.steps)   # diagnostic: same region without the nvidia-smi sampler

    # per-category CUDA-event profile of ONE extra step (same stream; events between consecutive launches)
    import ctypes as C
    _capi.check(_capi.lib().b200rl_model_profile(pol.handle, 1))
    device_step()
    ms_c, wk_c, cnt_c = (C.c_double * 9)(), (C.c_double * 9)(), (C.c_longlong * 9)()
    _capi.check(_capi.lib().b200rl_model_profile_read(pol.handle, ms_c, wk_c, cnt_c))
    _capi.check(_capi.lib().b200rl_model_profile(pol.handle, 0))
    prof = {CATS[i]: {"ms": ms_c[i], "work": wk_c[i], "launches": cnt_c[i]} for i in range(9)}

    tokens = N * T * world if not args.ragged else int(sum(len(a) for a in answers)) * world   # ragged: this rank's real tokens x N (same distribution)
    value = tokens / (ms_dev / 1e3)
    e2e_val = tokens / (ms_e2e / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    gemm = prof["gemm"]
    ach_tf = gemm["work"] / (gemm["ms"] / 1e3) / 1e12 if gemm["ms"] > 0 else 0.0
    step_flops = flops_per_sequence(P, T, args.lora_rank) * N * (args.layers / 28.0 if args.layers != 28 else 1.0)
    traffic = None
    try:  # dram__bytes_read + dram__bytes_write per GEMM launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_dram_traffic.json")))["per_launch_traffic_bytes"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_pair_kernel<256> (tcgen05 cta_group::2, base+LoRA mainloop)",
                "achieved": round(ach_tf, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach_tf / peak_tf, 4),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF sustained",
                "traffic": traffic, "traffic_unit": "bytes per GEMM launch (ncu dram read+write, profiles/r1_gemm_dram_traffic.json)",
                "gemm_share_of_step": round(gemm["ms"] / max(sum(p["ms"] for p in prof.values()), 1e-9), 4),
                "avg_launch_ms": round(gemm["ms"] / max(gemm["launches"], 1), 4),
                "flops_per_launch": gemm["work"] / max(gemm["launches"], 1),
                "reference_layout_equiv_tflops": round(step_flops * world / (ms_dev / 1e3) / 1e12 / world, 1),  # SURVEY 8d FLOPs of the unpacked layout / our time
                "how": "CUDA events between consecutive launches on the launching stream, one profiled step after the timed region"}
    h2d = int(sum(pk.h2d_bytes for pk in packed) + N * 8) if share else int(ids_h.numel() * 4 + am_h.numel() * 4 + ansm_h.numel() * 4 + N * 8)
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and args.layers == 28:
            tok_s, desc, _ = cpu_sample(w, threads=host_threads(), n_rows=args.cpu_rows)
            cpu = {"value": tok_s, "unit": "completion tokens/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": desc}
        line = {"metric": "learner tokens processed/sec (GRPO step, Qwen2.5-7B LoRA)", "value": value,
                "unit": "completion tokens/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config_dict(args),
                "e2e": {"value": e2e_val, "unit": "completion tokens/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
                "ms_per_step_no_sampler": ms_dev2, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "profile_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                "profile_launches": {k: int(v["launches"]) for k, v in prof.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
