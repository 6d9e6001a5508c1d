"""distrl_llm_b200 — B200-native GRPO/PG learner hot path behind DistRL-LLM's Learner API.

The compute lives in lib/libb200rl.so (hand-written sm_100a CUDA, C ABI in include/b200rl.h); this package is
the ctypes shim plus the Python mirror of the reference's learner classes.  Importing the package never loads
the library; constructing a Policy / calling an op does, and fails loudly if it is missing (no fallback).
"""
__all__ = ["_capi", "ops", "policy", "learner", "p2p", "trainer_prep"]
