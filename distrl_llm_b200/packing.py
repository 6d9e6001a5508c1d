"""Host-side construction of the packed ("shared-prompt") micro-batch layout.

The reference pads every (prompt, completion) pair to P + T tokens and runs the model on all B*(P+T) rows
(distributed_actor.py:217-243).  In its pipeline the n completions of a problem share the prompt verbatim
(`task["problem"] = [[p for _ in range(n)] for p in ...]`, distributed_actor.py:169-170), and under causal
attention the prompt positions never see the completion, so their activations are identical for the whole group.
The packed layout stores each distinct prompt of a micro-batch ONCE:

    rows = [ prompt 0 | prompt 1 | ... | completion 0 | completion 1 | ... ]      G*P + B*T rows

Completion queries attend to the whole prompt segment of their group (prefix) and causally to their own segment.
Everything here is integer bookkeeping (numpy); the arrays go to the device in one int32 blob and are consumed by
b200rl_model_microbatch_packed / the tcgen05 attention kernels through block descriptors.

Groups are runs of CONSECUTIVE sequences with identical padded prompt rows (ids and mask); a micro-batch without any
repeated prompt degenerates to G = B (same code path, no saving).

Ragged rows (`ragged=True`, the default): the pad tokens the reference feeds through the model (prompts left-padded to
P, completions right-padded to T, distributed_actor.py:217-239) are masked out of attention and of the loss, so they
influence nothing.  The packed layout simply does not store them: a prompt segment has its real length, a completion
segment its real length, and every kept row carries its ORIGINAL position (index in the padded row) for RoPE, which
keeps the arithmetic of the surviving rows identical.  Scored positions beyond a completion's length point at row 0
and have coefficient 0.  SURVEY.md 8(f) N2: mean completion length is 450-500 of T = 1200 in the reference's runs.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

QB_FIELDS = 8   # b200rl_attn_qblock: q_row0 q_rows q_local0 own_row0 own_len pre_row0 pre_len stat0
KB_FIELDS = 8   # b200rl_attn_kblock: k_row0 k_rows k_local0 q_row0 q_len causal stat0 out_row0


class PackedBatchC(C.Structure):
    _fields_ = [("rows", C.c_int), ("B", C.c_int), ("T", C.c_int), ("max_pos", C.c_int),
                ("n_qblocks", C.c_int), ("n_kblocks", C.c_int), ("part_rows", C.c_int),
                ("ids", C.c_void_p), ("pos", C.c_void_p), ("key_mask", C.c_void_p), ("score_src", C.c_void_p),
                ("targets", C.c_void_p), ("answer_mask", C.c_void_p), ("sc_start", C.c_void_p), ("sc_list", C.c_void_p),
                ("qblocks", C.c_void_p), ("kblocks", C.c_void_p), ("red_start", C.c_void_p), ("red_list", C.c_void_p),
                ("n_score", C.c_int), ("score_slot", C.c_void_p)]


@dataclass
class PackedHost:
    """numpy arrays of one packed micro-batch (all int32)."""
    rows: int
    B: int
    P: int
    T: int
    n_groups: int
    part_rows: int
    arrays: dict          # name -> np.int32 array
    seq_group: np.ndarray  # [B] group of every sequence
    prompt_row0: object = None   # [G] first packed row of every prompt segment
    prompt_ext: object = None    # [G] (first kept padded index, rows)
    comp_row0: object = None     # [B] first packed row of every completion segment
    comp_len: object = None      # [B] rows of every completion segment

    def blob(self):
        """Concatenate all arrays into one int32 vector; returns (blob, {name: (offset, length)})."""
        offs, parts, o = {}, [], 0
        for k, v in self.arrays.items():
            v = np.ascontiguousarray(v, dtype=np.int32).reshape(-1)
            pad = (-o) % 4  # keep every array 16-byte aligned
            if pad:
                parts.append(np.zeros(pad, np.int32))
                o += pad
            offs[k] = (o, v.size)
            parts.append(v)
            o += v.size
        return np.concatenate(parts), offs


def pack_microbatch(ids: np.ndarray, attn_mask: np.ndarray, P: int, T: int, ragged: bool = True,
                    compact_scored: bool = True) -> PackedHost:
    """ids / attn_mask: [B, P+T] (prompt left-padded to P, completion right-padded to T, reference layout).
    compact_scored: the head (final norm, lm_head, log-softmax, lm_head dX) runs on the scored positions with
    answer_mask 1 only; False keeps all B*T positions like the reference (distributed_actor.py:245-260)."""
    ids = np.asarray(ids, dtype=np.int32)
    am = np.asarray(attn_mask, dtype=np.int32)
    B, L = ids.shape
    assert L == P + T
    # ---- groups = runs of identical prompts ----
    seq_group = np.zeros(B, np.int32)
    g = 0
    for i in range(1, B):
        same = np.array_equal(ids[i, :P], ids[i - 1, :P]) and np.array_equal(am[i, :P], am[i - 1, :P])
        if not same:
            g += 1
        seq_group[i] = g
    G = g + 1
    first_of_group = [int(np.argmax(seq_group == k)) for k in range(G)]
    # ---- segments: (first padded index kept, number of rows kept).  Ragged: the contiguous run of real tokens
    # (left-padded prompt = suffix, right-padded completion = prefix); a mask with holes keeps the padded extent.
    def extent(mask, left_padded):
        if not ragged:
            return 0, mask.size
        nz = np.flatnonzero(mask)
        if nz.size == 0:
            return (mask.size, 0) if left_padded else (0, 0)
        if left_padded:
            return int(nz[0]), mask.size - int(nz[0])
        return 0, int(nz[-1]) + 1
    p_ext = [extent(am[first_of_group[k], :P], True) for k in range(G)]        # (start index in [0,P), rows)
    c_ext = [extent(am[i, P:], False) for i in range(B)]                       # (0, rows)
    p_row0 = np.zeros(G, np.int64)
    c_row0 = np.zeros(B, np.int64)
    r = 0
    for k in range(G):
        p_row0[k] = r
        r += p_ext[k][1]
    for i in range(B):
        c_row0[i] = r
        r += c_ext[i][1]
    rows = max(int(r), 1)
    p_ids = np.zeros(rows, np.int32)
    p_pos = np.zeros(rows, np.int32)
    p_km = np.zeros(rows, np.int32)
    for k in range(G):
        i, (st, n) = first_of_group[k], p_ext[k]
        r0 = int(p_row0[k])
        p_ids[r0:r0 + n] = ids[i, st:st + n]
        p_km[r0:r0 + n] = am[i, st:st + n]
        p_pos[r0:r0 + n] = st + np.arange(n)
    for i in range(B):
        r0, n = int(c_row0[i]), c_ext[i][1]
        p_ids[r0:r0 + n] = ids[i, P:P + n]
        p_km[r0:r0 + n] = am[i, P:P + n]
        p_pos[r0:r0 + n] = P + np.arange(n)
    # ---- scored rows: the logit at original position P-1+t predicts completion token t (distributed_actor.py:245-249)
    targets = ids[:, P:].reshape(-1).astype(np.int32)
    answer_mask = am[:, P:].reshape(-1).astype(np.int32)
    score_src = np.zeros(B * T, np.int32)
    live = np.zeros(B * T, bool)            # scored positions that exist in the packed rows
    for i in range(B):
        gk = int(seq_group[i])
        n = c_ext[i][1]
        if p_ext[gk][1] > 0 and n > 0:      # t = 0 reads the last prompt row (shared by the group)
            score_src[i * T] = p_row0[gk] + p_ext[gk][1] - 1
            live[i * T] = True
        if n > 1:
            score_src[i * T + 1:i * T + n] = c_row0[i] + np.arange(n - 1)
            live[i * T + 1:i * T + n] = True
    # positions without a source row must not be trained on (they are pad positions: answer_mask is already 0 there,
    # except when a sequence has an empty prompt, which the reference cannot produce)
    answer_mask = np.where(live, answer_mask, 0).astype(np.int32)
    if compact_scored:
        # scored rows = the positions that carry loss, in slot order (keeps at least one row: an all-masked micro-batch
        # still runs the head once, with coefficient 0)
        score_slot = np.flatnonzero(answer_mask != 0).astype(np.int32)
        if score_slot.size == 0:
            score_slot = np.zeros(1, np.int32)
        if score_slot.size == B * T:
            score_slot = np.zeros(0, np.int32)      # nothing to drop: identity, no indirection
        else:
            live_c = live[score_slot]
            score_src = score_src[score_slot]
            targets = targets[score_slot]
            live = live_c
    else:
        score_slot = np.zeros(0, np.int32)
    live_idx = np.flatnonzero(live)                  # indices of scored ROWS (compacted or not) that have a source row
    order = live_idx[np.argsort(score_src[live_idx], kind="stable")]
    counts = np.bincount(score_src[live_idx], minlength=rows)
    sc_start = np.zeros(rows + 1, np.int32)
    sc_start[1:] = np.cumsum(counts)
    sc_list = order.astype(np.int32)
    # ---- attention block descriptors ----
    segs = [(int(p_row0[k]), p_ext[k][1], 0, 0) for k in range(G)]                # (row0, len, pre_row0, pre_len)
    segs += [(int(c_row0[i]), c_ext[i][1], int(p_row0[seq_group[i]]), p_ext[int(seq_group[i])][1]) for i in range(B)]
    qb = []
    for (r0, ln, pr0, pl) in segs:
        for b0 in range(0, ln, 128):
            work = (pl + 127) // 128 + (min(b0 + 128, ln) + 127) // 128
            qb.append((work, [r0 + b0, min(128, ln - b0), b0, r0, ln, pr0, pl, r0 + b0]))
    qb.sort(key=lambda t: -t[0])                                                  # heavy blocks first
    qblocks = np.array([t[1] for t in qb], np.int32).reshape(-1, QB_FIELDS)
    kb, part_rows = [], 0
    deps = [[] for _ in range(G)]
    for i in range(B):
        deps[int(seq_group[i])].append(i)
    for si, (r0, ln, pr0, pl) in enumerate(segs):
        qsegs = [(r0, ln, 1)]                                                     # own queries (causal)
        if si < G:
            qsegs += [(int(c_row0[i]), c_ext[i][1], 0) for i in deps[si] if c_ext[i][1] > 0]   # the group's completions
        for b0 in range(0, ln, 128):
            k_rows = min(128, ln - b0)
            for (qr0, qlen, causal) in qsegs:
                nqb = (qlen + 63) // 64 - (b0 // 64 if causal else 0)
                kb.append((nqb, [r0 + b0, k_rows, b0, qr0, qlen, causal, qr0, part_rows]))
                part_rows += k_rows
    # reduction lists: key row r sums the partial rows of every (key block, query segment) unit that covers it, in the
    # order the units were created (fixed order -> deterministic dK/dV).  Vectorised: one (row, partial row) pair per
    # partial row, stable-sorted by key row.
    red_start = np.zeros(rows + 1, np.int32)
    if kb:
        kb_arr = np.array([t[1] for t in kb], np.int64).reshape(-1, KB_FIELDS)
        k_row0, k_rows_a, out0 = kb_arr[:, 0], kb_arr[:, 1], kb_arr[:, 7]
        rep = np.repeat(np.arange(kb_arr.shape[0]), k_rows_a)
        within = np.arange(int(k_rows_a.sum())) - np.repeat(np.cumsum(k_rows_a) - k_rows_a, k_rows_a)
        key_row = k_row0[rep] + within
        part_row = out0[rep] + within                                             # == arange(part_rows): creation order
        order_r = np.argsort(key_row, kind="stable")
        red_list = part_row[order_r].astype(np.int32)
        red_start[1:] = np.cumsum(np.bincount(key_row, minlength=rows))
    else:
        red_list = np.zeros(0, np.int32)
    kb.sort(key=lambda t: -t[0])
    kblocks = np.array([t[1] for t in kb], np.int32).reshape(-1, KB_FIELDS)
    arrays = {"ids": p_ids, "pos": p_pos, "key_mask": p_km, "score_src": score_src, "targets": targets,
              "answer_mask": answer_mask, "sc_start": sc_start, "sc_list": sc_list, "qblocks": qblocks.reshape(-1),
              "kblocks": kblocks.reshape(-1), "red_start": red_start, "red_list": red_list, "score_slot": score_slot}
    host = PackedHost(rows=rows, B=B, P=P, T=T, n_groups=G, part_rows=max(part_rows, 1), arrays=arrays, seq_group=seq_group)
    host.prompt_row0, host.prompt_ext = p_row0, p_ext       # (first kept padded index, rows) per group
    host.comp_row0, host.comp_len = c_row0, [e[1] for e in c_ext]
    return host


class PinnedStager:
    """Reusable pinned host staging buffers for the per-pass host->device copies (ids / masks / descriptors /
    advantages).  `tensor.pin_memory()` allocates and registers fresh pinned memory on every call (~0.1-1 ms each);
    here a small ring of buffers per power-of-two size class is allocated once, and a buffer is reused only after the
    copy that last read it has completed (CUDA event)."""

    def __init__(self, device, ring=4):
        self.device = torch.device(device)
        self.ring = ring
        self._bufs = {}     # size class -> [(uint8 pinned tensor, event or None)]
        self._next = {}

    def to_device(self, arr):
        """numpy array or CPU tensor -> new device tensor with the same dtype / shape (async copy on the current stream)."""
        t = torch.from_numpy(np.ascontiguousarray(arr)) if isinstance(arr, np.ndarray) else arr.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes == 0:
            return t.to(self.device)
        cls = 1 << max(12, (nbytes - 1).bit_length())
        slots = self._bufs.setdefault(cls, [])
        i = self._next.get(cls, 0)
        if len(slots) < self.ring:
            slots.append([torch.empty(cls, dtype=torch.uint8).pin_memory(), None])
            i = len(slots) - 1
        self._next[cls] = (i + 1) % self.ring
        buf, ev = slots[i]
        if ev is not None:
            ev.synchronize()
        stage = buf[:nbytes].view(t.dtype).view(t.shape)
        stage.copy_(t)
        out = stage.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slots[i][1] = ev
        return out


class PackedDevice:
    """Device-resident packed micro-batch: one int32 blob + the C descriptor struct pointing into it."""

    def __init__(self, host: PackedHost, device, pinned=True, stager: "PinnedStager" = None):
        blob, offs = host.blob()
        if stager is not None:
            self.blob = stager.to_device(blob)
        else:
            t = torch.from_numpy(blob)
            if pinned:
                t = t.pin_memory()
            self.blob = t.to(device, non_blocking=True)
        self.host = host
        self.h2d_bytes = blob.nbytes
        base = self.blob.data_ptr()

        def p(name):
            return base + 4 * offs[name][0]

        self.answer_mask = self.blob[offs["answer_mask"][0]:offs["answer_mask"][0] + offs["answer_mask"][1]].view(host.B, host.T)
        self.c = PackedBatchC(host.rows, host.B, host.T, host.P + host.T, offs["qblocks"][1] // QB_FIELDS,
                              offs["kblocks"][1] // KB_FIELDS, host.part_rows, p("ids"), p("pos"), p("key_mask"),
                              p("score_src"), p("targets"), p("answer_mask"), p("sc_start"), p("sc_list"),
                              p("qblocks"), p("kblocks"), p("red_start"), p("red_list"),
                              offs["score_slot"][1], p("score_slot") if offs["score_slot"][1] else None)
        self.n_score = offs["score_slot"][1] or host.B * host.T
