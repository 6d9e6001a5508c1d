// G4 on the 5th-gen tensor cores (head_dim 128): causal GQA flash attention forward + backward with tcgen05.mma,
// accumulators in TMEM, tiles by TMA.  Other head dims use the mma.sync kernels in attention.cu.
//
// The kernels are driven by small block descriptors so that ONE implementation serves two token layouts:
//   classic : B sequences of L = P + T rows each (the reference's padded batch, distributed_actor.py:233-239);
//             descriptors are computed from blockIdx on the device.
//   packed  : "shared-prompt" layout — every distinct prompt of the micro-batch is stored ONCE (a causal segment),
//             each completion is its own segment whose queries additionally see the whole prompt segment as a
//             prefix.  In GRPO all completions of a group share the prompt (distributed_actor.py:169-170 repeats it
//             n times), so the token-parallel work of the 28 layers drops from B*(P+T) to G*P + B*T rows with the
//             same attention inputs per query.  Descriptors come from arrays built on the host.
//
// Forward (one CTA per <=128-query block and q head):
//   warp 0 / 3 : TMA producers (Q + K ring / V ring)    warp 1 : issues S_j = Q.K_j^T    warp 2 : TMEM allocator, issues
//   O += P_j.V_j    (every single-lane role is entered through elect.sync, see common.cuh)
//   warps 4-11 : two softmax warpgroups; thread = query row (TMEM lane); warpgroup wg owns keys [64wg,64wg+64) of
//                each 128-key block and output columns [64wg, 64wg+64).  O lives in TMEM for the whole key loop (lazy
//                reference maximum); P_j is written with tcgen05.st over S_j and is the TMEM A operand of P_j.V_j.
// Backward: dQ kernel per query block (64-key inner blocks, dQ accumulates in TMEM); dK/dV kernel per key block and
// per query segment that sees it (64-query inner blocks over all q heads of the GQA group, dK/dV accumulate in TMEM;
// in the packed layout the per-(key block, query segment) partials are fp32 slabs summed in fixed order).  dS / P^T / dS^T
// are TMEM A operands as well; three MMA-issuing threads per CTA (scores, dP, gradients).
// The design follows three measurements: ~100 cycles per MMA for one issuing thread, the 128 B/clk
// shared-memory port, and the 56 B/clk TMEM read port.
#include "common.cuh"
#include "b200rl.h"
#include <string.h>

namespace b200rl {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;   // queries per CTA (fwd, dQ) / keys per CTA (dKV)
constexpr int BKV = 128;  // keys per block in the forward
constexpr int TILE_BYTES = 128 * HD * 2;  // 32 KB: [2 halves of 64 cols][128 rows][128 B]
constexpr int HALF_TILE = 64 * HD * 2;    // 16 KB: [2 d-halves of 8 KB][64 rows][128 B]

typedef b200rl_attn_qblock QBlock;
typedef b200rl_attn_kblock KBlock;

__device__ __forceinline__ void tma_load_rows(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(col), "r"(row)
      : "memory");
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same descriptor split into its halves so that an issuing thread builds it once per tile and only ADDS a byte
// offset per MMA (the address field holds addr >> 4 in bits 0-13; shared memory is < 256 KB, so no carry leaves it):
// one thread issues every MMA of a role and each instruction it spends per MMA is serial latency (profiles/r2_run11).
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t addr, uint32_t lbo) {
  return ((addr & 0x3FFFFu) >> 4) | (((lbo >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t DESC_HI_SBO1024 = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_at(uint32_t lo, uint32_t byte_off) {
  return ((uint64_t)DESC_HI_SBO1024 << 32) | (uint64_t)(lo + (byte_off >> 4));
}
constexpr uint32_t idesc_n(int n, bool b_mn) {  // M = 128, A K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
constexpr uint32_t idesc_128x128(bool b_mn) { return idesc_n(128, b_mn); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// this thread's row r of a [128 rows][64 cols] bf16 K-major SW128 tile: store columns col0..col0+31
__device__ __forceinline__ void store_row32_sw128(uint32_t tile, int r, int col0, const float* f) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int cc = col0 / 8 + u;  // 16-byte chunk 0..7 within the 128-byte row
    const uint32_t addr = tile + r * 128 + ((cc ^ (r & 7)) << 4);
    const bf16x8 pk = pack8(&f[u * 8]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk.u.x), "r"(pk.u.y), "r"(pk.u.z),
                 "r"(pk.u.w)
                 : "memory");
  }
}

struct AttnParams {
  const int* key_mask;   // [rows] 1 = real token
  bf16* out;             // fwd: [rows, nq*HD]
  float* lse2;           // fwd out / bwd in
  const float* delta;    // bwd in
  bf16* dqkv;            // bwd out [rows, (nq+2nkv)*HD]
  float* kv_part;        // packed dKV: fp32 partial slabs [part rows][2*nkv*HD]
  const QBlock* qblocks; // packed: descriptors (nullptr = classic)
  const KBlock* kblocks;
  int L;                 // classic: sequence length
  int stat_h;            // lse/delta head stride
  int nq, nkv;
  float scale, scale_log2;
  unsigned long long* prof;  // debug (b200rl_attn_set_prof): per-phase clock64 sums of the forward kernel, see PROF_*
};
// slots of AttnParams::prof (cycles summed over CTAs; [16] = key blocks, [17] = CTAs)
enum { PROF_S_WAIT = 0, PROF_S_LD, PROF_MAX_XCHG, PROF_ABSORB_WAIT, PROF_ABSORB, PROF_EXP_STORE, PROF_FENCE_ARRIVE, PROF_LOOP,
       PROF_PROLOGUE, PROF_EPILOGUE, PROF_M_KFULL, PROF_M_SEMPTY, PROF_M_PFULL, PROF_M_VFULL, PROF_M_OEMPTY, PROF_M_TOTAL,
       PROF_BLOCKS, PROF_CTAS, PROF_N };
#define PROF_T(var) const long long var = prof_on ? clock64() : 0

// classic layout descriptors from the launch grid: grid = (ceil(L/128), heads, B), heavy blocks first
__device__ __forceinline__ QBlock classic_qblock(const AttnParams& p) {
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x, b = blockIdx.z;
  QBlock d;
  d.q_row0 = b * p.L + qb * BQ;
  d.q_rows = min(BQ, p.L - qb * BQ);
  d.q_local0 = qb * BQ;
  d.own_row0 = b * p.L;
  d.own_len = p.L;
  d.pre_row0 = 0;
  d.pre_len = 0;
  d.stat0 = b * p.nq * p.L + qb * BQ;
  return d;
}

// key-block iteration shared by fwd (W = 128) and dQ (W = 64): prefix blocks first, then own (causal) blocks
template <int W>
struct KeyIter {
  int n_pre, n_tot;
  __device__ __forceinline__ KeyIter(const QBlock& d) {
    n_pre = (d.pre_len + W - 1) / W;
    const int need = d.q_local0 + d.q_rows;  // own keys with local index < need are visible to some query
    n_tot = n_pre + (min(need, d.own_len) + W - 1) / W;
  }
  // block j -> global row of its first key, number of in-range keys, local index of the first key (-1: prefix)
  __device__ __forceinline__ void get(const QBlock& d, int j, int& row0, int& valid, int& local0) const {
    if (j < n_pre) {
      row0 = d.pre_row0 + j * W;
      valid = min(W, d.pre_len - j * W);
      local0 = -1;
    } else {
      const int jo = j - n_pre;
      row0 = d.own_row0 + jo * W;
      valid = min(W, d.own_len - jo * W);
      local0 = jo * W;
    }
  }
};

// PROF = true: debug instantiation with per-phase clock64 counters (scripts/prof_attn_phases.py); the production
// instantiation carries none of it
// Forward, built around TMEM (profiles/r2_run22: registers <- TMEM runs at 56 B/clk per SM, so every fp32 tile the softmax
// warps read costs ~1170 cycles per 64 KB):
//   * O stays in TMEM for the whole key loop: P_j.V_j accumulates into it (tcgen05.mma, accumulate flag); the softmax
//     warps never read the partial products back.  The running maximum is LAZY: a row keeps its reference maximum until
//     a block exceeds it by more than 2^8, only then O (TMEM) and l are rescaled — after the first blocks that is rare.
//   * P_j (bf16 pairs) is written with tcgen05.st over the S_j accumulator it was computed from and is the TMEM A
//     operand of P_j.V_j: no P tile in shared memory, no proxy fence.
//   * two issuing threads (Q.K^T on warp 1, P.V on warp 2).
constexpr float FWD_RESCALE_LOG2 = 8.f;
template <bool PROF>
__global__ void __launch_bounds__(384, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], pv_done[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_mx[2][2][BQ];       // [stage][warpgroup][row] partial row max
  __shared__ float s_l[2][BQ];           // [warpgroup][row] partial row sums (final combine)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool prof_on = PROF;
  const long long t_start = prof_on ? clock64() : 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [Q 32K][K0 32K][V0 32K][K1 32K][V1 32K]
  const uint32_t sQ = smem_base, sKV = smem_base + TILE_BYTES;

  const QBlock d = p.qblocks ? p.qblocks[blockIdx.x] : classic_qblock(p);
  const int h = blockIdx.y;
  const int g = h / (p.nq / p.nkv);
  const KeyIter<BKV> kit(d);
  const int n_kb = kit.n_tot;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 256);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();
  // TMEM columns: S0 [0,128) S1 [128,256) O [256,384).  P_j (bf16 pairs) overwrites S_j: keys 64w..64w+63 -> columns
  // [64w, 64w+32) of the stage (the columns warpgroup w read its scores from)
  constexpr uint32_t COL_O = 256;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer: Q, then the K ring (a K slot is free as soon as QK_j retired) ====
    mbar_arrive_expect_tx(&q_full, TILE_BYTES);
    tma_load_rows(smem_gen, &tm, &q_full, h * HD, d.q_row0);
    tma_load_rows(smem_gen + TILE_BYTES / 2, &tm, &q_full, h * HD + 64, d.q_row0);
    const int kcol = (p.nq + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1u);
      uint8_t* k_dst = smem_gen + TILE_BYTES * (1 + 2 * st);
      mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
      tma_load_rows(k_dst, &tm, &k_full[st], kcol, row0);
      tma_load_rows(k_dst + TILE_BYTES / 2, &tm, &k_full[st], kcol + 64, row0);
    }
  } else if (warp == 3 && elect_one_sync()) {
    // ===================== TMA producer: V ring (a V slot is free when P.V_j retired) =====================
    const int vcol = (p.nq + p.nkv + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1u);
      uint8_t* v_dst = smem_gen + TILE_BYTES * (2 + 2 * st);
      mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
      tma_load_rows(v_dst, &tm, &v_full[st], vcol, row0);
      tma_load_rows(v_dst + TILE_BYTES / 2, &tm, &v_full[st], vcol + 64, row0);
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer 1: S_j = Q.K_j^T =====================
    constexpr uint32_t idesc_qk = idesc_128x128(false);
    mbar_wait(&q_full, 0);
    const uint32_t q_lo = desc_lo_sw128(sQ, 16);
    long long m_k = 0, m_s = 0;
    const long long m_t0 = prof_on ? clock64() : 0;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      PROF_T(b0);
      mbar_wait(&k_full[st], (j >> 1) & 1);
      PROF_T(b1);
      mbar_wait(&pv_done[st], ((j >> 1) & 1) ^ 1u);   // P_{j-2}.V_{j-2} has read P_{j-2}, which lives in S stage st
      PROF_T(b2);
      m_k += b1 - b0; m_s += b2 - b1;
      tc_fence_after();
      const uint32_t k_lo = desc_lo_sw128(sKV + TILE_BYTES * (2 * st), 16);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        umma_bf16(tmem_base + st * 128, desc_at(q_lo, off), desc_at(k_lo, off), idesc_qk, kk > 0 ? 1u : 0u);
      }
      umma_commit(&s_full[st]);
      umma_commit(&k_empty[st]);
    }
    if (prof_on) {
      atomicAdd(p.prof + PROF_M_KFULL, (unsigned long long)m_k);
      atomicAdd(p.prof + PROF_M_SEMPTY, (unsigned long long)m_s);
      atomicAdd(p.prof + PROF_M_TOTAL, (unsigned long long)(clock64() - m_t0));
    }
  } else if (warp == 2 && elect_one_sync()) {
    // ===================== MMA issuer 2: O += P_j.V_j (A = P_j in TMEM) =====================
    constexpr uint32_t idesc_pv = idesc_128x128(true);
    long long m_p = 0, m_v = 0;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      PROF_T(a0);
      mbar_wait(&p_full[st], (j >> 1) & 1);   // P_j written; O already rescaled if block j asked for it
      PROF_T(a1);
      mbar_wait(&v_full[st], (j >> 1) & 1);
      PROF_T(a2);
      m_p += a1 - a0; m_v += a2 - a1;
      tc_fence_after();
      // B = V stored [d-half][128 keys][128 B]: MN-major, LBO = 16 KB between the two 64-wide d slabs, SBO = 1 KB
      // between 8-key groups, 2 KB per K=16 step
      const uint32_t v_lo = desc_lo_sw128(sKV + TILE_BYTES * (2 * st + 1), TILE_BYTES / 2);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t a_col = st * 128 + (kk >> 2) * 64 + (kk & 3) * 8;
        umma_bf16_ts(tmem_base + COL_O, tmem_base + a_col, desc_at(v_lo, kk * 2048), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&pv_done[st]);
      umma_commit(&v_empty[st]);
    }
    if (prof_on) {
      atomicAdd(p.prof + PROF_M_PFULL, (unsigned long long)m_p);
      atomicAdd(p.prof + PROF_M_VFULL, (unsigned long long)m_v);
    }
  } else if (warp >= 4) {
    // ===================== softmax / output =====================
    // two warpgroups share every query row: warpgroup wg owns keys [64 wg, 64 wg + 64) of each block and output
    // columns d in [64 wg, 64 wg + 64); TMEM lane = row for both
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int ql = d.q_local0 + r;  // query index inside its segment
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;   // reference maximum (log2 domain, scaled) and row sum relative to it
    long long c_sw = 0, c_ld = 0, c_mx = 0, c_aw = 0, c_ab = 0, c_ex = 0, c_fa = 0;
    const long long t_loop0 = prof_on ? clock64() : 0;

    // validity of this warpgroup's 64 keys of a block (inside the segment AND a real token): every warp builds the two
    // 32-bit words itself with two coalesced loads + ballots, and the loads for block j+1 are issued during block j — no
    // shared-memory exchange, no named barrier, no global-load latency on the per-block critical path
    int mk0 = 0, mk1 = 0;
    auto fetch_mask = [&](int j) {
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      const int k0 = wg * 64 + lane;
      mk0 = (k0 < valid) ? p.key_mask[row0 + k0] : 0;
      mk1 = (k0 + 32 < valid) ? p.key_mask[row0 + k0 + 32] : 0;
    };
    if (n_kb > 0) fetch_mask(0);
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      PROF_T(e0);
      const uint32_t w0 = __ballot_sync(0xffffffffu, mk0 != 0), w1 = __ballot_sync(0xffffffffu, mk1 != 0);
      if (j + 1 < n_kb) fetch_mask(j + 1);
      // causal clipping is needed only when an own key of the block can lie after the tile's first query
      const bool diag = local0 >= 0 && (local0 + BKV - 1 > d.q_local0);
      const bool plain = !diag && (w0 & w1) == 0xFFFFFFFFu;          // no masking at all (the common case)
      PROF_T(e1);
      mbar_wait(&s_full[st], (j >> 1) & 1);
      PROF_T(e2);
      tc_fence_after();
      const uint32_t ts = tmem_base + st * 128 + lane_addr + wg * 64;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(ts, v0);
      tmem_ld_32x32(ts + 32, v1);
      tmem_ld_wait();
      PROF_T(e3);
      c_mx += e1 - e0; c_sw += e2 - e1; c_ld += e3 - e2;
      const int kbase = local0 + wg * 64;  // local index of this warpgroup's first key (own blocks)
      // ---- row max over the own 64 keys, then exchange with the other warpgroup ----
      // masked scores become -inf in place (UNSCALED: scale > 0 keeps -inf), so that the exp loop below is the same
      // straight-line code for plain and masked blocks
      const float sl2 = p.scale_log2;
      float mx = -INFINITY;
      if (!plain) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool ok0 = ((w0 >> i) & 1u) && (!diag || kbase + i <= ql);
          const bool ok1 = ((w1 >> i) & 1u) && (!diag || kbase + 32 + i <= ql);
          if (!ok0) v0[i] = 0xFF800000u;   // -inf
          if (!ok1) v1[i] = 0xFF800000u;
        }
      }
      {
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;   // scale > 0; -inf stays -inf
      }
      s_mx[st][wg][r] = mx;
      named_bar_sync(2, 256);
      mx = fmaxf(mx, s_mx[st][wg ^ 1][r]);
      PROF_T(e4);
      c_mx += e4 - e3;
      // ---- lazy reference maximum: both warpgroups see the same mx and m_ref, so they take the same decision ----
      const bool bump = mx > m_ref + FWD_RESCALE_LOG2 || (m_ref == -INFINITY && mx > -INFINITY);
      if (__any_sync(0xffffffffu, bump)) {
        const float m_new = bump ? mx : m_ref;
        const float corr = (bump && m_ref > -INFINITY) ? ex2_approx(m_ref - m_new) : 1.f;
        if (j > 0) {
          // O holds sum_{i<j} P_i.V_i relative to m_ref: rescale this warp's 32 rows x own 64 columns in place once
          // P_{j-1}.V_{j-1} (the last MMA that writes O) has retired
          PROF_T(w0c);
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          PROF_T(w1c);
          c_aw += w1c - w0c;
          tc_fence_after();
          const uint32_t to = tmem_base + COL_O + lane_addr + wg * 64;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t a[32];
            tmem_ld_32x32(to + half * 32, a);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = __float_as_uint(__uint_as_float(a[i]) * corr);
            tmem_st_32x32(to + half * 32, a);
          }
          if (prof_on) c_ab += clock64() - w1c;
        }
        l_run *= corr;
        m_ref = m_new;
      }
      const float mu = (m_ref == -INFINITY) ? 0.f : m_ref;
      PROF_T(e5);
      // ---- p = exp2(s - m_ref) -> bf16 pairs over this warpgroup's score columns: A operand of P.V ----
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float s0 = __uint_as_float(half == 0 ? v0[2 * i] : v1[2 * i]);
          const float s1 = __uint_as_float(half == 0 ? v0[2 * i + 1] : v1[2 * i + 1]);
          const float p0 = ex2_approx(fmaf(s0, sl2, -mu));   // masked: -inf * scale - mu = -inf -> 2^-inf = 0
          const float p1 = ex2_approx(fmaf(s1, sl2, -mu));
          rs4[(2 * i) & 3] += p0;
          rs4[(2 * i + 1) & 3] += p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x16(ts + half * 16, pk);
      }
      l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
      PROF_T(e6);
      tmem_st_wait();      // P_j and (if any) the rescaled O are in TMEM
      tc_fence_before();
      mbar_arrive(&p_full[st]);
      if (prof_on) { c_ex += e6 - e5; c_fa += clock64() - e6; }
    }
    const long long t_loop1 = prof_on ? clock64() : 0;
    s_l[wg][r] = l_run;
    named_bar_sync(3, 256);
    const float l_tot = l_run + s_l[wg ^ 1][r];
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    if (n_kb > 0) mbar_wait(&pv_done[(n_kb - 1) & 1], ((n_kb - 1) >> 1) & 1);
    tc_fence_after();
    {
      // the TMEM loads are .sync.aligned: every lane executes them, only the stores are predicated
      const bool row_ok = r < d.q_rows;
      if (row_ok && wg == 0)
        p.lse2[(long long)d.stat0 + (long long)h * p.stat_h + r] = l_tot > 0.f ? m_ref + log2f(l_tot) : INFINITY;
      bf16* dst = p.out + (long long)(d.q_row0 + (row_ok ? r : 0)) * p.nq * HD + h * HD + wg * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t a[32];
        tmem_ld_32x32(tmem_base + COL_O + lane_addr + wg * 64 + half * 32, a);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = n_kb > 0 ? __uint_as_float(a[i + u]) * inv : 0.f;
            *reinterpret_cast<bf16x8*>(dst + half * 32 + i) = pack8(f);
          }
        }
      }
    }
    if (prof_on && (threadIdx.x == 128 || threadIdx.x == 256)) {   // one thread of each softmax warpgroup
      atomicAdd(p.prof + PROF_S_WAIT, (unsigned long long)c_sw);
      atomicAdd(p.prof + PROF_S_LD, (unsigned long long)c_ld);
      atomicAdd(p.prof + PROF_MAX_XCHG, (unsigned long long)c_mx);
      atomicAdd(p.prof + PROF_ABSORB_WAIT, (unsigned long long)c_aw);
      atomicAdd(p.prof + PROF_ABSORB, (unsigned long long)c_ab);
      atomicAdd(p.prof + PROF_EXP_STORE, (unsigned long long)c_ex);
      atomicAdd(p.prof + PROF_FENCE_ARRIVE, (unsigned long long)c_fa);
      atomicAdd(p.prof + PROF_LOOP, (unsigned long long)(t_loop1 - t_loop0));
      atomicAdd(p.prof + PROF_PROLOGUE, (unsigned long long)(t_loop0 - t_start));
      atomicAdd(p.prof + PROF_EPILOGUE, (unsigned long long)(clock64() - t_loop1));
      if (threadIdx.x == 128) {
        atomicAdd(p.prof + PROF_BLOCKS, (unsigned long long)n_kb);
        atomicAdd(p.prof + PROF_CTAS, 1ull);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==================================================================================================
// BACKWARD on tcgen05.  Deterministic two-kernel split like the mma.sync version:
//   dQ  kernel: CTA = (128-query block, q head, sequence), inner loop over 64-key blocks j up to the diagonal:
//       S_j = Q.K_j^T, dP_j = dO.V_j^T (TMEM, double buffered)  ->  softmax warps: P = exp2(S*c - lse2),
//       dS = scale * P * (dP - delta) written (bf16) over S_j in TMEM  ->  dQ += dS_j.K_j (TMEM accumulator).
//   dKV kernel: CTA = (128-key block, kv head, sequence), inner loop over (q head of the GQA group, 64-query
//       block at or after the key block): S^T = K.Q^T, dP^T = V.dO^T (TMEM, double buffered) -> P^T, dS^T written
//       over them -> dV += P^T.dO, dK += dS^T.Q (TMEM accumulators, summed over the whole group: no atomics).
// The 64-row streamed tiles are stored [d-half][64 rows][128 B]; the SAME smem tile is read K-major (rows = N)
// by the score MMAs and MN-major (rows = K) by the gradient MMAs.
// ==================================================================================================
// ---------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------
// dQ: dS_j never touches shared memory — it is written (bf16 pairs, tcgen05.st) over the S_j accumulator it was computed
// from and is the TMEM A operand of dQ += dS_j.K_j (tcgen05.mma [d], [a], b-desc).  (profiles/r2_run14: with dS staged
// in shared memory the 128 B/clk shared-memory port carried 176 KB per 64-key block and was the floor.)
constexpr int DQ_KV_STAGES = 4;
template <bool PROF>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_do128,
                      const __grid_constant__ CUtensorMap tm_kv64, const AttnParams p) {
  constexpr bool prof_on = PROF;
  const long long t_start = prof_on ? clock64() : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t qdo_full, kv_full[DQ_KV_STAGES], kv_empty[DQ_KV_STAGES], sp_full[2], ds_full[2], ds_empty[2], dq_full;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [Q 32K][dO 32K][DQ_KV_STAGES x (K 16K, V 16K)] = 192 KB
  const uint32_t sQ = smem_base, sdO = sQ + TILE_BYTES, sKV = sdO + TILE_BYTES;
  const QBlock d = p.qblocks ? p.qblocks[blockIdx.x] : classic_qblock(p);
  const int h = blockIdx.y;
  const int g = h / (p.nq / p.nkv);
  const KeyIter<64> kit(d);
  const int n_kb = kit.n_tot;

  if (threadIdx.x == 0) {
    mbar_init(&qdo_full, 1);
    mbar_init(&dq_full, 1);
    for (int s = 0; s < DQ_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sp_full[s], 2);
      mbar_init(&ds_full[s], 256);
      mbar_init(&ds_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();
  // TMEM columns: S0 [0,64) S1 [64,128) dP0 [128,192) dP1 [192,256) dQ [256,384)
  // dS_j (bf16 pairs, column i = keys 2i, 2i+1) overwrites S_j: keys 32c..32c+31 -> columns [32c, 32c+16) of the stage.
  // Q and dO stay in shared memory (TMA): with dS in TMEM the shared-memory port carries 144 KB per key block, below
  // what the TMEM read port (56 B/clk, profiles/r2_run22) allows the dS warps anyway, and a Q / dO copy into TMEM
  // through registers cost 5000 cycles of prologue per CTA (profiles/r2_run18).
  constexpr uint32_t COL_DP = 128, COL_DQ = 256;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer: Q, dO once, then K_j / V_j =====================
    mbar_arrive_expect_tx(&qdo_full, 2 * TILE_BYTES);
    tma_load_rows(smem_gen, &tm_q128, &qdo_full, h * HD, d.q_row0);
    tma_load_rows(smem_gen + TILE_BYTES / 2, &tm_q128, &qdo_full, h * HD + 64, d.q_row0);
    tma_load_rows(smem_gen + TILE_BYTES, &tm_do128, &qdo_full, h * HD, d.q_row0);
    tma_load_rows(smem_gen + TILE_BYTES + TILE_BYTES / 2, &tm_do128, &qdo_full, h * HD + 64, d.q_row0);
    const int kcol = (p.nq + g) * HD, vcol = (p.nq + p.nkv + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j % DQ_KV_STAGES;
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      mbar_wait(&kv_empty[st], ((j / DQ_KV_STAGES) & 1) ^ 1u);
      uint8_t* kd = smem_gen + 2 * TILE_BYTES + st * 2 * HALF_TILE;
      uint8_t* vd = kd + HALF_TILE;
      mbar_arrive_expect_tx(&kv_full[st], 2 * HALF_TILE);
      tma_load_rows(kd, &tm_kv64, &kv_full[st], kcol, row0);
      tma_load_rows(kd + HALF_TILE / 2, &tm_kv64, &kv_full[st], kcol + 64, row0);
      tma_load_rows(vd, &tm_kv64, &kv_full[st], vcol, row0);
      tma_load_rows(vd + HALF_TILE / 2, &tm_kv64, &kv_full[st], vcol + 64, row0);
    }
  } else if ((warp == 1 || warp == 2) && elect_one_sync()) {
    // ===================== MMA issuers 1 / 2: S_j = Q.K_j^T (warp 1), dP_j = dO.V_j^T (warp 2) =====================
    // Three issuing threads on three SM sub-partitions (profiles/r2_run11, r2_run17: one thread needs 80-100 cycles per
    // tcgen05.mma while an N = 64 MMA executes in < 50, so the issuer, not the tensor pipe, paced the kernel).  The
    // tensor pipe runs the MMAs in arrival order; the order that matters is carried by the mbarriers.
    constexpr uint32_t id_s = idesc_n(64, false);    // [128 q] x [64 keys], B K-major over d
    const bool is_dp = warp == 2;
    const uint32_t a_lo = desc_lo_sw128(is_dp ? sdO : sQ, 16);
    const uint32_t d_col = tmem_base + (is_dp ? COL_DP : 0u);
    mbar_wait(&qdo_full, 0);
    tc_fence_after();
    long long m_k = 0, m_s = 0;
    const long long m_t0 = prof_on ? clock64() : 0;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1, k3 = j % DQ_KV_STAGES;
      PROF_T(b0);
      mbar_wait(&kv_full[k3], (j / DQ_KV_STAGES) & 1);
      PROF_T(b1);
      mbar_wait(&ds_empty[st], ((j >> 1) & 1) ^ 1u);   // dQ_{j-2} has consumed dS_{j-2}, which lives in S stage st
      if (prof_on) { m_k += b1 - b0; m_s += clock64() - b1; }
      tc_fence_after();
      const uint32_t b_lo = desc_lo_sw128(sKV + k3 * 2 * HALF_TILE + (is_dp ? HALF_TILE : 0), 16);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {  // d = 128 = 8 x K16, two d-halves
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(d_col + st * 64, desc_at(a_lo, aoff), desc_at(b_lo, boff), id_s, kk > 0 ? 1u : 0u);
      }
      umma_commit(&sp_full[st]);   // count 2: S_j and dP_j
    }
    if (prof_on && !is_dp) {
      atomicAdd(p.prof + PROF_M_KFULL, (unsigned long long)m_k);
      atomicAdd(p.prof + PROF_M_SEMPTY, (unsigned long long)m_s);
      atomicAdd(p.prof + PROF_M_TOTAL, (unsigned long long)(clock64() - m_t0));
    }
  } else if (warp == 3 && elect_one_sync()) {
    // ===================== MMA issuer 3: dQ += dS_j.K_j =====================
    constexpr uint32_t id_dq = idesc_n(128, true);   // [128 q] x [128 d], B = K_j MN-major (k = keys)
    long long m_d = 0;
    const long long m_t0 = prof_on ? clock64() : 0;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1, k3 = j % DQ_KV_STAGES;
      PROF_T(a0);
      mbar_wait(&kv_full[k3], (j / DQ_KV_STAGES) & 1);  // K_j landed (long ago: dS_j was computed from it); orders this thread's reads
      mbar_wait(&ds_full[st], (j >> 1) & 1);
      if (prof_on) m_d += clock64() - a0;
      tc_fence_after();
      const uint32_t b_lo = desc_lo_sw128(sKV + k3 * 2 * HALF_TILE, HALF_TILE / 2);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 64 keys = 4 x K16; keys 32c.. sit at columns 32c.. of the S stage
        const uint32_t a_col = st * 64 + (kk >> 1) * 32 + (kk & 1) * 8;
        umma_bf16_ts(tmem_base + COL_DQ, tmem_base + a_col, desc_at(b_lo, kk * 2048), id_dq, (j > 0 || kk > 0) ? 1u : 0u);
      }
      // S_j / dP_j completed before dS_j existed, so K_j / V_j have no reader left once dQ_j is done
      umma_commit(&kv_empty[k3]);
      umma_commit(&ds_empty[st]);   // count 1; both S / dP issuers wait on it
    }
    umma_commit(&dq_full);
    if (prof_on) {
      atomicAdd(p.prof + PROF_M_PFULL, (unsigned long long)m_d);
      atomicAdd(p.prof + PROF_M_VFULL, (unsigned long long)(clock64() - m_t0));
    }
  } else if (warp >= 4) {
    // ===================== dS producer: two warpgroups, each 32 of the 64 keys of a block =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int ql = d.q_local0 + r;
    const bool q_ok = r < d.q_rows;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long sidx = (long long)d.stat0 + (long long)h * p.stat_h + r;
    const float lse = q_ok ? p.lse2[sidx] : INFINITY;
    const float del = q_ok ? p.delta[sidx] : 0.f;
    // validity word of this warpgroup's 32-key chunk: built by every warp itself (one coalesced load + ballot), the load
    // for block j+1 issued during block j — no shared-memory exchange, no named barrier, no load latency per block
    const int c = wg;
    int mk = 0;
    auto fetch_mask = [&](int j) {
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      const int k0 = c * 32 + lane;
      mk = (k0 < valid) ? p.key_mask[row0 + k0] : 0;
    };
    if (n_kb > 0) fetch_mask(0);
    long long c_sw = 0, c_ld = 0, c_ma = 0, c_fa = 0, c_mk = 0;
    const long long t_loop0 = prof_on ? clock64() : 0;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      int row0, valid, local0;
      kit.get(d, j, row0, valid, local0);
      PROF_T(e0);
      const uint32_t mw = __ballot_sync(0xffffffffu, mk != 0);
      if (j + 1 < n_kb) fetch_mask(j + 1);
      const int kc0 = local0 + c * 32;  // local index of the chunk's first key (own blocks)
      // causal clipping only when an own key of the chunk can lie after the tile's first query
      const bool diag = local0 >= 0 && (kc0 + 31 > d.q_local0);
      const bool plain = !diag && mw == 0xFFFFFFFFu;
      PROF_T(e1);
      mbar_wait(&sp_full[st], (j >> 1) & 1);
      PROF_T(e2);
      tc_fence_after();
      {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem_base + st * 64 + lane_addr + c * 32, sv);
        tmem_ld_32x32(tmem_base + COL_DP + st * 64 + lane_addr + c * 32, dv);
        tmem_ld_wait();
        PROF_T(e4);
        c_mk += e1 - e0; c_sw += e2 - e1; c_ld += e4 - e2;
        float f[32];
        if (plain) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float pr = ex2_approx(__uint_as_float(sv[i]) * p.scale_log2 - lse);
            f[i] = p.scale * pr * (__uint_as_float(dv[i]) - del);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool ok = ((mw >> i) & 1u) && (!diag || kc0 + i <= ql);
            const float pr = ok ? ex2_approx(__uint_as_float(sv[i]) * p.scale_log2 - lse) : 0.f;
            f[i] = p.scale * pr * (__uint_as_float(dv[i]) - del);
          }
        }
        // dS (bf16 pairs) back into the columns this thread just read S from: A operand of dQ += dS.K
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
        tmem_st_32x16(tmem_base + st * 64 + lane_addr + c * 32, pk);
        if (prof_on) c_ma += clock64() - e4;
      }
      PROF_T(e5);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[st]);
      if (prof_on) c_fa += clock64() - e5;
    }
    const long long t_loop1 = prof_on ? clock64() : 0;
    // ---- write dQ ----
    mbar_wait(&dq_full, 0);
    tc_fence_after();
    {
      // the TMEM loads are .sync.aligned: every lane executes them, only the stores are predicated
      bf16* dst = p.dqkv + (long long)(d.q_row0 + (q_ok ? r : 0)) * (long long)(p.nq + 2 * p.nkv) * HD + h * HD;
#pragma unroll
      for (int c = 2 * wg; c < 2 * wg + 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + COL_DQ + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[u * 8 + i]);
            *reinterpret_cast<bf16x8*>(dst + c * 32 + u * 8) = pack8(f);
          }
        }
      }
    }
    if (prof_on && (threadIdx.x == 128 || threadIdx.x == 256)) {
      atomicAdd(p.prof + PROF_S_WAIT, (unsigned long long)c_sw);
      atomicAdd(p.prof + PROF_S_LD, (unsigned long long)c_ld);
      atomicAdd(p.prof + PROF_MAX_XCHG, (unsigned long long)c_mk);
      atomicAdd(p.prof + PROF_EXP_STORE, (unsigned long long)c_ma);
      atomicAdd(p.prof + PROF_FENCE_ARRIVE, (unsigned long long)c_fa);
      atomicAdd(p.prof + PROF_LOOP, (unsigned long long)(t_loop1 - t_loop0));
      atomicAdd(p.prof + PROF_PROLOGUE, (unsigned long long)(t_loop0 - t_start));
      atomicAdd(p.prof + PROF_EPILOGUE, (unsigned long long)(clock64() - t_loop1));
      if (threadIdx.x == 128) {
        atomicAdd(p.prof + PROF_BLOCKS, (unsigned long long)n_kb);
        atomicAdd(p.prof + PROF_CTAS, 1ull);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// dK, dV
// ---------------------------------------------------------------------------------------------------
// dK / dV: P^T_j and dS^T_j never touch shared memory — they are written (bf16 pairs, tcgen05.st) over the S^T_j / dP^T_j
// accumulators they were computed from and feed dV += P^T.dO, dK += dS^T.Q as TMEM A operands.
constexpr int DKV_QD_STAGES = 4;
template <bool PROF>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_kv128, const __grid_constant__ CUtensorMap tm_q64,
                       const __grid_constant__ CUtensorMap tm_do64, const AttnParams p) {
  constexpr bool prof_on = PROF;
  const long long t_start = prof_on ? clock64() : 0;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, qd_full[DKV_QD_STAGES], qd_empty[DKV_QD_STAGES], sp_full[2], ds_full[2], ds_empty[2], out_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_lse[2][64], s_del[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [K 32K][V 32K][DKV_QD_STAGES x (Q 16K, dO 16K)] = 192 KB
  const uint32_t sK = smem_base, sV = sK + TILE_BYTES, sQD = sV + TILE_BYTES;
  KBlock d;
  if (p.kblocks) {
    d = p.kblocks[blockIdx.x];
  } else {  // classic: grid = (ceil(L/128), nkv, B); key block 0 (most work) first
    const int kb = blockIdx.x, b = blockIdx.z;
    d.k_row0 = b * p.L + kb * BQ;
    d.k_rows = min(BQ, p.L - kb * BQ);
    d.k_local0 = kb * BQ;
    d.q_row0 = b * p.L;
    d.q_len = p.L;
    d.causal = 1;
    d.stat0 = b * p.nq * p.L;
    d.out_row0 = 0;
  }
  const int g = blockIdx.y;
  const int group = p.nq / p.nkv;
  const int qb0 = d.causal ? d.k_local0 / 64 : 0;   // first 64-query block that can see these keys
  const int nqb = (d.q_len + 63) / 64 - qb0;        // 64-query blocks per head
  const int n_it = group * nqb;

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    mbar_init(&out_full, 1);
    for (int s = 0; s < DKV_QD_STAGES; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sp_full[s], 2);
      mbar_init(&ds_full[s], 256);
      mbar_init(&ds_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();
  // TMEM columns: S^T0 [0,64) S^T1 [64,128) dP^T0 [128,192) dP^T1 [192,256) dK [256,384) dV [384,512)
  // P^T_j / dS^T_j (bf16 pairs) overwrite S^T_j / dP^T_j: queries 32c..32c+31 -> columns [32c, 32c+16) of the stage

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    const int kcol = (p.nq + g) * HD, vcol = (p.nq + p.nkv + g) * HD;
    mbar_arrive_expect_tx(&kv_full, 2 * TILE_BYTES);
    tma_load_rows(smem_gen, &tm_kv128, &kv_full, kcol, d.k_row0);
    tma_load_rows(smem_gen + TILE_BYTES / 2, &tm_kv128, &kv_full, kcol + 64, d.k_row0);
    tma_load_rows(smem_gen + TILE_BYTES, &tm_kv128, &kv_full, vcol, d.k_row0);
    tma_load_rows(smem_gen + TILE_BYTES + TILE_BYTES / 2, &tm_kv128, &kv_full, vcol + 64, d.k_row0);
    for (int it = 0; it < n_it; ++it) {
      const int st = it % DKV_QD_STAGES;
      const int h = g * group + it / nqb;
      const int qrow = d.q_row0 + (qb0 + it % nqb) * 64;
      mbar_wait(&qd_empty[st], ((it / DKV_QD_STAGES) & 1) ^ 1u);
      uint8_t* qd = smem_gen + 2 * TILE_BYTES + st * 2 * HALF_TILE;
      uint8_t* dd = qd + HALF_TILE;
      mbar_arrive_expect_tx(&qd_full[st], 2 * HALF_TILE);
      tma_load_rows(qd, &tm_q64, &qd_full[st], h * HD, qrow);
      tma_load_rows(qd + HALF_TILE / 2, &tm_q64, &qd_full[st], h * HD + 64, qrow);
      tma_load_rows(dd, &tm_do64, &qd_full[st], h * HD, qrow);
      tma_load_rows(dd + HALF_TILE / 2, &tm_do64, &qd_full[st], h * HD + 64, qrow);
    }
  } else if ((warp == 1 || warp == 2) && elect_one_sync()) {
    // ===================== MMA issuers 1 / 2: S^T = K.Q^T (warp 1), dP^T = V.dO^T (warp 2); see the dQ kernel ==========
    constexpr uint32_t id_s = idesc_n(64, false);    // [128 keys] x [64 queries], K-major over d
    const bool is_dp = warp == 2;
    const uint32_t a_lo = desc_lo_sw128(is_dp ? sV : sK, 16);
    const uint32_t d_col = tmem_base + (is_dp ? 128u : 0u);
    mbar_wait(&kv_full, 0);
    long long m_k = 0, m_s = 0;
    const long long m_t0 = prof_on ? clock64() : 0;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1, q3 = it % DKV_QD_STAGES;
      PROF_T(b0);
      mbar_wait(&qd_full[q3], (it / DKV_QD_STAGES) & 1);
      PROF_T(b1);
      mbar_wait(&ds_empty[st], ((it >> 1) & 1) ^ 1u);   // the gradient MMAs of block it-2 have consumed P^T / dS^T of this stage
      if (prof_on) { m_k += b1 - b0; m_s += clock64() - b1; }
      tc_fence_after();
      const uint32_t b_lo = desc_lo_sw128(sQD + q3 * 2 * HALF_TILE + (is_dp ? HALF_TILE : 0), 16);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(d_col + st * 64, desc_at(a_lo, aoff), desc_at(b_lo, boff), id_s, kk > 0 ? 1u : 0u);
      }
      umma_commit(&sp_full[st]);   // count 2: S^T and dP^T
    }
    if (prof_on && !is_dp) {
      atomicAdd(p.prof + PROF_M_KFULL, (unsigned long long)m_k);
      atomicAdd(p.prof + PROF_M_SEMPTY, (unsigned long long)m_s);
      atomicAdd(p.prof + PROF_M_TOTAL, (unsigned long long)(clock64() - m_t0));
    }
  } else if (warp == 3 && elect_one_sync()) {
    // ===================== MMA issuer 3: dV += P^T.dO and dK += dS^T.Q =====================
    constexpr uint32_t id_g = idesc_n(128, true);    // [128 keys] x [128 d], B = dO / Q MN-major (k = queries)
    long long m_d = 0;
    const long long m_t0 = prof_on ? clock64() : 0;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1, q3 = it % DKV_QD_STAGES;
      PROF_T(a0);
      mbar_wait(&qd_full[q3], (it / DKV_QD_STAGES) & 1);  // landed long ago (P^T was computed from it); orders this thread's reads
      mbar_wait(&ds_full[st], (it >> 1) & 1);
      if (prof_on) m_d += clock64() - a0;
      tc_fence_after();
      const uint32_t q_lo = desc_lo_sw128(sQD + q3 * 2 * HALF_TILE, HALF_TILE / 2);
      const uint32_t d_lo = desc_lo_sw128(sQD + q3 * 2 * HALF_TILE + HALF_TILE, HALF_TILE / 2);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 64 queries = 4 x K16; queries 32c.. sit at columns 32c.. of the stage
        const uint32_t a_col = st * 64 + (kk >> 1) * 32 + (kk & 1) * 8;
        umma_bf16_ts(tmem_base + 384, tmem_base + a_col, desc_at(d_lo, kk * 2048), id_g, (it > 0 || kk > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t a_col = 128 + st * 64 + (kk >> 1) * 32 + (kk & 1) * 8;
        umma_bf16_ts(tmem_base + 256, tmem_base + a_col, desc_at(q_lo, kk * 2048), id_g, (it > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&qd_empty[q3]);
      umma_commit(&ds_empty[st]);
    }
    umma_commit(&out_full);
    if (prof_on) {
      atomicAdd(p.prof + PROF_M_PFULL, (unsigned long long)m_d);
      atomicAdd(p.prof + PROF_M_VFULL, (unsigned long long)(clock64() - m_t0));
    }
  } else if (warp >= 4) {
    // ===================== P^T / dS^T producer: one KEY row per thread, two warpgroups x 32 queries =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int t = threadIdx.x - 128;
    const int kl = d.k_local0 + r;  // local key index (meaningful when causal)
    const bool row_ok = r < d.k_rows;
    const bool key_ok = row_ok && p.key_mask[d.k_row0 + (row_ok ? r : 0)] != 0;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    // softmax statistics of the 64 queries of an inner block: warps 4-5 load them one iteration AHEAD into registers and
    // only publish them to shared memory here, so the global-load latency is off the per-iteration critical path
    float lse_n = INFINITY, del_n = 0.f;
    auto fetch_stats = [&](int it) {
      if (t < 64) {
        const int h = g * group + it / nqb;
        const int qi = (qb0 + it % nqb) * 64 + t;
        const long long sidx = (long long)d.stat0 + (long long)h * p.stat_h + qi;
        lse_n = qi < d.q_len ? p.lse2[sidx] : INFINITY;   // +inf -> probability 0 for queries past the end
        del_n = qi < d.q_len ? p.delta[sidx] : 0.f;
      }
    };
    if (n_it > 0) fetch_stats(0);
    long long c_sw = 0, c_ld = 0, c_ma = 0, c_fa = 0, c_mk = 0;
    const long long t_loop0 = prof_on ? clock64() : 0;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int qs = (qb0 + it % nqb) * 64;  // local index of the block's first query
      PROF_T(e0);
      if (t < 64) {
        s_lse[st][t] = lse_n;
        s_del[st][t] = del_n;
      }
      if (it + 1 < n_it) fetch_stats(it + 1);
      named_bar_sync(1, 256);
      const int c = wg;
      const int qc0 = qs + c * 32;
      const bool need_cmp = d.causal && (qc0 < d.k_local0 + BQ - 1);  // some (key, query) pair may violate key <= query
      PROF_T(e1);
      mbar_wait(&sp_full[st], (it >> 1) & 1);
      PROF_T(e2);
      tc_fence_after();
      {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem_base + st * 64 + lane_addr + c * 32, sv);
        tmem_ld_32x32(tmem_base + 128 + st * 64 + lane_addr + c * 32, dv);
        tmem_ld_wait();
        PROF_T(e4);
        c_mk += e1 - e0; c_sw += e2 - e1; c_ld += e4 - e2;
        uint32_t pk_p[16], pk_s[16];
        const float4* l4 = reinterpret_cast<const float4*>(&s_lse[st][c * 32]);
        const float4* d4 = reinterpret_cast<const float4*>(&s_del[st][c * 32]);
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const float4 lv = l4[v4], dl = d4[v4];
          const float ls[4] = {lv.x, lv.y, lv.z, lv.w}, ds4[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
          float fp[4], fs[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = v4 * 4 + u;
            float pr = ex2_approx(__uint_as_float(sv[i]) * p.scale_log2 - ls[u]);
            if (!key_ok || (need_cmp && kl > qc0 + i)) pr = 0.f;  // select, never 0 * inf
            fp[u] = pr;
            fs[u] = p.scale * pr * (__uint_as_float(dv[i]) - ds4[u]);
          }
          pk_p[2 * v4] = pack_bf16x2(fp[0], fp[1]);
          pk_p[2 * v4 + 1] = pack_bf16x2(fp[2], fp[3]);
          pk_s[2 * v4] = pack_bf16x2(fs[0], fs[1]);
          pk_s[2 * v4 + 1] = pack_bf16x2(fs[2], fs[3]);
        }
        // P^T / dS^T back into the columns this thread just read: A operands of dV += P^T.dO and dK += dS^T.Q
        tmem_st_32x16(tmem_base + st * 64 + lane_addr + c * 32, pk_p);
        tmem_st_32x16(tmem_base + 128 + st * 64 + lane_addr + c * 32, pk_s);
        if (prof_on) c_ma += clock64() - e4;
      }
      PROF_T(e5);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&ds_full[st]);
      if (prof_on) c_fa += clock64() - e5;
    }
    const long long t_loop1 = prof_on ? clock64() : 0;
    // ---- write dK (warpgroup 0) / dV (warpgroup 1): bf16 into dqkv (classic) or an fp32 partial slab (packed) ----
    mbar_wait(&out_full, 0);
    tc_fence_after();
    const int which = wg;
    const long long stride = (long long)(p.nq + 2 * p.nkv) * HD;
    bf16* dst16 = p.dqkv + (long long)(d.k_row0 + (row_ok ? r : 0)) * stride +
                  (which == 0 ? (p.nq + g) : (p.nq + p.nkv + g)) * HD;
    float* dst32 = p.kv_part ? p.kv_part + ((long long)(d.out_row0 + (row_ok ? r : 0)) * 2 * p.nkv + which * p.nkv + g) * HD
                             : nullptr;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + 256 + which * 128 + lane_addr + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = n_it > 0 ? __uint_as_float(v[u * 8 + i]) : 0.f;
          if (dst32) {
            *reinterpret_cast<float4*>(dst32 + c * 32 + u * 8) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(dst32 + c * 32 + u * 8 + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            *reinterpret_cast<bf16x8*>(dst16 + c * 32 + u * 8) = pack8(f);
          }
        }
      }
    }
    if (prof_on && (threadIdx.x == 128 || threadIdx.x == 256)) {
      atomicAdd(p.prof + PROF_S_WAIT, (unsigned long long)c_sw);
      atomicAdd(p.prof + PROF_S_LD, (unsigned long long)c_ld);
      atomicAdd(p.prof + PROF_MAX_XCHG, (unsigned long long)c_mk);
      atomicAdd(p.prof + PROF_EXP_STORE, (unsigned long long)c_ma);
      atomicAdd(p.prof + PROF_FENCE_ARRIVE, (unsigned long long)c_fa);
      atomicAdd(p.prof + PROF_LOOP, (unsigned long long)(t_loop1 - t_loop0));
      atomicAdd(p.prof + PROF_PROLOGUE, (unsigned long long)(t_loop0 - t_start));
      atomicAdd(p.prof + PROF_EPILOGUE, (unsigned long long)(clock64() - t_loop1));
      if (threadIdx.x == 128) {
        atomicAdd(p.prof + PROF_BLOCKS, (unsigned long long)n_it);
        atomicAdd(p.prof + PROF_CTAS, 1ull);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// packed layout: dK/dV of key row m = sum (fixed order) of its fp32 partial rows -> bf16 into dqkv
__global__ void kv_reduce_kernel(const float* __restrict__ part, const int* __restrict__ row_start,
                                 const int* __restrict__ row_list, bf16* __restrict__ dqkv, int nq, int nkv) {
  pdl_enter();
  const int m = blockIdx.x;
  const int s0 = row_start[m], s1 = row_start[m + 1];
  const int width = 2 * nkv * HD;
  const long long stride = (long long)(nq + 2 * nkv) * HD;
  for (int c = threadIdx.x * 8; c < width; c += blockDim.x * 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = s0; s < s1; ++s) {
      const float* src = part + (long long)row_list[s] * width + c;
      const float4 a = *reinterpret_cast<const float4*>(src), b2 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b2.x; acc[5] += b2.y; acc[6] += b2.z; acc[7] += b2.w;
    }
    // partial row layout [which][g][HD] == dqkv columns (nq + which*nkv + g)*HD + d
    *reinterpret_cast<bf16x8*>(dqkv + (long long)m * stride + (long long)nq * HD + c) = pack8(acc);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D map over x[rows][cols] bf16, box = 64 cols x box_rows rows, 128B swizzle (rows past the end are zero-filled;
// rows of a neighbouring segment that fall into a box are masked by index in the kernels)
int make_rows_map(CUtensorMap* tm, const void* base, long long rows, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
  return 0;
}

constexpr int SMEM_FWD = 5 * TILE_BYTES + 1024;
constexpr int SMEM_DQ = 2 * TILE_BYTES + DQ_KV_STAGES * 2 * HALF_TILE + 1024;
constexpr int SMEM_DKV = 2 * TILE_BYTES + DKV_QD_STAGES * 2 * HALF_TILE + 1024;

int set_attrs() {
  static DeviceOnce once;   // kernel attributes are per device
  if (!once.first()) return 0;
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FWD));
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
  B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
  return 0;
}

unsigned long long* g_attn_prof = nullptr;   // device buffer of PROF_N counters (debug; scripts/prof_attn_phases.py)
int g_attn_prof_which = 0;                   // which kernel writes them: 0 forward, 1 dQ, 2 dK/dV

AttnParams base_params(const int* key_mask, int nq, int nkv, float scale) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.key_mask = key_mask;
  p.nq = nq;
  p.nkv = nkv;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.prof = g_attn_prof;
  return p;
}

struct BwdMaps {
  CUtensorMap q128, q64, d128, d64;
};
int make_bwd_maps(BwdMaps& m, const void* qkv, const void* dout, long long rows, int nq, int nkv) {
  const long long qcols = (long long)(nq + 2 * nkv) * HD, ocols = (long long)nq * HD;
  int rc;
  if ((rc = make_rows_map(&m.q128, qkv, rows, qcols, 128))) return rc;
  if ((rc = make_rows_map(&m.q64, qkv, rows, qcols, 64))) return rc;
  if ((rc = make_rows_map(&m.d128, dout, rows, ocols, 128))) return rc;
  if ((rc = make_rows_map(&m.d64, dout, rows, ocols, 64))) return rc;
  return set_attrs();
}

}  // namespace

// ---- classic layout -----------------------------------------------------------------------------------
int attn_fwd_tc_launch(const void* qkv, const int* key_mask, void* out, float* lse, int B, int L, int nq, int nkv,
                       float scale, cudaStream_t stream) {
  CUtensorMap tm;
  int rc = make_rows_map(&tm, qkv, (long long)B * L, (long long)(nq + 2 * nkv) * HD, 128);
  if (rc) return rc;
  if ((rc = set_attrs())) return rc;
  AttnParams p = base_params(key_mask, nq, nkv, scale);
  p.out = (bf16*)out;
  p.lse2 = lse;
  p.L = L;
  p.stat_h = L;
  dim3 grid((L + BQ - 1) / BQ, nq, B);
  B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 0 ? attn_fwd_tc_kernel<true> : attn_fwd_tc_kernel<false>, dim3(grid), dim3(384), SMEM_FWD, stream, tm, p));
  B200RL_LAUNCH_OK();
  return 0;
}

int attn_bwd_tc_launch(const void* qkv, const int* key_mask, const void* dout, const float* lse, const float* delta,
                       void* dqkv, int B, int L, int nq, int nkv, float scale, cudaStream_t stream) {
  BwdMaps m;
  int rc = make_bwd_maps(m, qkv, dout, (long long)B * L, nq, nkv);
  if (rc) return rc;
  AttnParams p = base_params(key_mask, nq, nkv, scale);
  p.lse2 = const_cast<float*>(lse);
  p.delta = delta;
  p.dqkv = (bf16*)dqkv;
  p.L = L;
  p.stat_h = L;
  {
    dim3 grid((L + BQ - 1) / BQ, nq, B);
    B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 1 ? attn_bwd_dq_tc_kernel<true> : attn_bwd_dq_tc_kernel<false>, dim3(grid), dim3(384), SMEM_DQ, stream, m.q128, m.d128, m.q64, p));
    B200RL_LAUNCH_OK();
  }
  {
    dim3 grid((L + BQ - 1) / BQ, nkv, B);
    B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 2 ? attn_bwd_dkv_tc_kernel<true> : attn_bwd_dkv_tc_kernel<false>, dim3(grid), dim3(384), SMEM_DKV, stream, m.q128, m.q64, m.d64, p));
    B200RL_LAUNCH_OK();
  }
  return 0;
}

// ---- packed (shared-prompt) layout: lse / delta are indexed [head][row] (stat_h = rows) -------------------------
int attn_fwd_seg_launch(const void* qkv, const int* key_mask, void* out, float* lse, long long rows, int nq, int nkv,
                        float scale, const QBlock* qblocks_dev, int n_qblocks, cudaStream_t stream) {
  CUtensorMap tm;
  int rc = make_rows_map(&tm, qkv, rows, (long long)(nq + 2 * nkv) * HD, 128);
  if (rc) return rc;
  if ((rc = set_attrs())) return rc;
  AttnParams p = base_params(key_mask, nq, nkv, scale);
  p.out = (bf16*)out;
  p.lse2 = lse;
  p.qblocks = qblocks_dev;
  p.stat_h = (int)rows;
  dim3 grid(n_qblocks, nq, 1);
  B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 0 ? attn_fwd_tc_kernel<true> : attn_fwd_tc_kernel<false>, dim3(grid), dim3(384), SMEM_FWD, stream, tm, p));
  B200RL_LAUNCH_OK();
  return 0;
}

int attn_bwd_seg_launch(const void* qkv, const int* key_mask, const void* dout, const float* lse, const float* delta,
                        void* dqkv, float* kv_part, long long rows, int nq, int nkv, float scale,
                        const QBlock* qblocks_dev, int n_qblocks, const KBlock* kblocks_dev, int n_kblocks,
                        const int* red_start_dev, const int* red_list_dev, cudaStream_t stream) {
  BwdMaps m;
  int rc = make_bwd_maps(m, qkv, dout, rows, nq, nkv);
  if (rc) return rc;
  AttnParams p = base_params(key_mask, nq, nkv, scale);
  p.lse2 = const_cast<float*>(lse);
  p.delta = delta;
  p.dqkv = (bf16*)dqkv;
  p.kv_part = kv_part;
  p.qblocks = qblocks_dev;
  p.kblocks = kblocks_dev;
  p.stat_h = (int)rows;
  {
    dim3 grid(n_qblocks, nq, 1);
    B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 1 ? attn_bwd_dq_tc_kernel<true> : attn_bwd_dq_tc_kernel<false>, dim3(grid), dim3(384), SMEM_DQ, stream, m.q128, m.d128, m.q64, p));
    B200RL_LAUNCH_OK();
  }
  {
    dim3 grid(n_kblocks, nkv, 1);
    B200RL_CUDA_OK(launch_pdl(p.prof && g_attn_prof_which == 2 ? attn_bwd_dkv_tc_kernel<true> : attn_bwd_dkv_tc_kernel<false>, dim3(grid), dim3(384), SMEM_DKV, stream, m.q128, m.q64, m.d64, p));
    B200RL_LAUNCH_OK();
  }
  B200RL_CUDA_OK(launch_pdl(kv_reduce_kernel, dim3((unsigned)rows), dim3(128), 0, stream, kv_part, red_start_dev, red_list_dev, (bf16*)dqkv, nq, nkv));
  B200RL_LAUNCH_OK();
  return 0;
}

void g_attn_prof_set(unsigned long long* b, int which) { g_attn_prof = b; g_attn_prof_which = which; }

}  // namespace b200rl

// debug: per-phase cycle counters of the tcgen05 attention forward kernel (18 x uint64 device buffer, or NULL = off)
extern "C" int b200rl_attn_set_prof(void* buf_dev, int which) {
  b200rl::g_attn_prof_set(reinterpret_cast<unsigned long long*>(buf_dev), which);
  return 0;
}
