// G4 on the 5th-gen tensor cores: causal GQA flash attention FORWARD with tcgen05.mma, S and the per-block
// P.V product in TMEM, Q/K/V tiles by TMA (head_dim 128; other head dims use the mma.sync kernels in
// attention.cu).  Same contract as attn_fwd_kernel: qkv [B*L, (nq+2nkv)*128] bf16 (RoPE applied), key-padding
// mask, out [B*L, nq*128] bf16, lse2 [B, nq, L] fp32 (log2 domain, +inf for rows without any visible key).
//
// One CTA per (128-query block, q head, sequence); key blocks of 128:
//   warp 0  : TMA producer  - Q once, then (K_j, V_j) into a 2-stage ring (3-D tensor map over [B][L][cols] so
//             rows beyond a sequence's end are zero-filled)
//   warp 1  : MMA issuer    - S_j = Q.K_j^T  (8 x UMMA 128x128x16, both operands K-major) into TMEM S[j&1];
//                             O_j = P_j.V_j  (A = P_j from smem, B = V_j MN-major) into TMEM O[j&1]; software
//                             pipelined: QK_{j+1} is issued before P.V_j so the tensor core overlaps the softmax
//   warp 2  : TMEM allocator (512 columns: S0 S1 O0 O1)
//   warps 4-7: softmax, one query row per thread (TMEM lane = row): pass 1 row max, pass 2 p = exp2(s - m) ->
//             bf16 P tile in the canonical 128B-swizzled K-major layout; running (m, l) and the output row live
//             in registers, O_reg = O_reg * exp2(m_old - m_new) + O_j (no TMEM read-modify-write).
#include "common.cuh"

namespace b200rl {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;   // queries per CTA
constexpr int BKV = 128;  // keys per block
constexpr int TILE_BYTES = 128 * HD * 2;  // 32 KB: [2 halves of 64 cols][128 rows][128 B]

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
constexpr uint32_t idesc_128x128(bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(128 >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct FwdParams {
  const int* key_mask;
  bf16* out;
  float* lse2;
  int L, nq, nkv;
  float scale_log2;
};

__global__ void __launch_bounds__(384, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], s_empty[2], o_full[2], o_empty[2], p_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ uint32_t s_maskw[2][4];     // key-padding bitmask of the 128 keys of a block
  __shared__ float s_mx[2][2][BQ];       // [stage][warpgroup][row] partial row max
  __shared__ float s_l[2][BQ];           // [warpgroup][row] partial row sums (final combine)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [Q 32K][K0 32K][V0 32K][K1 32K][V1 32K][P 32K]
  const uint32_t sQ = smem_base, sKV = smem_base + TILE_BYTES, sP = smem_base + 5 * TILE_BYTES;

  // heavy (late) query blocks first: causal work per CTA grows with the block index
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int g = h / (p.nq / p.nkv);
  const int q0 = qb * BQ;
  const int n_kb = min(qb + 1, (p.L + BKV - 1) / BKV);  // causal: key blocks 0..qb

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 256);
      mbar_init(&o_full[s], 1);
      mbar_init(&o_empty[s], 256);
    }
    mbar_init(&p_full, 256);
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

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer: Q, then the K ring (a K slot is free as soon as QK_j retired) ====
    mbar_arrive_expect_tx(&q_full, TILE_BYTES);
    tma_load_3d(smem_gen, &tm, &q_full, h * HD, q0, b);
    tma_load_3d(smem_gen + TILE_BYTES / 2, &tm, &q_full, h * HD + 64, q0, b);
    const int kcol = (p.nq + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1u);
      uint8_t* k_dst = smem_gen + TILE_BYTES * (1 + 2 * st);
      mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
      tma_load_3d(k_dst, &tm, &k_full[st], kcol, j * BKV, b);
      tma_load_3d(k_dst + TILE_BYTES / 2, &tm, &k_full[st], kcol + 64, j * BKV, b);
    }
  } else if (warp == 3 && lane == 0) {
    // ===================== TMA producer: V ring (a V slot is free when P.V_j retired) =====================
    const int vcol = (p.nq + p.nkv + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1u);
      uint8_t* v_dst = smem_gen + TILE_BYTES * (2 + 2 * st);
      mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
      tma_load_3d(v_dst, &tm, &v_full[st], vcol, j * BKV, b);
      tma_load_3d(v_dst + TILE_BYTES / 2, &tm, &v_full[st], vcol + 64, j * BKV, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = idesc_128x128(false);
    constexpr uint32_t idesc_pv = idesc_128x128(true);
    mbar_wait(&q_full, 0);
    auto issue_pv = [&](int j) {
      const int st = j & 1;
      mbar_wait(&p_full, j & 1);
      mbar_wait(&v_full[st], (j >> 1) & 1);
      mbar_wait(&o_empty[st], ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t v = sKV + TILE_BYTES * (2 * st + 1);
      const uint32_t tmem_o = tmem_base + 256 + st * 128;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        // A = P: K-major over keys, two 64-key halves; B = V stored [d-half][128 keys][128 B]: MN-major,
        // LBO = 16 KB between the two 64-wide d slabs, SBO = 1 KB between 8-key groups, 2 KB per K=16 step
        const uint64_t da = smem_desc_sw128(sP + (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32, 16, 1024);
        const uint64_t db = smem_desc_sw128(v + kk * 2048, TILE_BYTES / 2, 1024);
        umma_bf16(tmem_o, da, db, idesc_pv, kk > 0 ? 1u : 0u);
      }
      umma_commit(&o_full[st]);
      umma_commit(&v_empty[st]);
    };
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t k = sKV + TILE_BYTES * (2 * st);
      const uint32_t tmem_s = tmem_base + st * 128;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = smem_desc_sw128(sQ + (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32, 16, 1024);
        const uint64_t db = smem_desc_sw128(k + (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32, 16, 1024);
        umma_bf16(tmem_s, da, db, idesc_qk, kk > 0 ? 1u : 0u);
      }
      umma_commit(&s_full[st]);
      umma_commit(&k_empty[st]);
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(n_kb - 1);
  } else if (warp >= 4) {
    // ===================== softmax / output =====================
    // two warpgroups share every query row: warpgroup wg owns keys [64 wg, 64 wg + 64) of each block (one
    // swizzle atom of the P tile) and output columns d in [64 wg, 64 wg + 64); TMEM lane = row for both
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, corr_prev = 1.f;

    auto absorb = [&](int j, float corr) {  // O_reg = O_reg * corr + O_j   (own 64 columns)
      const int st = j & 1;
      mbar_wait(&o_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t a0[32], a1[32];
      const uint32_t to = tmem_base + 256 + st * 128 + lane_addr + wg * 64;
      tmem_ld_32x32(to, a0);
      tmem_ld_32x32(to + 32, a1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = o[i] * corr + __uint_as_float(a0[i]);
        o[32 + i] = o[32 + i] * corr + __uint_as_float(a1[i]);
      }
      tc_fence_before();
      mbar_arrive(&o_empty[st]);
    };

    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      const int k0 = j * BKV;
      if (wg == 0) {  // key-padding bitmask of this block
        const int mk = (k0 + r < p.L) ? p.key_mask[(long long)b * p.L + k0 + r] : 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, mk != 0);
        if (lane == 0) s_maskw[st][quad] = bal;
      }
      named_bar_sync(1, 256);
      const uint32_t w0 = s_maskw[st][2 * wg], w1 = s_maskw[st][2 * wg + 1];
      const bool diag = (j == qb);                                   // only the diagonal block needs key <= q
      const bool plain = !diag && (w0 & w1) == 0xFFFFFFFFu;          // no masking at all (the common case)
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_base + st * 128 + lane_addr + wg * 64;
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(ts, v0);
      tmem_ld_32x32(ts + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[st]);  // scores are in registers: S[st] may be overwritten by QK_{j+2}
      const int kbase = k0 + wg * 64;
      // ---- row max over the own 64 keys, then exchange with the other warpgroup ----
      float mx = -INFINITY;
      if (plain) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
        mx *= p.scale_log2;  // scale > 0
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool ok0 = ((w0 >> i) & 1u) && (!diag || kbase + i <= q);
          const bool ok1 = ((w1 >> i) & 1u) && (!diag || kbase + 32 + i <= q);
          const float a = ok0 ? __uint_as_float(v0[i]) * p.scale_log2 : -INFINITY;
          const float c = ok1 ? __uint_as_float(v1[i]) * p.scale_log2 : -INFINITY;
          v0[i] = __float_as_uint(a);  // keep the masked, scaled score
          v1[i] = __float_as_uint(c);
          mx = fmaxf(mx, fmaxf(a, c));
        }
      }
      s_mx[st][wg][r] = mx;
      named_bar_sync(2, 256);
      mx = fmaxf(mx, s_mx[st][wg ^ 1][r]);
      const float m_new = fmaxf(m_run, mx);
      const float mu = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = ex2_approx(m_run - mu);
      // fold the previous block's P.V into the register accumulator (also guarantees the previous P.V has
      // finished reading the P tile before it is overwritten below)
      if (j > 0) absorb(j - 1, corr_prev);
      // ---- p = exp2(s - m) -> bf16 into this warpgroup's 64-key atom of the P tile ----
      float rs = 0.f;
      const uint32_t prow = sP + wg * (TILE_BYTES / 2) + r * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float pf[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = __uint_as_float(half == 0 ? v0[i] : v1[i]);
          const float pv = plain ? ex2_approx(sv * p.scale_log2 - mu) : ex2_approx(sv - mu);  // masked -> 2^-inf = 0
          pf[i] = pv;
          rs += pv;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = half * 4 + u;  // 16-byte chunk 0..7 within the 128-byte row of the atom
          const bf16x8 pk = pack8(&pf[u * 8]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((cc ^ (r & 7)) << 4)),
                       "r"(pk.u.x), "r"(pk.u.y), "r"(pk.u.z), "r"(pk.u.w)
                       : "memory");
        }
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      corr_prev = corr;
      fence_proxy_async_smem();  // P tile visible to the tensor core (async proxy)
      mbar_arrive(&p_full);
    }
    // corr bookkeeping: O_reg before absorbing block j is relative to m_{j-1}; corr_j = exp2(m_{j-1} - m_j) was
    // computed when block j's scores were processed and O_j (from P_j) is relative to m_j.
    absorb(n_kb - 1, corr_prev);
    s_l[wg][r] = l_run;
    named_bar_sync(3, 256);
    const float l_tot = l_run + s_l[wg ^ 1][r];
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    if (q < p.L) {
      if (wg == 0) p.lse2[((long long)b * p.nq + h) * p.L + q] = l_tot > 0.f ? m_run + log2f(l_tot) : INFINITY;
      bf16* dst = p.out + ((long long)b * p.L + q) * p.nq * HD + h * HD + wg * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = o[i + u] * inv;
        *reinterpret_cast<bf16x8*>(dst + i) = pack8(f);
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
//       dS = scale * P * (dP - delta) as a bf16 smem tile  ->  dQ += dS_j.K_j (TMEM accumulator).
//   dKV kernel: CTA = (128-key block, kv head, sequence), inner loop over (q head of the GQA group, 64-query
//       block at or after the key block): S^T = K.Q^T, dP^T = V.dO^T (TMEM, double buffered) -> P^T, dS^T tiles
//       -> dV += P^T.dO, dK += dS^T.Q (TMEM accumulators, summed over the whole group: no atomics).
// The 64-row streamed tiles are stored [d-half][64 rows][128 B]; the SAME smem tile is read K-major (rows = N)
// by the score MMAs and MN-major (rows = K) by the gradient MMAs.
// ==================================================================================================
constexpr int HALF_TILE = 64 * HD * 2;  // 16 KB: [2 d-halves of 8 KB][64 rows][128 B]

constexpr uint32_t idesc_n(int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

struct BwdParams {
  const int* key_mask;
  const float* lse2;
  const float* delta;
  bf16* dqkv;
  int L, nq, nkv;
  float scale, scale_log2;
};

// this thread's row r of a [128 rows][64 cols] bf16 K-major SW128 tile: store columns col0..col0+31
__device__ __forceinline__ void store_row32_sw128(uint32_t tile, int r, int col0, const float* f) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int cc = col0 / 8 + u;  // 16-byte chunk 0..7 within the 128-byte row
    const uint32_t addr = tile + r * 128 + ((cc ^ (r & 7)) << 4);
    const bf16x8 pk = pack8(&f[u * 8]);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                 "r"(pk.u.x), "r"(pk.u.y), "r"(pk.u.z), "r"(pk.u.w)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// dQ
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_q128, const __grid_constant__ CUtensorMap tm_do128,
                      const __grid_constant__ CUtensorMap tm_kv64, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t qdo_full, kv_full[3], kv_empty[3], sp_full[2], sp_empty[2], ds_full[2], ds_empty[2], dq_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int s_mask[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [Q 32K][dO 32K][3 x (K 16K, V 16K)][dS0 16K][dS1 16K]  = 192 KB
  const uint32_t sQ = smem_base, sdO = sQ + TILE_BYTES, sKV = sdO + TILE_BYTES, sDS = sKV + 6 * HALF_TILE;
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int g = h / (p.nq / p.nkv);
  const int q0 = qb * BQ;
  const int n_kb = min((q0 + BQ + 63) / 64, (p.L + 63) / 64);  // 64-key blocks that intersect keys <= q0+127

  if (threadIdx.x == 0) {
    mbar_init(&qdo_full, 1);
    mbar_init(&dq_full, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sp_full[s], 1);
      mbar_init(&sp_empty[s], 256);
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
  // TMEM columns: S0 [0,64) S1 [64,128) dP0 [128,192) dP1 [192,256) dQ [256,384)

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    mbar_arrive_expect_tx(&qdo_full, 2 * TILE_BYTES);
    tma_load_3d(smem_gen, &tm_q128, &qdo_full, h * HD, q0, b);
    tma_load_3d(smem_gen + TILE_BYTES / 2, &tm_q128, &qdo_full, h * HD + 64, q0, b);
    tma_load_3d(smem_gen + TILE_BYTES, &tm_do128, &qdo_full, h * HD, q0, b);
    tma_load_3d(smem_gen + TILE_BYTES + TILE_BYTES / 2, &tm_do128, &qdo_full, h * HD + 64, q0, b);
    const int kcol = (p.nq + g) * HD, vcol = (p.nq + p.nkv + g) * HD;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j % 3;
      mbar_wait(&kv_empty[st], ((j / 3) & 1) ^ 1u);
      uint8_t* kd = smem_gen + 2 * TILE_BYTES + st * 2 * HALF_TILE;
      uint8_t* vd = kd + HALF_TILE;
      mbar_arrive_expect_tx(&kv_full[st], 2 * HALF_TILE);
      tma_load_3d(kd, &tm_kv64, &kv_full[st], kcol, j * 64, b);
      tma_load_3d(kd + HALF_TILE / 2, &tm_kv64, &kv_full[st], kcol + 64, j * 64, b);
      tma_load_3d(vd, &tm_kv64, &kv_full[st], vcol, j * 64, b);
      tma_load_3d(vd + HALF_TILE / 2, &tm_kv64, &kv_full[st], vcol + 64, j * 64, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t id_s = idesc_n(64, false);    // [128 q] x [64 keys], both K-major over d
    constexpr uint32_t id_dq = idesc_n(128, true);   // [128 q] x [128 d], B = K_j MN-major (k = keys)
    mbar_wait(&qdo_full, 0);
    auto issue_dq = [&](int j) {
      const int st = j & 1, k3 = j % 3;
      mbar_wait(&ds_full[st], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t kt = sKV + k3 * 2 * HALF_TILE;
      const uint32_t ds = sDS + st * HALF_TILE;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 64 keys = 4 x K16
        const uint64_t da = smem_desc_sw128(ds + kk * 32, 16, 1024);
        const uint64_t db = smem_desc_sw128(kt + kk * 2048, HALF_TILE / 2, 1024);
        umma_bf16(tmem_base + 256, da, db, id_dq, (j > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&kv_empty[k3]);
      umma_commit(&ds_empty[st]);
    };
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1, k3 = j % 3;
      mbar_wait(&kv_full[k3], (j / 3) & 1);
      mbar_wait(&sp_empty[st], ((j >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t kt = sKV + k3 * 2 * HALF_TILE, vt = kt + HALF_TILE;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {  // d = 128 = 8 x K16, two d-halves
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(tmem_base + st * 64, smem_desc_sw128(sQ + aoff, 16, 1024), smem_desc_sw128(kt + boff, 16, 1024),
                  id_s, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(tmem_base + 128 + st * 64, smem_desc_sw128(sdO + aoff, 16, 1024),
                  smem_desc_sw128(vt + boff, 16, 1024), id_s, kk > 0 ? 1u : 0u);
      }
      umma_commit(&sp_full[st]);
      if (j > 0) issue_dq(j - 1);
    }
    issue_dq(n_kb - 1);
    umma_commit(&dq_full);
  } else if (warp >= 4) {
    // ===================== dS producer: two warpgroups, each 32 of the 64 keys of a block =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int t = threadIdx.x - 128;
    const int q = q0 + r;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const long long sidx = ((long long)b * p.nq + h) * p.L;
    const float lse = q < p.L ? p.lse2[sidx + q] : INFINITY;
    const float del = q < p.L ? p.delta[sidx + q] : 0.f;
    for (int j = 0; j < n_kb; ++j) {
      const int st = j & 1;
      const int k0 = j * 64;
      if (t < 64) {  // warps 4 and 5: key-padding bitmask of the block's two 32-key chunks
        const int mk = (k0 + t < p.L) ? p.key_mask[(long long)b * p.L + k0 + t] : 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, mk != 0);
        if (lane == 0) s_mask[st][t >> 5] = (int)bal;
      }
      named_bar_sync(1, 256);
      const int c = wg;
      const uint32_t mw = (uint32_t)s_mask[st][c];
      const int kc0 = k0 + c * 32;
      // no masking at all when every key of the chunk is real and at or before the tile's first query
      const bool plain = (mw == 0xFFFFFFFFu) && (kc0 + 31 <= q0);
      mbar_wait(&sp_full[st], (j >> 1) & 1);
      mbar_wait(&ds_empty[st], ((j >> 1) & 1) ^ 1u);  // dQ MMA of block j-2 finished reading dS[st]
      tc_fence_after();
      {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem_base + st * 64 + lane_addr + c * 32, sv);
        tmem_ld_32x32(tmem_base + 128 + st * 64 + lane_addr + c * 32, dv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sp_empty[st]);
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
            const bool ok = ((mw >> i) & 1u) && (kc0 + i <= q);
            const float pr = ok ? ex2_approx(__uint_as_float(sv[i]) * p.scale_log2 - lse) : 0.f;
            f[i] = p.scale * pr * (__uint_as_float(dv[i]) - del);
          }
        }
        store_row32_sw128(sDS + st * HALF_TILE, r, c * 32, f);
      }
      fence_proxy_async_smem();
      mbar_arrive(&ds_full[st]);
    }
    // ---- write dQ ----
    mbar_wait(&dq_full, 0);
    tc_fence_after();
    {
      // the TMEM loads are .sync.aligned: every lane executes them, only the stores are predicated
      bf16* dst = p.dqkv + ((long long)b * p.L + (q < p.L ? q : 0)) * (long long)(p.nq + 2 * p.nkv) * HD + h * HD;
#pragma unroll
      for (int c = 2 * wg; c < 2 * wg + 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + 256 + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (q < p.L) {
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
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_kv128, const __grid_constant__ CUtensorMap tm_q64,
                       const __grid_constant__ CUtensorMap tm_do64, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t kv_full, qd_full[3], qd_empty[3], sp_full[2], sp_empty[2], ds_full[2], ds_empty[2], out_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_lse[2][64], s_del[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // [K 32K][V 32K][3 x (Q 16K, dO 16K)][PT0 16K][dST0 16K][PT1 16K][dST1 16K] = 224 KB
  const uint32_t sK = smem_base, sV = sK + TILE_BYTES, sQD = sV + TILE_BYTES, sPD = sQD + 6 * HALF_TILE;
  const int kb = blockIdx.x, g = blockIdx.y, b = blockIdx.z;   // key block 0 (most work) is scheduled first
  const int group = p.nq / p.nkv;
  const int k0 = kb * BKV;
  const int qb0 = k0 / 64;                         // first 64-query block that can see these keys
  const int nqb = (p.L + 63) / 64 - qb0;           // 64-query blocks per head
  const int n_it = group * nqb;

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    mbar_init(&out_full, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sp_full[s], 1);
      mbar_init(&sp_empty[s], 256);
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
  // TMEM columns: S^T0 [0,64) S^T1 [64,128) dP^T0 [128,192) dP^T1 [192,256) dK [256,384) dV [384,512)

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    const int kcol = (p.nq + g) * HD, vcol = (p.nq + p.nkv + g) * HD;
    mbar_arrive_expect_tx(&kv_full, 2 * TILE_BYTES);
    tma_load_3d(smem_gen, &tm_kv128, &kv_full, kcol, k0, b);
    tma_load_3d(smem_gen + TILE_BYTES / 2, &tm_kv128, &kv_full, kcol + 64, k0, b);
    tma_load_3d(smem_gen + TILE_BYTES, &tm_kv128, &kv_full, vcol, k0, b);
    tma_load_3d(smem_gen + TILE_BYTES + TILE_BYTES / 2, &tm_kv128, &kv_full, vcol + 64, k0, b);
    for (int it = 0; it < n_it; ++it) {
      const int st = it % 3;
      const int h = g * group + it / nqb;
      const int qs = (qb0 + it % nqb) * 64;
      mbar_wait(&qd_empty[st], ((it / 3) & 1) ^ 1u);
      uint8_t* qd = smem_gen + 2 * TILE_BYTES + st * 2 * HALF_TILE;
      uint8_t* dd = qd + HALF_TILE;
      mbar_arrive_expect_tx(&qd_full[st], 2 * HALF_TILE);
      tma_load_3d(qd, &tm_q64, &qd_full[st], h * HD, qs, b);
      tma_load_3d(qd + HALF_TILE / 2, &tm_q64, &qd_full[st], h * HD + 64, qs, b);
      tma_load_3d(dd, &tm_do64, &qd_full[st], h * HD, qs, b);
      tma_load_3d(dd + HALF_TILE / 2, &tm_do64, &qd_full[st], h * HD + 64, qs, b);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t id_s = idesc_n(64, false);    // [128 keys] x [64 queries], K-major over d
    constexpr uint32_t id_g = idesc_n(128, true);    // [128 keys] x [128 d], B = dO / Q MN-major (k = queries)
    mbar_wait(&kv_full, 0);
    auto issue_grad = [&](int it) {
      const int st = it & 1, q3 = it % 3;
      mbar_wait(&ds_full[st], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t qt = sQD + q3 * 2 * HALF_TILE, dt = qt + HALF_TILE;
      const uint32_t pt = sPD + st * 2 * HALF_TILE, dst = pt + HALF_TILE;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // 64 queries = 4 x K16
        umma_bf16(tmem_base + 384, smem_desc_sw128(pt + kk * 32, 16, 1024),
                  smem_desc_sw128(dt + kk * 2048, HALF_TILE / 2, 1024), id_g, (it > 0 || kk > 0) ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        umma_bf16(tmem_base + 256, smem_desc_sw128(dst + kk * 32, 16, 1024),
                  smem_desc_sw128(qt + kk * 2048, HALF_TILE / 2, 1024), id_g, (it > 0 || kk > 0) ? 1u : 0u);
      }
      umma_commit(&qd_empty[q3]);
      umma_commit(&ds_empty[st]);
    };
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1, q3 = it % 3;
      mbar_wait(&qd_full[q3], (it / 3) & 1);
      mbar_wait(&sp_empty[st], ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t qt = sQD + q3 * 2 * HALF_TILE, dt = qt + HALF_TILE;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(tmem_base + st * 64, smem_desc_sw128(sK + aoff, 16, 1024), smem_desc_sw128(qt + boff, 16, 1024),
                  id_s, kk > 0 ? 1u : 0u);
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t aoff = (kk >> 2) * (TILE_BYTES / 2) + (kk & 3) * 32;
        const uint32_t boff = (kk >> 2) * (HALF_TILE / 2) + (kk & 3) * 32;
        umma_bf16(tmem_base + 128 + st * 64, smem_desc_sw128(sV + aoff, 16, 1024),
                  smem_desc_sw128(dt + boff, 16, 1024), id_s, kk > 0 ? 1u : 0u);
      }
      umma_commit(&sp_full[st]);
      if (it > 0) issue_grad(it - 1);
    }
    if (n_it > 0) issue_grad(n_it - 1);
    umma_commit(&out_full);
  } else if (warp >= 4) {
    // ===================== P^T / dS^T producer: one KEY row per thread, two warpgroups x 32 queries =====================
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int t = threadIdx.x - 128;
    const int key = k0 + r;
    const bool key_ok = key < p.L && p.key_mask[(long long)b * p.L + key] != 0;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int st = it & 1;
      const int h = g * group + it / nqb;
      const int qs = (qb0 + it % nqb) * 64;
      if (t < 64) {
        const int qi = qs + t;
        const long long sidx = ((long long)b * p.nq + h) * p.L;
        s_lse[st][t] = qi < p.L ? p.lse2[sidx + qi] : INFINITY;   // +inf -> probability 0 for queries past the end
        s_del[st][t] = qi < p.L ? p.delta[sidx + qi] : 0.f;
      }
      named_bar_sync(1, 256);
      const int c = wg;
      const int qc0 = qs + c * 32;
      const bool need_cmp = qc0 < k0 + BKV - 1;   // some (key, query) pair of this chunk may violate key <= query
      mbar_wait(&sp_full[st], (it >> 1) & 1);
      mbar_wait(&ds_empty[st], ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t pt = sPD + st * 2 * HALF_TILE, dst = pt + HALF_TILE;
      {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem_base + st * 64 + lane_addr + c * 32, sv);
        tmem_ld_32x32(tmem_base + 128 + st * 64 + lane_addr + c * 32, dv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&sp_empty[st]);
        float fp[32], fs[32];
        const float4* l4 = reinterpret_cast<const float4*>(&s_lse[st][c * 32]);
        const float4* d4 = reinterpret_cast<const float4*>(&s_del[st][c * 32]);
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4) {
          const float4 lv = l4[v4], dl = d4[v4];
          const float ls[4] = {lv.x, lv.y, lv.z, lv.w}, ds4[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = v4 * 4 + u;
            float pr = ex2_approx(__uint_as_float(sv[i]) * p.scale_log2 - ls[u]);
            if (!key_ok || (need_cmp && key > qc0 + i)) pr = 0.f;  // select, never 0 * inf
            fp[i] = pr;
            fs[i] = p.scale * pr * (__uint_as_float(dv[i]) - ds4[u]);
          }
        }
        store_row32_sw128(pt, r, c * 32, fp);
        store_row32_sw128(dst, r, c * 32, fs);
      }
      fence_proxy_async_smem();
      mbar_arrive(&ds_full[st]);
    }
    // ---- write dK, dV ----
    mbar_wait(&out_full, 0);
    tc_fence_after();
    const long long stride = (long long)(p.nq + 2 * p.nkv) * HD;
    bf16* drow = p.dqkv + ((long long)b * p.L + (key < p.L ? key : 0)) * stride;
    {
      const int which = wg;
      bf16* dst = drow + (which == 0 ? (p.nq + g) : (p.nq + p.nkv + g)) * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + 256 + which * 128 + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (key < p.L) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = n_it > 0 ? __uint_as_float(v[u * 8 + i]) : 0.f;
            *reinterpret_cast<bf16x8*>(dst + c * 32 + u * 8) = pack8(f);
          }
        }
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

}  // namespace

// 3-D map over x[B][L][cols] bf16, box = 64 cols x 128 rows x 1 sequence, 128B swizzle
int make_seq_map(CUtensorMap* tm, const void* base, int B, int L, long long cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)L};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
  return 0;
}

int attn_fwd_tc_launch(const void* qkv, const int* key_mask, void* out, float* lse, int B, int L, int nq, int nkv,
                       float scale, cudaStream_t stream) {
  CUtensorMap tm;
  int rc = make_seq_map(&tm, qkv, B, L, (long long)(nq + 2 * nkv) * HD, 128);
  if (rc) return rc;
  FwdParams p;
  p.key_mask = key_mask;
  p.out = (bf16*)out;
  p.lse2 = lse;
  p.L = L;
  p.nq = nq;
  p.nkv = nkv;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int smem = 6 * TILE_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((L + BQ - 1) / BQ, nq, B);
  attn_fwd_tc_kernel<<<grid, 384, smem, stream>>>(tm, p);
  B200RL_LAUNCH_OK();
  return 0;
}


int attn_bwd_tc_launch(const void* qkv, const int* key_mask, const void* dout, const float* lse, const float* delta,
                       void* dqkv, int B, int L, int nq, int nkv, float scale, cudaStream_t stream) {
  const long long qcols = (long long)(nq + 2 * nkv) * HD, ocols = (long long)nq * HD;
  CUtensorMap q128, q64, d128, d64;
  int rc;
  if ((rc = make_seq_map(&q128, qkv, B, L, qcols, 128))) return rc;
  if ((rc = make_seq_map(&q64, qkv, B, L, qcols, 64))) return rc;
  if ((rc = make_seq_map(&d128, dout, B, L, ocols, 128))) return rc;
  if ((rc = make_seq_map(&d64, dout, B, L, ocols, 64))) return rc;
  BwdParams p;
  p.key_mask = key_mask;
  p.lse2 = lse;
  p.delta = delta;
  p.dqkv = (bf16*)dqkv;
  p.L = L;
  p.nq = nq;
  p.nkv = nkv;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  const int smem_dq = 2 * TILE_BYTES + 8 * HALF_TILE + 1024;
  const int smem_dkv = 2 * TILE_BYTES + 10 * HALF_TILE + 1024;
  if (!attr_set) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
    B200RL_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv));
    attr_set = true;
  }
  {
    dim3 grid((L + BQ - 1) / BQ, nq, B);
    attn_bwd_dq_tc_kernel<<<grid, 384, smem_dq, stream>>>(q128, d128, q64, p);
    B200RL_LAUNCH_OK();
  }
  {
    dim3 grid((L + BKV - 1) / BKV, nkv, B);
    attn_bwd_dkv_tc_kernel<<<grid, 384, smem_dkv, stream>>>(q128, q64, d64, p);
    B200RL_LAUNCH_OK();
  }
  return 0;
}

}  // namespace b200rl
