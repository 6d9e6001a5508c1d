// G1 (CTA-pair variant): tcgen05.mma.cta_group::2 GEMM — two CTAs of a 2-CTA cluster (one TPC) share a
// 256 x BN output tile.  Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 columns),
// the leader CTA issues one M=256 UMMA per K=16 step that reads both CTAs' shared memory, and each CTA
// keeps its 128 x BN half of the accumulator in its own TMEM.  Versus the single-CTA kernel this halves
// the B-operand shared-memory traffic per SM (the limiter measured at ~78 % of the tensor peak) and buys
// two more pipeline stages.
//
// Same contract as gemm_kernel (gemm_tcgen05.cu): C = alpha*(A1.B1^T + A2.B2^T) (+bias) (+residual),
// A K-major, B K-major (TN) or MN-major (dX form).  Synchronisation:
//   full[s]   (leader only, count 2) : leader arrive.expect_tx(2 x stage bytes) + peer remote arrive;
//                                      both CTAs' TMA loads complete_tx on the LEADER's barrier
//   empty[s]  (each CTA, count 1)    : tcgen05.commit.cta_group::2 ... multicast to both CTAs
//   tmem_full[a]  (each CTA, count 1): multicast commit after the last k-block of a tile
//   tmem_empty[a] (leader, count 2 EW): one arrive per epilogue warp of both CTAs (peer arrives remotely); EW = 4,
//                                      or 8 for the fused SwiGLU epilogues (FUSE 1 / 2, see GemmArgs::fuse)
//
// Tail split: with T tiles on P CTA pairs the last wave holds r = T mod P tiles (Qwen2.5-7B, M = 4446: the N = 3584
// GEMMs have 252 tiles on 74 pairs -> 3.4 waves, the 4th wave is 40 % full).  When 0 < r <= P/2 the last r tiles are
// each cut into S = min(4, P / r) K-ranges, so the last wave costs 1/S of a tile time (3.5 instead of 4 waves).
#include "gemm_common.cuh"
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>

namespace b200rl {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the LEADER's
// barrier (peer bit of the shared::cluster address cleared, as in cute SM100_TMA_2SM_LOAD).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint64_t* bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// work unit -> (tile, k-block range, part).  part < 0: whole tile.
__device__ __forceinline__ void unit_decode(const GemmParams& p, int unit, int kb_total, int& tile, int& kb0,
                                            int& kb1, int& part, int& sidx) {
  if (unit < p.tail_first) {
    tile = unit; kb0 = 0; kb1 = kb_total; part = -1; sidx = 0;
    return;
  }
  const int idx = unit - p.tail_first;
  sidx = idx / p.tail_split;
  part = idx - sidx * p.tail_split;
  tile = p.tail_first + sidx;
  const int per = (kb_total + p.tail_split - 1) / p.tail_split;
  kb0 = part * per;
  kb1 = min(kb_total, kb0 + per);
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// NT = 256-column sub-tiles per cluster tile.  NT = 1: one 256 x BN tile per work unit, accumulators double-buffered in
// TMEM (the epilogue of tile i overlaps the mainloop of tile i+1).  NT = 2 (BN = 256 only): a 256 x 512 "wide" tile —
// every k-block stages the A tile ONCE and two B sub-tiles, and issues two N = 256 UMMAs into the two TMEM accumulators.
// Per flop this moves 25 % fewer bytes from L2 into shared memory (48 KB instead of 64 KB per CTA per 2 x 256 x 256 x 64
// MACs): the 256 x 256 mainloop sits at the L2 -> SM fill limit, so the operand bytes, not the tensor
// pipe, set its speed.  Price: both accumulators belong to the same tile, so a sub-tile's epilogue only overlaps the
// next tile's mainloop from the moment ITS accumulator has been drained (tmem_empty per sub-tile).
// NF4 level table (same values as csrc/nf4.cu; __constant__ symbols do not link across translation units)
__constant__ float c_nf4_levels[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                                       -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                                       0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f,
                                       0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f,
                                       0.7229568362236023f, 1.0f};

// 64 NF4 codes (32 bytes: q0, q1) of one 128-byte K-/MN-major shared-memory row -> bf16(level * absmax), written as the
// eight 16-byte chunks of the row in the 128B-swizzle pattern TMA would have produced (chunk c at c ^ (row & 7)).
// Bit-identical to nf4_dequant_kernel (csrc/nf4.cu): fp32 product, round-to-nearest bf16, even element in the high nibble.
__device__ __forceinline__ void nf4_row_to_smem(uint32_t row_addr, int rsw, const uint4& q0, const uint4& q1, float am,
                                                const float2* lut) {
  const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float2 v = lut[(w[c] >> (8 * b)) & 0xFFu];
      o[b] = pack_bf16x2(v.x * am, v.y * am);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((c ^ rsw) << 4)), "r"(o[0]), "r"(o[1]),
                 "r"(o[2]), "r"(o[3])
                 : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {  // release at cluster scope
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

template <int BN, bool B_MN, int NT = 1>
struct PairCfg {
  static_assert(NT == 1 || BN == 256, "wide tiles are 2 x 256 columns");
  static constexpr int BH = BN / 2;                      // B columns staged per CTA (per sub-tile)
  static constexpr int B_SLABS = (BH + 63) / 64;         // MN-major: 64-column swizzle slabs (the last may be partly used)
  static constexpr int B_TILE_BYTES = B_MN ? B_SLABS * 8192 : BH * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + NT * B_TILE_BYTES;  // per CTA
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int STAGES_RAW = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

// EW = epilogue warps per CTA (4 or 8): the fused SwiGLU epilogues use 8, two per TMEM lane quadrant, each pair
// splitting the accumulator columns.
// NF4 = true: segment 1 of the B operand comes from NF4 storage (p.nf4_q / p.nf4_am).  Four extra producer warps per CTA
// (warps 4+EW ..) expand the 4-bit codes of every k-block into the stage's swizzled bf16 B tile(s) — the same bytes TMA
// would have delivered from a dequantised copy — and arrive on the leader's full barrier next to the TMA transaction of
// the A tile; the MMA issuer, the K-extension (LoRA) segment, the ext units and every epilogue are unchanged.
template <int BN, bool B_MN, int FUSE, int EW, int NT, bool NF4>
__global__ void __launch_bounds__(128 + 32 * EW + (NF4 ? 128 : 0), 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmExt, const GemmParams p) {
  using C = PairCfg<BN, B_MN, NT>;
  constexpr int STAGES = C::STAGES;
  constexpr int BM2 = 2 * BM;  // rows per pair tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float2 s_lut[NF4 ? 256 : 1];   // byte -> (level[hi nibble], level[lo nibble])
  if constexpr (NF4) {
    if (threadIdx.x < 256) s_lut[threadIdx.x] = make_float2(c_nf4_levels[threadIdx.x >> 4], c_nf4_levels[threadIdx.x & 15]);
  }

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NF4 ? 2 + 8 : 2);   // + the four dequant warps of both CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * EW);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmExt);
  }
  if (warp == 2) tmem_alloc2(&tmem_base_smem, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();  // prologue done: let the next kernel start its own, then wait for the previous kernel's data

  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;  // num_m_blocks counts 256-row pair tiles
  const int kb_total = p.kb1 + p.kb2;
  // work units: [0, n_ext) ext units (LoRA intermediate of m-block `unit`), then the tiles / tail K-ranges
  const int n_ext = p.n_ext;
  const int num_units = n_ext + (num_tiles <= p.tail_first ? num_tiles : p.tail_first + (num_tiles - p.tail_first) * p.tail_split);
  // bytes one CTA stages per k-block of an ext unit: its 128 rows of A + its half of the ext operand
  const uint32_t ext_b_bytes = B_MN ? ((p.ext_n / 2 + 63) / 64) * 8192u : (uint32_t)(p.ext_n / 2) * BK * 2;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      if (unit < n_ext) {
        // ---- ext unit: A rows of m-block `unit` x the whole ext operand, over K1 ----
        const int row0 = unit * BM2 + (int)rank * BM;
        for (int kb = 0; kb < p.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem_gen + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (A_TILE_BYTES + ext_b_bytes));
          tma_load_2d_pair(sa, &tmA1, &full_bar[stage], kb * BK, row0);
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tmExt, &full_bar[stage], kb * BK, (int)rank * (p.ext_n / 2));
          } else {
            for (int h = 0; h < (p.ext_n / 2 + 63) / 64; ++h)
              tma_load_2d_pair(sb + h * 8192, &tmExt, &full_bar[stage], (int)rank * (p.ext_n / 2) + h * 64, kb * BK);
          }
          if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        continue;
      }
      int tile, kb_begin, kb_end, part, sidx, m_blk, n_blk;
      unit_decode(p, unit - n_ext, kb_total, tile, kb_begin, kb_end, part, sidx);
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, p.gm, m_blk, n_blk);
      const int row0 = m_blk * BM2 + (int)rank * BM;
      bool ext_ready = n_ext == 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem_gen + stage * C::STAGE_BYTES;
        const bool seg2 = kb >= p.kb1;
        if (seg2 && !ext_ready) {
          // the K-extension reads U rows of this CTA, written by the four epilogue warps of the same rank of the ext unit
          // of this m-block (a lower-numbered unit: it is running or done, never queued behind this one)
          const int* f = p.ext_flags + (m_blk * 2 + (int)rank) * 4;
          for (int q = 0; q < 4; ++q) {
            long long t0 = clock64();
            while (ld_acquire_gpu(f + q) != p.ext_epoch) {
              if (clock64() - t0 > 40000000000LL) {
                printf("b200rl: gemm ext-unit flag wait timed out (block %d, m-block %d)\n", blockIdx.x, m_blk);
                __trap();
              }
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of U -> this thread's TMA (async proxy) reads
          ext_ready = true;
        }
        const CUtensorMap* ta = seg2 ? &tmA2 : &tmA1;
        const CUtensorMap* tb = seg2 ? &tmB2 : &tmB1;
        const int k0 = (seg2 ? kb - p.kb1 : kb) * BK;
        const bool b_by_warps = NF4 && !seg2;   // the dequant warps fill the B tile(s) of this stage
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], b_by_warps ? 2 * A_TILE_BYTES : 2 * C::STAGE_BYTES);
        tma_load_2d_pair(sa, ta, &full_bar[stage], k0, row0);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (b_by_warps) break;
          const int n_sub = n_blk * NT + t;
          // FUSE 1: CTA 0 stages 128 gate rows of the weight, CTA 1 the 128 up rows with the same index
          const int col0 = FUSE == 1 ? n_sub * C::BH + (int)rank * p.fuse_I : n_sub * BN + (int)rank * C::BH;
          uint8_t* sb = sa + A_TILE_BYTES + t * C::B_TILE_BYTES;
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, tb, &full_bar[stage], k0, col0);
          } else {
#pragma unroll
            for (int h = 0; h < C::B_SLABS; ++h)
              tma_load_2d_pair(sb + h * 8192, tb, &full_bar[stage], col0 + h * 64, k0);
          }
        }
        if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && leader && elect_one_sync()) {
    // ===================== MMA issuer (leader CTA only) =====================
    // instruction M = 256 (both CTAs), N = BN
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM2 >> 4) << 24);
    const uint32_t idesc_ext = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(p.ext_n >> 3) << 17) | ((uint32_t)(BM2 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;    // units processed so far (NT = 1: accumulator slot = local & 1; NT = 2: uses of slot 0)
    int local1 = 0;   // NT = 2: uses of slot 1 (ext units only touch slot 0)
    for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++local) {
      if (unit < n_ext) {
        const int acc = NT == 1 ? (local & 1) : 0;
        const uint32_t ph = NT == 1 ? ((local >> 1) & 1) : (local & 1);
        mbar_wait(&tmem_empty_bar[acc], ph ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < p.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t a_lo = smem_desc_lo(sa, 16);
          const uint32_t b_lo = smem_desc_lo(sa + A_TILE_BYTES, B_MN ? 8192 : 16);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_pair(tmem_d, smem_desc_at(a_lo, k * 32), smem_desc_at(b_lo, B_MN ? k * 2048 : k * 32), idesc_ext,
                           (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_pair(&tmem_full_bar[acc]);
        continue;
      }
      int tile, kb_begin, kb_end, part, sidx;
      unit_decode(p, unit - n_ext, kb_total, tile, kb_begin, kb_end, part, sidx);
      // accumulator slot / barrier phase: NT = 1 alternates the two slots unit by unit; NT = 2 uses slot t for
      // sub-tile t of every tile (slot 0 also serves the ext units, hence one use counter per slot)
      const int acc0 = NT == 1 ? (local & 1) : 0;
      const uint32_t acc_phase = NT == 1 ? ((local >> 1) & 1) : (local & 1);
      const uint32_t acc_phase1 = local1 & 1;
      if constexpr (NT == 1) {
        mbar_wait(&tmem_empty_bar[acc0], acc_phase ^ 1u);
        tc_fence_after();
      }
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t a_lo = smem_desc_lo(sa, 16);   // descriptors: built once per k-block, an offset added per MMA
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (NT == 2 && kb == kb_begin) {   // sub-tile t may start as soon as the epilogue drained ITS accumulator
            mbar_wait(&tmem_empty_bar[t], (t == 0 ? acc_phase : acc_phase1) ^ 1u);
            tc_fence_after();
          }
          const uint32_t tmem_d = tmem_base + (acc0 + t) * C::ACC_STRIDE;
          const uint32_t b_lo = smem_desc_lo(sa + A_TILE_BYTES + t * C::B_TILE_BYTES, B_MN ? 8192 : 16);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_pair(tmem_d, smem_desc_at(a_lo, k * 32), smem_desc_at(b_lo, B_MN ? k * 2048 : k * 32), idesc,
                           (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit_pair(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) umma_commit_pair(&tmem_full_bar[acc0 + t]);
      ++local1;
    }
  } else if (NF4 && warp >= 4 + EW) {
    // ===================== NF4 dequant producers (both CTAs, 4 warps = 128 threads) =====================
    // K-major B (forward, W [N, K1]): thread dt owns row dt of each 128-row sub-tile: 64 codes = 32 packed bytes per
    // k-block, one absmax.  MN-major B (dX form, W [K1, N]): thread dt owns k-row dt/2 and the 64-column half dt%2.
    const int dt = (int)threadIdx.x - (128 + 32 * EW);
    const long long kblocks_per_row = B_MN ? (p.N >> 6) : (p.K1 >> 6);   // absmax entries per matrix row
    int stage = 0;
    uint32_t phase = 0;
    auto arrive_full = [&](int st) {
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&full_bar[st]);
        else mbar_arrive_cluster(&full_bar[st], 0);
      }
    };
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      if (unit < n_ext) {   // ext units stage TMA operands only: arrive to keep the barrier count uniform
        for (int kb = 0; kb < p.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          arrive_full(stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      int tile, kb_begin, kb_end, part, sidx, m_blk, n_blk;
      unit_decode(p, unit - n_ext, kb_total, tile, kb_begin, kb_end, part, sidx);
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, p.gm, m_blk, n_blk);
      // global coordinates of this thread's codes inside sub-tile t (k-block independent part)
      long long row_off[NT];      // element offset of the first code of k-block 0
      long long am_off[NT];       // absmax index of k-block 0
      bool in_range[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int n_sub = n_blk * NT + t;
        const int col0 = FUSE == 1 ? n_sub * C::BH + (int)rank * p.fuse_I : n_sub * BN + (int)rank * C::BH;
        if constexpr (!B_MN) {
          const int n = col0 + dt;                       // matrix row (output feature)
          in_range[t] = n < (FUSE == 1 ? 2 * p.fuse_I : p.N) && (FUSE != 1 || n_sub * 128 < p.fuse_I);
          row_off[t] = (long long)n * p.K1;
          am_off[t] = (long long)n * kblocks_per_row;
        } else {
          const int n = col0 + (dt & 1) * 64;            // first of this thread's 64 columns
          in_range[t] = n < p.N;
          row_off[t] = n;                                // + k * N per k-row
          am_off[t] = n >> 6;
        }
      }
      constexpr int PF = 4;                              // k-blocks of packed codes in flight per thread
      uint4 q[PF][NT][2];
      float am[PF][NT];
      auto issue = [&](int slot, int kb) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const bool ok = in_range[t] && kb < kb_end && kb < p.kb1;
          long long e, a;
          if constexpr (!B_MN) {
            e = row_off[t] + (long long)kb * BK;
            a = am_off[t] + kb;
          } else {
            const long long k = (long long)kb * BK + (dt >> 1);
            e = k * p.N + row_off[t];
            a = k * kblocks_per_row + am_off[t];
          }
          if (ok) {
            const uint4* src = reinterpret_cast<const uint4*>(p.nf4_q + (e >> 1));
            q[slot][t][0] = __ldg(src);
            q[slot][t][1] = __ldg(src + 1);
            am[slot][t] = __ldg(p.nf4_am + a);
          } else {
            q[slot][t][0] = make_uint4(0x77777777u, 0x77777777u, 0x77777777u, 0x77777777u);   // code 7 = level 0.0
            q[slot][t][1] = q[slot][t][0];
            am[slot][t] = 0.f;
          }
        }
      };
#pragma unroll
      for (int i = 0; i < PF; ++i) issue(i, kb_begin + i);
      for (int kb0 = kb_begin; kb0 < kb_end; kb0 += PF) {
#pragma unroll
        for (int i = 0; i < PF; ++i) {
          const int kb = kb0 + i;
          if (kb < kb_end) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (kb < p.kb1) {
              const uint32_t sb0 = smem_base + stage * C::STAGE_BYTES + A_TILE_BYTES;
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                uint32_t row_addr;
                int rsw;
                if constexpr (!B_MN) {
                  row_addr = sb0 + t * C::B_TILE_BYTES + dt * 128;
                  rsw = dt & 7;
                } else {
                  row_addr = sb0 + t * C::B_TILE_BYTES + (dt & 1) * 8192 + (dt >> 1) * 128;
                  rsw = (dt >> 1) & 7;
                }
                nf4_row_to_smem(row_addr, rsw, q[i][t][0], q[i][t][1], am[i][t], s_lut);
              }
              fence_proxy_async_smem();   // generic-proxy smem writes -> the UMMA's async-proxy reads
            }
            arrive_full(stage);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            issue(i, kb + PF);            // refill this slot: codes of the k-block PF steps ahead
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + EW) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;  // 0, or 1 for the second warp of a quadrant (EW == 8)
    constexpr int NH = EW / 4;          // column shares
    int local = 0, local1 = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++local) {
      if (unit < n_ext) {
        // ---- ext unit: U[rows of m-block `unit`, ext_n] = ext_alpha * accumulator, bf16; then raise this warp's flag ----
        const int acc = NT == 1 ? (local & 1) : 0;
        const uint32_t ph = NT == 1 ? ((local >> 1) & 1) : (local & 1);
        mbar_wait(&tmem_full_bar[acc], ph);
        tc_fence_after();
        if (half == 0) {
          const int row = unit * BM2 + (int)rank * BM + quad * 32 + lane;
          const uint32_t taddr0 = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
          for (int c = 0; c < p.ext_n / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr0 + c * 32, r);
            tmem_ld_wait();
            if (row < p.M) {
              bf16* dst = p.ext_out + (long long)row * p.ld_ext + c * 32;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * p.ext_alpha;
                *reinterpret_cast<bf16x8*>(dst + g * 8) = pack8(v);
              }
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // U is read back through TMA (async proxy) by the consumers
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_gpu(p.ext_flags + (unit * 2 + (int)rank) * 4 + quad, p.ext_epoch);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty_bar[acc]);
          else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
        }
        continue;
      }
      int tile, kb_begin, kb_end, part, sidx, m_blk, n_blk;
      unit_decode(p, unit - n_ext, kb_total, tile, kb_begin, kb_end, part, sidx);
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, p.gm, m_blk, n_blk);
      const int row = m_blk * BM2 + (int)rank * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      if (part == 0) {
        // K-range 0 of a tail tile: wait for the same warp (rank, quad[, half]) of every other K-range of this tile
        if (lane == 0) {
          for (int s = 1; s < p.tail_split; ++s) {
            const int* f = NT == 2 ? p.tail_flags + ((sidx * p.tail_split + s) * 8 + rank * 4 + quad) * 2 + half
                                   : p.tail_flags + (sidx * p.tail_split + s) * 8 + rank * 4 + quad;
            long long t0 = clock64();
            while (ld_acquire_gpu(f) != p.tail_epoch) {
              if (clock64() - t0 > 40000000000LL) {
                printf("b200rl: gemm tail-split flag wait timed out (block %d)\n", blockIdx.x);
                __trap();
              }
            }
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
      const int acc = NT == 1 ? (local & 1) : t;
      const uint32_t acc_phase = NT == 1 ? ((local >> 1) & 1) : ((t == 0 ? local : local1) & 1);
      const int n_sub = n_blk * NT + t;   // 256-column sub-tile index along N (== n_blk for NT = 1)
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
      if constexpr (NT == 2) {
        // ---- wide tiles: registers are the second accumulator buffer ----
        // Both TMEM accumulators belong to this tile, so the next tile's MMAs can only start once they are drained.  Each
        // of the 8 epilogue warps therefore pulls ITS 128 accumulator columns of this sub-tile into registers first
        // (4 x tcgen05.ld 32x32b.x32), releases the accumulator (tmem_empty arrive: ~300 cycles after the commit instead of
        // the ~2-3 k cycles of the conversion + global stores) and only then converts / stores from registers.
        static_assert(EW == 8, "wide tiles use 8 epilogue warps: two per TMEM lane quadrant, 128 columns each");
        uint32_t r[4][32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // FUSE 1: gate columns [64 half, 64 half + 64) and the up columns with the same index (accumulator cols + 128)
          const int col = FUSE == 1 ? (j < 2 ? half * 64 + j * 32 : 128 + half * 64 + (j - 2) * 32) : half * 128 + j * 32;
          tmem_ld_32x32(taddr0 + col, r[j]);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tmem_empty_bar[acc]);
          else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
        }
        if constexpr (FUSE == 1) {
          if (n_sub * 128 < p.fuse_I && row_ok) {
            bf16* gu_row = reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc;
            bf16* act_row = p.aux_out + (long long)row * p.ld_aux;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = n_sub * 128 + half * 64 + j * 32 + g * 8;
                float a[8], b[8], o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  a[i] = __uint_as_float(r[j][g * 8 + i]) * p.alpha;
                  b[i] = __uint_as_float(r[j + 2][g * 8 + i]) * p.alpha;
                }
                const bf16x8 ga = pack8(a), ub = pack8(b);
                *reinterpret_cast<bf16x8*>(gu_row + col) = ga;
                *reinterpret_cast<bf16x8*>(gu_row + p.fuse_I + col) = ub;
                unpack8(ga, a);  // the activation is computed from the bf16-rounded gate / up, like the row kernel
                unpack8(ub, b);
                swiglu_fwd8(a, b, o);
                *reinterpret_cast<bf16x8*>(act_row + col) = pack8(o);
              }
            }
          }
        } else if constexpr (FUSE == 2) {
          if (row_ok) {
            const bf16* gu_row = p.aux_in + (long long)row * p.ld_aux;
            bf16* dgu_row = reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = n_sub * BN + half * 128 + j * 32 + g * 8;
                if (col < p.N) {
                  float a[8], b[8], d[8], og[8], ou[8];
                  unpack8(*reinterpret_cast<const bf16x8*>(gu_row + col), a);
                  unpack8(*reinterpret_cast<const bf16x8*>(gu_row + p.fuse_I + col), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(r[j][g * 8 + i]) * p.alpha;
                  unpack8(pack8(d), d);  // dact is a bf16 tensor in the unfused path
                  swiglu_bwd8(a, b, d, og, ou);
                  *reinterpret_cast<bf16x8*>(dgu_row + col) = pack8(og);
                  *reinterpret_cast<bf16x8*>(dgu_row + p.fuse_I + col) = pack8(ou);
                }
              }
            }
          }
        } else if (part > 0) {
          // K-range 1..S-1 of a tail tile: raw fp32 accumulators -> workspace; flag after the last sub-tile
          float* ws = p.tail_ws + (((long long)(sidx * (p.tail_split - 1) + (part - 1)) * 2 + rank) * NT + t) * (BM * BN);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = half * 4 + j;
            float4* dst = reinterpret_cast<float4*>(ws + ((c * 4 + quad) * 32 + lane) * 32);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              dst[g] = make_float4(__uint_as_float(r[j][4 * g]), __uint_as_float(r[j][4 * g + 1]),
                                   __uint_as_float(r[j][4 * g + 2]), __uint_as_float(r[j][4 * g + 3]));
          }
          if (t == NT - 1) {
            __threadfence();
            __syncwarp();
            // 8 warps: the two warps of a quadrant share one flag slot pair -> one flag per (rank, quad, half)
            if (lane == 0) st_release_gpu(p.tail_flags + ((sidx * p.tail_split + part) * 8 + rank * 4 + quad) * 2 + half, p.tail_epoch);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = half * 4 + j;
            if (part == 0) {
              for (int s2 = 1; s2 < p.tail_split; ++s2) {
                const float4* src = reinterpret_cast<const float4*>(
                    p.tail_ws + (((long long)(sidx * (p.tail_split - 1) + (s2 - 1)) * 2 + rank) * NT + t) * (BM * BN) +
                    ((c * 4 + quad) * 32 + lane) * 32);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  const float4 v = __ldcg(src + g);
                  r[j][4 * g] = __float_as_uint(__uint_as_float(r[j][4 * g]) + v.x);
                  r[j][4 * g + 1] = __float_as_uint(__uint_as_float(r[j][4 * g + 1]) + v.y);
                  r[j][4 * g + 2] = __float_as_uint(__uint_as_float(r[j][4 * g + 2]) + v.z);
                  r[j][4 * g + 3] = __float_as_uint(__uint_as_float(r[j][4 * g + 3]) + v.w);
                }
              }
            }
            if (row_ok) epilogue_store32(p, r[j], row, n_sub * BN + c * 32, 0);
          }
        }
        continue;
      }
      if constexpr (FUSE == 1) {
        if (n_sub * 128 >= p.fuse_I) {   // second sub-tile of the last wide tile when I / 128 is odd: nothing to store
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(&tmem_empty_bar[acc]);
            else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
          }
          continue;
        }
        // accumulator columns [0,128) = gate(j0..), [128,256) = up(j0..) with j0 = n_sub * 128
        bf16* gu_row = reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc;
        bf16* act_row = p.aux_out + (long long)row * p.ld_aux;
#pragma unroll 1
        for (int c = half * (4 / NH); c < (half + 1) * (4 / NH); ++c) {
          uint32_t rg[32], ru[32];
          tmem_ld_32x32(taddr0 + c * 32, rg);
          tmem_ld_32x32(taddr0 + 128 + c * 32, ru);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = n_sub * 128 + c * 32 + g * 8;
              float a[8], b[8], o[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                a[i] = __uint_as_float(rg[g * 8 + i]) * p.alpha;
                b[i] = __uint_as_float(ru[g * 8 + i]) * p.alpha;
              }
              const bf16x8 ga = pack8(a), ub = pack8(b);
              *reinterpret_cast<bf16x8*>(gu_row + col) = ga;
              *reinterpret_cast<bf16x8*>(gu_row + p.fuse_I + col) = ub;
              unpack8(ga, a);  // the activation is computed from the bf16-rounded gate / up, like the row kernel
              unpack8(ub, b);
              swiglu_fwd8(a, b, o);
              *reinterpret_cast<bf16x8*>(act_row + col) = pack8(o);
            }
          }
        }
      } else if constexpr (FUSE == 2) {
        const bf16* gu_row = p.aux_in + (long long)row * p.ld_aux;
        bf16* dgu_row = reinterpret_cast<bf16*>(p.C) + (long long)row * p.ldc;
#pragma unroll 1
        for (int c = half * (BN / 32 / NH); c < (half + 1) * (BN / 32 / NH); ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + c * 32, r);
          const int colc = n_sub * BN + c * 32;
          bf16x8 gq[4], uq[4];
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (colc + g * 8 < p.N) {
                gq[g] = *reinterpret_cast<const bf16x8*>(gu_row + colc + g * 8);
                uq[g] = *reinterpret_cast<const bf16x8*>(gu_row + p.fuse_I + colc + g * 8);
              }
          }
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = colc + g * 8;
              if (col < p.N) {
                float a[8], b[8], d[8], og[8], ou[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(r[g * 8 + i]) * p.alpha;
                unpack8(pack8(d), d);  // dact is a bf16 tensor in the unfused path
                unpack8(gq[g], a);
                unpack8(uq[g], b);
                swiglu_bwd8(a, b, d, og, ou);
                *reinterpret_cast<bf16x8*>(dgu_row + col) = pack8(og);
                *reinterpret_cast<bf16x8*>(dgu_row + p.fuse_I + col) = pack8(ou);
              }
            }
          }
        }
      } else
      if (part > 0) {
        // K-range 1..S-1 of a tail tile: raw fp32 accumulators -> workspace, then raise this warp's flag
        float* ws = p.tail_ws + (((long long)(sidx * (p.tail_split - 1) + (part - 1)) * 2 + rank) * NT + t) * (BM * BN);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + c * 32, r);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(ws + ((c * 4 + quad) * 32 + lane) * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            dst[g] = make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                                 __uint_as_float(r[4 * g + 3]));
        }
        if (t == NT - 1) {   // every sub-tile of this K-range is in the workspace: raise this warp's flag
          __threadfence();
          __syncwarp();
          if (lane == 0) st_release_gpu(p.tail_flags + (sidx * p.tail_split + part) * 8 + rank * 4 + quad, p.tail_epoch);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + c * 32, r);
          tmem_ld_wait();
          if (part == 0) {
            for (int s = 1; s < p.tail_split; ++s) {
              const float4* src = reinterpret_cast<const float4*>(
                  p.tail_ws + (((long long)(sidx * (p.tail_split - 1) + (s - 1)) * 2 + rank) * NT + t) * (BM * BN) +
                  ((c * 4 + quad) * 32 + lane) * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 v = __ldcg(src + g);
                r[4 * g] = __float_as_uint(__uint_as_float(r[4 * g]) + v.x);
                r[4 * g + 1] = __float_as_uint(__uint_as_float(r[4 * g + 1]) + v.y);
                r[4 * g + 2] = __float_as_uint(__uint_as_float(r[4 * g + 2]) + v.z);
                r[4 * g + 3] = __float_as_uint(__uint_as_float(r[4 * g + 3]) + v.w);
              }
            }
          }
          if (row_ok) epilogue_store32(p, r, row, n_sub * BN + c * 32, 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
      }
      }  // sub-tile loop
      ++local1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's smem / TMEM stay valid until the leader's last MMA has retired
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C::TMEM_COLS);
  }
}

static int g_pair_enabled = -1;
bool gemm_pair_enabled() {
  if (g_pair_enabled < 0) {
    const char* e = getenv("B200RL_GEMM_CTA_PAIR");
    g_pair_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pair_enabled != 0;
}

// ---- tail-split workspace: one per (device, stream), grown lazily; flags compare against a per-workspace epoch ----
struct TailWs {
  float* ws = nullptr;
  int* flags = nullptr;
  size_t ws_bytes = 0;
  int epoch = 0;
  int* ext_flags = nullptr;   // ext units: [EXT_MAX_MBLOCKS][2][4]
  int ext_epoch = 0;
};
static constexpr int EXT_MAX_MBLOCKS = 4096;   // 1M rows per launch
static std::mutex g_tail_mu;
static std::map<std::pair<int, cudaStream_t>, TailWs> g_tail_ws;
static int g_tail_enabled = -1;
static constexpr int TAIL_MAX_FLAGS = 128 * 4 * 16;  // <= 128 tail tiles x 4 parts x 16 warps (8 per CTA with wide tiles)

void gemm_pair_set_tail_split(int enable) { g_tail_enabled = enable ? 1 : 0; }

static int tail_workspace(cudaStream_t stream, size_t bytes, TailWs** out) {
  int dev = 0;
  B200RL_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tail_mu);
  TailWs& w = g_tail_ws[std::make_pair(dev, stream)];
  if (bytes > 0 && w.ws_bytes < bytes) {
    if (w.ws) B200RL_CUDA_OK(cudaFree(w.ws));  // synchronises: no kernel still uses the old buffer
    w.ws = nullptr;
    w.ws_bytes = 0;
    B200RL_CUDA_OK(cudaMalloc(&w.ws, bytes));
    w.ws_bytes = bytes;
  }
  if (!w.flags) {
    B200RL_CUDA_OK(cudaMalloc(&w.flags, TAIL_MAX_FLAGS * sizeof(int)));
    B200RL_CUDA_OK(cudaMemset(w.flags, 0, TAIL_MAX_FLAGS * sizeof(int)));
  }
  if (!w.ext_flags) {
    B200RL_CUDA_OK(cudaMalloc(&w.ext_flags, EXT_MAX_MBLOCKS * 8 * sizeof(int)));
    B200RL_CUDA_OK(cudaMemset(w.ext_flags, 0, EXT_MAX_MBLOCKS * 8 * sizeof(int)));
  }
  *out = &w;
  return 0;
}

template <int BN, bool B_MN, int FUSE = 0, int EW = 4, int NT = 1, bool NF4 = false>
static int launch_pair(const GemmArgs& a, cudaStream_t stream) {
  using C = PairCfg<BN, B_MN, NT>;
  GemmParams p;
  p.M = a.M;
  p.N = a.N;
  p.kb1 = (a.K1 + BK - 1) / BK;
  p.kb2 = (a.K2 + BK - 1) / BK;
  p.num_m_blocks = (a.M + 2 * BM - 1) / (2 * BM);
  p.num_n_blocks = FUSE == 1 ? (a.N / BN + NT - 1) / NT : (a.N + BN * NT - 1) / (BN * NT);
  p.splits = 1;
  p.kb_per_split = p.kb1 + p.kb2;
  p.C = a.C;
  p.ldc = a.ldc;
  p.c_split_stride = 0;
  p.c_fp32 = a.c_fp32;
  p.bias = reinterpret_cast<const bf16*>(a.bias);
  p.residual = reinterpret_cast<const bf16*>(a.residual);
  p.ldr = a.ldr;
  p.alpha = a.alpha;
  p.gm = raster_group((long long)a.M * (a.K1 + a.K2) * 2, p.num_m_blocks);
  if (FUSE) {
    p.fuse_I = FUSE == 1 ? a.N / 2 : a.N;
    p.aux_in = reinterpret_cast<const bf16*>(a.aux);
    p.aux_out = reinterpret_cast<bf16*>(a.aux);
    p.ld_aux = a.ld_aux;
  }
  CUtensorMap tA1, tB1, tA2, tB2, tExt;
  int rc;
  if ((rc = make_map(&tA1, a.A1, a.K1, a.M, a.lda1, BK, BM))) return rc;
  if constexpr (NF4) {
    B200RL_REQUIRE(a.nf4_packed && a.nf4_absmax && a.K1 % 64 == 0 && a.N % 64 == 0,
                   "gemm(nf4): needs packed codes + absmax, K1 %% 64 == 0 and N %% 64 == 0 (K1=%d N=%d)", a.K1, a.N);
    p.nf4_q = reinterpret_cast<const uint8_t*>(a.nf4_packed);
    p.nf4_am = a.nf4_absmax;
    p.K1 = a.K1;
    tB1 = tA1;   // unused: the dequant warps write the B tiles
  } else {
    if ((rc = B_MN ? make_map(&tB1, a.B1, a.N, a.K1, a.ldb1, 64, BK) : make_map(&tB1, a.B1, a.K1, a.N, a.ldb1, BK, C::BH))) return rc;
  }
  if (a.K2 > 0) {
    if ((rc = make_map(&tA2, a.A2, a.K2, a.M, a.lda2, BK, BM))) return rc;
    if ((rc = B_MN ? make_map(&tB2, a.B2, a.N, a.K2, a.ldb2, 64, BK) : make_map(&tB2, a.B2, a.K2, a.N, a.ldb2, BK, C::BH))) return rc;
  } else {
    tA2 = tA1;
    tB2 = tB1;
  }
  const bool ext = a.ext_B != nullptr;
  if (ext) {
    B200RL_REQUIRE(a.K2 == 64 || a.K2 == 128, "gemm(ext): the LoRA intermediate must be 64 or 128 wide (K2=%d)", a.K2);
    B200RL_REQUIRE(p.num_m_blocks <= EXT_MAX_MBLOCKS, "gemm(ext): too many row blocks (%d)", p.num_m_blocks);
    // forward: Acat [K2, K1] K-major, each CTA stages K2/2 rows; dX form: Bcat [K1, K2] MN-major, 64-column slabs
    if ((rc = B_MN ? make_map(&tExt, a.ext_B, a.K2, a.K1, a.ld_ext_b, 64, BK)
                   : make_map(&tExt, a.ext_B, a.K1, a.K2, a.ld_ext_b, BK, a.K2 / 2))) return rc;
    p.n_ext = p.num_m_blocks;
    p.ext_n = a.K2;
    p.ext_out = reinterpret_cast<bf16*>(const_cast<void*>(a.A2));
    p.ld_ext = a.lda2;
    p.ext_alpha = a.ext_alpha;
  } else {
    tExt = tB1;
  }
  auto kern = gemm_pair_kernel<BN, B_MN, FUSE, EW, NT, NF4>;
  static DeviceOnce attr_once;   // per template instantiation, per device
  if (attr_once.first())
    B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int slots = tiles + p.n_ext;   // work units before any tail split
  int clusters = num_sms() / 2;
  if (a.max_ctas > 0 && a.max_ctas / 2 < clusters) clusters = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  if (slots < clusters) clusters = slots;
  if (g_tail_enabled < 0) {
    const char* e = getenv("B200RL_GEMM_TAIL_SPLIT");
    g_tail_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int kb_total = p.kb1 + p.kb2;
  // the K-ranges of a tail tile must be co-resident (range 0 waits for the others): with ext units in front, the last
  // round holds (tiles + n_ext) mod clusters units
  int S = (g_tail_enabled && FUSE == 0) ? pair_tail_split(slots, clusters, kb_total) : 1;
  if (S > 1 && slots % clusters > tiles) S = 1;
  TailWs* wx = nullptr;
  if (ext) {
    int rc2 = tail_workspace(stream, 0, &wx);
    if (rc2) return rc2;
    p.ext_flags = wx->ext_flags;
    p.ext_epoch = ++wx->ext_epoch;
  }
  if (S > 1) {
    const int rem = slots % clusters;
    TailWs* w = nullptr;
    int rc2 = tail_workspace(stream, (size_t)rem * (S - 1) * 2 * NT * BM * BN * sizeof(float), &w);
    if (rc2) return rc2;
    p.tail_first = tiles - rem;
    p.tail_split = S;
    p.tail_ws = w->ws;
    p.tail_flags = w->flags;
    p.tail_epoch = ++w->epoch;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(128 + 32 * EW + (NF4 ? 128 : 0));
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 2 : 1;
  B200RL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tA1, tB1, tA2, tB2, tExt, p));
  B200RL_LAUNCH_OK();
  return 0;
}

// Called by gemm_dispatch for A K-major layouts (TN / dX) without split-K; bn in {128, 192, 224, 256}.
static int fuse_epilogue_warps() {
  static int ew = 0;
  if (!ew) {
    const char* e = getenv("B200RL_GEMM_FUSE_EW");
    ew = (e && e[0] == '4') ? 4 : 8;
  }
  return ew;
}

bool gemm_fuse_supported(int M, int I) { return gemm_pair_enabled() && M > BM && I % 128 == 0; }

// LoRA-in-kernel: -1 = env B200RL_GEMM_EXT (default on), 0 / 1 forced by b200rl_gemm_set_ext (tests, A/B runs)
static int g_ext = -1;
bool gemm_ext_supported(int M, int N, int K2) {
  if (g_ext < 0) {
    const char* e = getenv("B200RL_GEMM_EXT");
    g_ext = (e && e[0] == '0') ? 0 : 1;
  }
  // needs the CTA-pair kernel (same condition as gemm_dispatch) and a 64- or 128-wide intermediate
  return g_ext != 0 && gemm_pair_enabled() && M > BM && N >= 256 && (K2 == 64 || K2 == 128);
}

// Wide (256 x 512) tiles: -1 = env B200RL_GEMM_WIDE (default 2), else forced by b200rl_gemm_set_wide (tests, A/B runs).
// Measured on B200 (profiles/r2_run02_wide_ext_ab.txt, r2_run05_wide_modes_ab.txt), config 2, ms per learner step:
//   first version (epilogue drains TMEM while the tensor pipe waits): mode 1 1061 vs mode 0 1033 — slower although the SM
//   clock under the power cap rose from 1522 to 1620 MHz (fewer operand bytes per flop);
//   with the early-release epilogue (accumulator -> registers, free TMEM, then convert / store):
//   mode 0 1031 / 1039, mode 1 1029, mode 2 1016  -> default 2: wide tiles for the K-long GEMMs only.
// Modes: 0 = never, 1 = wherever every pair gets a wide tile, 2 = additionally only for K-long GEMMs (>= 256 k-blocks:
// down forward, gate|up dX, lm_head dX), where the un-overlapped epilogue is < 3 % of a tile and the DRAM re-reads of the
// 256 x 256 schedule are largest.
static int g_wide = -1;
static int wide_mode() {
  if (g_wide < 0) {
    const char* e = getenv("B200RL_GEMM_WIDE");
    g_wide = e ? atoi(e) : 2;
    if (g_wide < 0 || g_wide > 2) g_wide = 2;
  }
  return g_wide;
}
bool gemm_pair_wide_enabled() { return wide_mode() != 0; }
bool gemm_pair_wide_for(int M, int n_cols, int k_total) {
  const int mode = wide_mode();
  if (mode == 0 || n_cols < 512) return false;
  static int min_kb = -1;   // B200RL_GEMM_WIDE_MIN_KB: k-blocks from which a GEMM counts as K-long (A/B knob)
  if (min_kb < 0) {
    const char* e = getenv("B200RL_GEMM_WIDE_MIN_KB");
    min_kb = e ? atoi(e) : 256;
    if (min_kb <= 0) min_kb = 256;
  }
  if (mode == 2 && (k_total + BK - 1) / BK < min_kb) return false;
  const long long mb = (M + 2 * BM - 1) / (2 * BM);
  return mb * ((n_cols + 511) / 512) >= num_sms() / 2;
}
static bool want_wide(const GemmArgs& a, int n_cols) { return gemm_pair_wide_for(a.M, n_cols, a.K1 + a.K2); }

// NF4 in the mainloop: same conditions as the CTA-pair kernel itself, 256-column (sub-)tiles only
bool gemm_nf4_supported(int M, int N, int K1) {
  return gemm_pair_enabled() && M > BM && N >= 256 && N % 64 == 0 && K1 % 64 == 0;
}

static int gemm_pair_dispatch_nf4(const GemmArgs& a, int bn, cudaStream_t stream) {
  const bool b_mn = (a.mn_major & 2) != 0;
  if (a.fuse == 1) {
    B200RL_REQUIRE(!b_mn && !a.c_fp32 && !a.bias && !a.residual && a.aux && a.N % 256 == 0 && a.ld_aux % 8 == 0,
                   "gemm(fused swiglu fwd): needs TN layout, bf16 C, N = 2I with I %% 128 == 0, no bias/residual");
    return want_wide(a, a.N) ? launch_pair<256, false, 1, 8, 2, true>(a, stream) : launch_pair<256, false, 1, 8, 1, true>(a, stream);
  }
  if (a.fuse == 2) {
    B200RL_REQUIRE(b_mn && !a.c_fp32 && !a.bias && !a.residual && a.aux && a.N % 8 == 0 && a.ld_aux % 8 == 0,
                   "gemm(fused swiglu bwd): needs dX layout, bf16 C, no bias/residual");
    return want_wide(a, a.N) ? launch_pair<256, true, 2, 8, 2, true>(a, stream) : launch_pair<256, true, 2, 8, 1, true>(a, stream);
  }
  if (bn == 512) return b_mn ? launch_pair<256, true, 0, 8, 2, true>(a, stream) : launch_pair<256, false, 0, 8, 2, true>(a, stream);
  return b_mn ? launch_pair<256, true, 0, 4, 1, true>(a, stream) : launch_pair<256, false, 0, 4, 1, true>(a, stream);
}

int gemm_pair_dispatch(const GemmArgs& a, int bn, cudaStream_t stream) {
  const bool b_mn = (a.mn_major & 2) != 0;
  if (a.nf4_packed) return gemm_pair_dispatch_nf4(a, bn, stream);
  if (a.fuse == 1) {
    B200RL_REQUIRE(!b_mn && !a.c_fp32 && !a.bias && !a.residual && a.aux && a.N % 256 == 0 && a.ld_aux % 8 == 0,
                   "gemm(fused swiglu fwd): needs TN layout, bf16 C, N = 2I with I %% 128 == 0, no bias/residual");
    if (want_wide(a, a.N)) return launch_pair<256, false, 1, 8, 2>(a, stream);
    return fuse_epilogue_warps() == 8 ? launch_pair<256, false, 1, 8>(a, stream) : launch_pair<256, false, 1, 4>(a, stream);
  }
  if (a.fuse == 2) {
    B200RL_REQUIRE(b_mn && !a.c_fp32 && !a.bias && !a.residual && a.aux && a.N % 8 == 0 && a.ld_aux % 8 == 0,
                   "gemm(fused swiglu bwd): needs dX layout, bf16 C, no bias/residual");
    if (want_wide(a, a.N)) return launch_pair<256, true, 2, 8, 2>(a, stream);
    return fuse_epilogue_warps() == 8 ? launch_pair<256, true, 2, 8>(a, stream) : launch_pair<256, true, 2, 4>(a, stream);
  }
  if (bn == 512) return b_mn ? launch_pair<256, true, 0, 8, 2>(a, stream) : launch_pair<256, false, 0, 8, 2>(a, stream);
  if (bn == 256) return b_mn ? launch_pair<256, true>(a, stream) : launch_pair<256, false>(a, stream);
  if (bn == 224) return b_mn ? launch_pair<224, true>(a, stream) : launch_pair<224, false>(a, stream);
  if (bn == 192) return b_mn ? launch_pair<192, true>(a, stream) : launch_pair<192, false>(a, stream);
  if (bn == 128) return b_mn ? launch_pair<128, true>(a, stream) : launch_pair<128, false>(a, stream);
  return set_error(B200RL_ERR_UNSUPPORTED, "gemm(pair): BN=%d not instantiated", bn);
}

}  // namespace b200rl

// test / bisection switch: 1 = use CTA-pair kernels where applicable (default), 0 = single-CTA only
extern "C" int b200rl_gemm_set_cta_pair(int enable) {
  b200rl::g_pair_enabled = enable ? 1 : 0;
  return 0;
}
// test / A-B switch for the LoRA-in-kernel ext units (model driver): 1 = default, 0 = separate skinny GEMM + reduce
extern "C" int b200rl_gemm_set_ext(int enable) {
  b200rl::g_ext = enable ? 1 : 0;
  return 0;
}
// test / A-B switch for the wide (256 x 512) pair tiles: 1 = use them where applicable (default), 0 = 256 x 256 only
namespace b200rl { int g_gemm_raster_forced = 0; }
extern "C" int b200rl_gemm_set_raster(int gm) {
  b200rl::g_gemm_raster_forced = gm > 0 ? gm : 0;
  return 0;
}

extern "C" int b200rl_gemm_set_wide(int mode) {
  b200rl::g_wide = (mode < 0 || mode > 2) ? 0 : mode;
  return 0;
}
// test / bisection switch for the K-split of the last partial wave (default on; env B200RL_GEMM_TAIL_SPLIT=0 disables)
extern "C" int b200rl_gemm_set_tail_split(int enable) {
  b200rl::gemm_pair_set_tail_split(enable);
  return 0;
}
