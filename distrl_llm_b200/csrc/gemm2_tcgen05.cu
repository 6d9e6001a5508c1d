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
//   tmem_empty[a] (leader, count 8)  : one arrive per epilogue warp of both CTAs (peer arrives remotely)
#include "gemm_common.cuh"
#include <stdlib.h>
#include <string.h>

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

template <int BN>
struct PairCfg {
  static constexpr int BH = BN / 2;                      // B columns staged per CTA
  static constexpr int B_TILE_BYTES = BH * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;  // per CTA
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int STAGES_RAW = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const GemmParams p) {
  using C = PairCfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int BM2 = 2 * BM;  // rows per pair tile
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
  }
  if (warp == 2) tmem_alloc2(&tmem_base_smem, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;  // num_m_blocks counts 256-row pair tiles
  const int kb_total = p.kb1 + p.kb2;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk;
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      const int row0 = m_blk * BM2 + (int)rank * BM;
      const int col0 = n_blk * BN + (int)rank * C::BH;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem_gen + stage * C::STAGE_BYTES;
        uint8_t* sb = sa + A_TILE_BYTES;
        const bool seg2 = kb >= p.kb1;
        const CUtensorMap* ta = seg2 ? &tmA2 : &tmA1;
        const CUtensorMap* tb = seg2 ? &tmB2 : &tmB1;
        const int k0 = (seg2 ? kb - p.kb1 : kb) * BK;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
        tma_load_2d_pair(sa, ta, &full_bar[stage], k0, row0);
        if constexpr (!B_MN) {
          tma_load_2d_pair(sb, tb, &full_bar[stage], k0, col0);
        } else {
#pragma unroll
          for (int h = 0; h < C::BH / 64; ++h)
            tma_load_2d_pair(sb + h * 8192, tb, &full_bar[stage], col0 + h * 64, k0);
        }
        if (!leader) mbar_arrive_remote(&full_bar[stage], 0);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    // instruction M = 256 (both CTAs), N = BN
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM2 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024)
                                   : make_smem_desc(sb + k * 32, 16, 1024);
          umma_bf16_pair(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit_pair(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_pair(&tmem_full_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int quad = warp & 3;
    int local = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++local) {
      int m_blk, n_blk;
      tile_coords(tile, p.num_m_blocks, p.num_n_blocks, m_blk, n_blk);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM2 + (int)rank * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr0 + c * 32, r);
        tmem_ld_wait();
        if (row_ok) epilogue_store32(p, r, row, n_blk * BN + c * 32, 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_remote(&tmem_empty_bar[acc], 0);
      }
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

template <int BN, bool B_MN>
static int launch_pair(const GemmArgs& a, cudaStream_t stream) {
  using C = PairCfg<BN>;
  GemmParams p;
  p.M = a.M;
  p.N = a.N;
  p.kb1 = (a.K1 + BK - 1) / BK;
  p.kb2 = (a.K2 + BK - 1) / BK;
  p.num_m_blocks = (a.M + 2 * BM - 1) / (2 * BM);
  p.num_n_blocks = (a.N + BN - 1) / BN;
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
  CUtensorMap tA1, tB1, tA2, tB2;
  int rc;
  if ((rc = make_map(&tA1, a.A1, a.K1, a.M, a.lda1, BK, BM))) return rc;
  if ((rc = B_MN ? make_map(&tB1, a.B1, a.N, a.K1, a.ldb1, 64, BK) : make_map(&tB1, a.B1, a.K1, a.N, a.ldb1, BK, C::BH))) return rc;
  if (a.K2 > 0) {
    if ((rc = make_map(&tA2, a.A2, a.K2, a.M, a.lda2, BK, BM))) return rc;
    if ((rc = B_MN ? make_map(&tB2, a.B2, a.N, a.K2, a.ldb2, 64, BK) : make_map(&tB2, a.B2, a.K2, a.N, a.ldb2, BK, C::BH))) return rc;
  } else {
    tA2 = tA1;
    tB2 = tB1;
  }
  auto kern = gemm_pair_kernel<BN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  int clusters = num_sms() / 2;
  if (a.max_ctas > 0 && a.max_ctas / 2 < clusters) clusters = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B200RL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tA1, tB1, tA2, tB2, p));
  B200RL_LAUNCH_OK();
  return 0;
}

// Called by gemm_dispatch for A K-major layouts (TN / dX) without split-K; bn in {128, 256}.
int gemm_pair_dispatch(const GemmArgs& a, int bn, cudaStream_t stream) {
  const bool b_mn = (a.mn_major & 2) != 0;
  if (bn == 256) return b_mn ? launch_pair<256, true>(a, stream) : launch_pair<256, false>(a, stream);
  if (bn == 128) return b_mn ? launch_pair<128, true>(a, stream) : launch_pair<128, false>(a, stream);
  return set_error(B200RL_ERR_UNSUPPORTED, "gemm(pair): BN=%d not instantiated", bn);
}

}  // namespace b200rl

// test / bisection switch: 1 = use CTA-pair kernels where applicable (default), 0 = single-CTA only
extern "C" int b200rl_gemm_set_cta_pair(int enable) {
  b200rl::g_pair_enabled = enable ? 1 : 0;
  return 0;
}
