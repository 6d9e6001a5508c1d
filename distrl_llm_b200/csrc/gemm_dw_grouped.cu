// G1 (grouped dW form): all LoRA weight-gradient GEMMs of one decoder layer in ONE persistent tcgen05 launch.
//
// The backward of a layer needs, for each of its four projection groups (qkv, o, gate|up, down), two skinny GEMMs
// with the reduction over tokens (reference: autograd of the PEFT LoRA linear, reached through loss.backward(),
// distributed_actor.py:385 / :483):
//     dBcat [out, 64] = dY^T . u          dAcat^T [in, 64] = x^T . du
// Each is HBM-bound (it streams an [tokens, out|in] activation once) and far too small to fill 148 SMs on its own;
// as 8 separate launches they also pay 8 launch/drain gaps.  Here the (problem, 128-row tile, K-range) work units of
// all 8 problems are spread round-robin over one persistent grid.  Both operands are read as stored (MN-major UMMA
// descriptors, like gemm_kernel<64, true, true>), results go to per-problem fp32 slabs [splits][rows][64]
// (deterministic: the caller sums the K-ranges in fixed order).
#include "gemm_common.cuh"

namespace b200rl {

static constexpr int DW_MAX = 8;
static constexpr int DW_BN = 64;
static constexpr int DW_STAGE_BYTES = A_TILE_BYTES + DW_BN * BK * 2;  // 16 KB + 8 KB
static constexpr int DW_STAGES = 8;
static constexpr int DW_SMEM_BYTES = DW_STAGES * DW_STAGE_BYTES + 1024;
static constexpr int DW_ACC_STRIDE = 64;

struct DwMaps {
  CUtensorMap y[DW_MAX];  // Y [tokens][rows]  -> dims {rows, tokens}, box {64, 64}
  CUtensorMap u[DW_MAX];  // U [tokens][64]    -> dims {64, tokens},   box {64, 64}
};
struct DwParams {
  float* C[DW_MAX];
  long long split_stride[DW_MAX];
  int rows[DW_MAX];
  int m_blocks[DW_MAX];
  int unit0[DW_MAX + 1];  // first work unit of each problem
  int nprob;
  int splits, kb_total, kb_per_split;
};

__device__ __forceinline__ void dw_decode(const DwParams& p, int unit, int& prob, int& m_blk, int& split) {
  prob = 0;
#pragma unroll 1
  while (prob + 1 < p.nprob && unit >= p.unit0[prob + 1]) ++prob;
  const int local = unit - p.unit0[prob];
  split = local / p.m_blocks[prob];
  m_blk = local - split * p.m_blocks[prob];
}

__global__ void __launch_bounds__(256, 1)
dw_grouped_kernel(const __grid_constant__ DwMaps maps, const DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[DW_STAGES], empty_bar[DW_STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < DW_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, 2 * DW_ACC_STRIDE);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();

  const int num_units = p.unit0[p.nprob];
  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int prob, m_blk, split;
      dw_decode(p, unit, prob, m_blk, split);
      const CUtensorMap* ty = &maps.y[prob];
      const CUtensorMap* tu = &maps.u[prob];
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem_gen + stage * DW_STAGE_BYTES;
        uint8_t* sb = sa + A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], DW_STAGE_BYTES);
        // both operands MN-major (stored [token][mn]): boxes of 64 (mn) x 64 (tokens), 8 KB each
        tma_load_2d(sa, ty, &full_bar[stage], m_blk * BM, kb * BK);
        tma_load_2d(sa + 8192, ty, &full_bar[stage], m_blk * BM + 64, kb * BK);
        tma_load_2d(sb, tu, &full_bar[stage], 0, kb * BK);
        if (++stage == DW_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(DW_BN, true, true);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
      int prob, m_blk, split;
      dw_decode(p, unit, prob, m_blk, split);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * DW_ACC_STRIDE;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * DW_STAGE_BYTES;
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 2048, 8192, 1024);
          const uint64_t db = make_smem_desc(sb + k * 2048, 8192, 1024);
          umma_bf16(tmem_d, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == DW_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tmem_full_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue: fp32 accumulators -> this K-range's slab =====================
    const int quad = warp & 3;
    int local = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
      int prob, m_blk, split;
      dw_decode(p, unit, prob, m_blk, split);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.rows[prob];
      const uint32_t taddr0 = tmem_base + acc * DW_ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
      float* dst = p.C[prob] + (long long)split * p.split_stride[prob] + (long long)row * DW_BN;
#pragma unroll
      for (int c = 0; c < DW_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr0 + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(dst + c * 32 + g * 4) =
                make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                            __uint_as_float(r[4 * g + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * DW_ACC_STRIDE);
  }
}

// K-split factor shared by all problems of a grouped launch: fewest "waves x range length".
int dw_grouped_splits(int total_m_blocks, int kb_total) {
  const int sms = num_sms();
  int best_s = 1;
  double best = 1e30;
  for (int s = 1; s <= 4 && s * 8 <= kb_total; ++s) {
    const long long units = (long long)total_m_blocks * s;
    const double t = (double)((units + sms - 1) / sms) / s;
    if (t < best - 1e-9) {
      best = t;
      best_s = s;
    }
  }
  return best_s;
}

// Y_i [tokens, rows_i] (ld ldy_i), U_i [tokens, 64] (ld ldu_i), C_i fp32 [splits][rows_i][64] (split_stride_i apart).
int dw_grouped_dispatch(int nprob, const void* const* Y, const long long* ldy, const int* rows, const void* const* U,
                        const long long* ldu, float* const* C, const long long* split_stride, int tokens, int splits,
                        cudaStream_t stream) {
  B200RL_REQUIRE(nprob >= 1 && nprob <= DW_MAX && tokens > 0 && splits >= 1, "dw_grouped: bad problem count / shape");
  DwMaps maps;
  DwParams p;
  memset(&p, 0, sizeof(p));
  p.nprob = nprob;
  p.kb_total = (tokens + BK - 1) / BK;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  int unit = 0, rc;
  for (int i = 0; i < nprob; ++i) {
    B200RL_REQUIRE(rows[i] > 0 && rows[i] % 8 == 0 && Y[i] && U[i] && C[i], "dw_grouped: bad problem %d", i);
    if ((rc = make_map(&maps.y[i], Y[i], rows[i], tokens, ldy[i], 64, BK))) return rc;
    if ((rc = make_map(&maps.u[i], U[i], DW_BN, tokens, ldu[i], 64, BK))) return rc;
    p.C[i] = C[i];
    p.split_stride[i] = split_stride[i];
    p.rows[i] = rows[i];
    p.m_blocks[i] = (rows[i] + BM - 1) / BM;
    p.unit0[i] = unit;
    unit += p.m_blocks[i] * p.splits;
  }
  for (int i = nprob; i < DW_MAX; ++i) {
    maps.y[i] = maps.y[0];
    maps.u[i] = maps.u[0];
  }
  for (int i = nprob; i <= DW_MAX; ++i) p.unit0[i] = unit;
  static DeviceOnce attr_once;   // kernel attributes are per device
  if (attr_once.first())
    B200RL_CUDA_OK(cudaFuncSetAttribute(dw_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));
  int ctas = num_sms();
  if (unit < ctas) ctas = unit;
  B200RL_CUDA_OK(launch_pdl(dw_grouped_kernel, dim3(ctas), dim3(256), DW_SMEM_BYTES, stream, maps, p));
  B200RL_LAUNCH_OK();
  return p.splits;  // > 0: the split count actually used
}

}  // namespace b200rl

using namespace b200rl;

// C ABI (tests / other hosts): see include/b200rl.h
extern "C" int b200rl_gemm_dw_grouped(int nprob, const void* const* Y, const long long* ldy, const int* rows,
                                      const void* const* U, const long long* ldu, float* const* C,
                                      const long long* split_stride, int tokens, int splits, void* stream) {
  const int rc = dw_grouped_dispatch(nprob, Y, ldy, rows, U, ldu, C, split_stride, tokens, splits,
                                     reinterpret_cast<cudaStream_t>(stream));
  return rc;
}
