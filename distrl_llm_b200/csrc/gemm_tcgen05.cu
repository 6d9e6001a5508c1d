// G1: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
// Replaces the cuBLAS / bitsandbytes / Unsloth-Triton matmuls that the reference reaches through
// `policy(...)` and `loss.backward()` (reference distributed_actor.py:241-243, :385, :483;
// SURVEY.md §2.1 rows K1, K6).
//
//   C[M,N] = alpha * ( A1[M,K1] . B1[N,K1]^T  +  A2[M,K2] . B2[N,K2]^T ) (+ bias[N]) (+ R[M,N])
//
// * Operands are bf16, accumulation fp32 in TMEM, output bf16 or fp32.
// * The second (A2,B2) segment is the LoRA side path: A2 = s*(X.A^T) (the rank-r intermediate),
//   B2 = LoRA B, so "base GEMM + LoRA" is one mainloop over K1+K2 with no extra pass over C.
// * Three operand layouts, selected per operand through the UMMA descriptors (no transposes in HBM):
//   TN (both K-major: activations x weights^T), "dX" (A K-major, B MN-major: dY . W with W as
//   stored [out][in]) and "dW" (both MN-major: C[N_out, r] = Y^T . U, reduction over tokens) with
//   optional split-K into fp32 partial slabs (deterministic: summed in fixed order by the caller).
//
// Structure (one CTA per SM, 256 threads):
//   warp 0   : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1   : MMA issuer     (one thread: tcgen05.mma kind::f16, M=128, N=BN, K=16 per instr)
//   warp 2   : TMEM allocator (2 accumulator stages so the epilogue overlaps the next mainloop)
//   warps 4-7: epilogue       (tcgen05.ld 32x32b -> registers -> alpha/bias/residual -> global)
#include "gemm_common.cuh"

namespace b200rl {

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int STAGES_RAW = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
            const GemmParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1024-byte aligned tile ring (SWIZZLE_128B atoms are 1024 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB2);
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_enter();

  const int num_tiles = p.num_m_blocks * p.num_n_blocks * p.splits;
  const int kb_total = p.kb1 + p.kb2;

  if (warp == 0 && elect_one_sync()) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = p.num_m_blocks * p.num_n_blocks;
      const int split = tile / mn;
      int m_blk, n_blk;
      tile_coords(tile - split * mn, p.num_m_blocks, p.num_n_blocks, p.gm, m_blk, n_blk);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(kb_begin + p.kb_per_split, kb_total);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem_gen + stage * C::STAGE_BYTES;
        uint8_t* sb = sa + A_TILE_BYTES;
        const bool seg2 = kb >= p.kb1;
        const CUtensorMap* ta = seg2 ? &tmA2 : &tmA1;
        const CUtensorMap* tb = seg2 ? &tmB2 : &tmB1;
        const int k0 = (seg2 ? kb - p.kb1 : kb) * BK;
        mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
        // K-major operand: one box of [rows][64 k].  MN-major operand (stored [k][mn]): boxes of
        // 64 (mn) x 64 (k), one 8 KB slab per 64 mn.
        if constexpr (!A_MN) {
          tma_load_2d(sa, ta, &full_bar[stage], k0, m_blk * BM);
        } else {
#pragma unroll
          for (int h = 0; h < BM / 64; ++h)
            tma_load_2d(sa + h * 8192, ta, &full_bar[stage], m_blk * BM + h * 64, k0);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sb, tb, &full_bar[stage], k0, n_blk * BN);
        } else {
#pragma unroll
          for (int h = 0; h < BN / 64; ++h)
            tma_load_2d(sb + h * 8192, tb, &full_bar[stage], n_blk * BN + h * 64, k0);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && elect_one_sync()) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int split = tile / (p.num_m_blocks * p.num_n_blocks);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(kb_begin + p.kb_per_split, kb_total);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * C::ACC_STRIDE;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart; +32 B per K=16 step.
          // MN-major: 64(mn) x 8(k) atoms of 1024 B; next 64 mn at +8192 (LBO), next 8 k at +1024
          // (SBO); K=16 step = two k-groups = +2048 B.
          const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024)
                                   : make_smem_desc(sa + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024)
                                   : make_smem_desc(sb + k * 32, 16, 1024);
          umma_bf16(tmem_d, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int mn = p.num_m_blocks * p.num_n_blocks;
      const int split = tile / mn;
      int m_blk, n_blk;
      tile_coords(tile - split * mn, p.num_m_blocks, p.num_n_blocks, p.gm, m_blk, n_blk);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr0 = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr0 + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok) epilogue_store32(p, r, row, col0, split);
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
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const GemmArgs& a, cudaStream_t stream) {
  using C = Cfg<BN>;
  GemmParams p;
  p.M = a.M;
  p.N = a.N;
  p.kb1 = (a.K1 + BK - 1) / BK;
  p.kb2 = (a.K2 + BK - 1) / BK;
  p.num_m_blocks = (a.M + BM - 1) / BM;
  p.num_n_blocks = (a.N + BN - 1) / BN;
  const int kb_total = p.kb1 + p.kb2;
  p.splits = a.splits < 1 ? 1 : (a.splits > kb_total ? kb_total : a.splits);
  p.kb_per_split = (kb_total + p.splits - 1) / p.splits;
  p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
  p.C = a.C;
  p.ldc = a.ldc;
  p.c_split_stride = a.c_split_stride;
  p.c_fp32 = a.c_fp32;
  p.bias = reinterpret_cast<const bf16*>(a.bias);
  p.residual = reinterpret_cast<const bf16*>(a.residual);
  p.ldr = a.ldr;
  p.alpha = a.alpha;
  p.gm = raster_group((long long)a.M * (a.K1 + a.K2) * 2, p.num_m_blocks);

  CUtensorMap tA1, tB1, tA2, tB2;
  int rc;
  // K-major: X[rows][K] -> dims {K, rows}, box {64, tile rows}.  MN-major: X[K][mn] -> dims {mn, K}, box {64, 64}.
  if ((rc = A_MN ? make_map(&tA1, a.A1, a.M, a.K1, a.lda1, 64, BK) : make_map(&tA1, a.A1, a.K1, a.M, a.lda1, BK, BM))) return rc;
  if ((rc = B_MN ? make_map(&tB1, a.B1, a.N, a.K1, a.ldb1, 64, BK) : make_map(&tB1, a.B1, a.K1, a.N, a.ldb1, BK, BN))) return rc;
  if (a.K2 > 0) {
    if ((rc = A_MN ? make_map(&tA2, a.A2, a.M, a.K2, a.lda2, 64, BK) : make_map(&tA2, a.A2, a.K2, a.M, a.lda2, BK, BM))) return rc;
    if ((rc = B_MN ? make_map(&tB2, a.B2, a.N, a.K2, a.ldb2, 64, BK) : make_map(&tB2, a.B2, a.K2, a.N, a.ldb2, BK, BN))) return rc;
  } else {
    tA2 = tA1;
    tB2 = tB1;
  }
  auto kern = gemm_kernel<BN, A_MN, B_MN>;
  static DeviceOnce attr_once;  // per template instantiation, per device
  if (attr_once.first())
    B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  const int tiles = p.num_m_blocks * p.num_n_blocks * p.splits;
  int ctas = num_sms();
  if (a.max_ctas > 0 && a.max_ctas < ctas) ctas = a.max_ctas;
  if (tiles < ctas) ctas = tiles;
  B200RL_CUDA_OK(launch_pdl(kern, dim3(ctas), dim3(256), C::SMEM_BYTES, stream, tA1, tB1, tA2, tB2, p));
  B200RL_LAUNCH_OK();
  return 0;
}

// relative per-flop cost of the narrower pair tiles (measured on B200, scripts/bench_gemm.py)
static double g_pair_cost224 = 1.12, g_pair_cost192 = 1.35;

int gemm_dispatch(const GemmArgs& a, cudaStream_t stream) {
  B200RL_REQUIRE(a.M > 0 && a.N > 0 && a.K1 > 0 && a.K2 >= 0, "gemm: bad shape M=%d N=%d K1=%d K2=%d",
                 a.M, a.N, a.K1, a.K2);
  B200RL_REQUIRE(a.A1 && (a.B1 || a.nf4_packed) && a.C, "gemm: null operand");
  if (a.nf4_packed)
    B200RL_REQUIRE(!(a.mn_major & 1) && a.splits <= 1 && gemm_nf4_supported(a.M, a.N, a.K1) &&
                       (a.force_bn == 0 || a.force_bn == 256 || a.force_bn == 512),
                   "gemm(nf4): in-kernel dequant needs the CTA-pair kernel (M > 128, N >= 256, N %% 64 == 0, K1 %% 64 == 0)");
  B200RL_REQUIRE(a.K2 == 0 || (a.A2 && a.B2), "gemm: K2 > 0 needs A2/B2");
  B200RL_REQUIRE(a.N % 8 == 0 && a.ldc % 8 == 0, "gemm: N and ldc must be multiples of 8 (N=%d ldc=%lld)",
                 a.N, a.ldc);
  B200RL_REQUIRE((reinterpret_cast<uintptr_t>(a.C) & 15u) == 0, "gemm: C must be 16-byte aligned");
  B200RL_REQUIRE(a.splits <= 1 || a.c_fp32, "gemm: split-K needs fp32 output slabs");
  if (a.residual)
    B200RL_REQUIRE(a.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15u) == 0,
                   "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  if (a.bias)
    B200RL_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15u) == 0, "gemm: bias must be 16-byte aligned");
  int bn = a.force_bn;
  const bool a_mn = (a.mn_major & 1) != 0, b_mn = (a.mn_major & 2) != 0;
  if (a.ext_B)
    B200RL_REQUIRE(!a_mn && a.splits <= 1 && a.A2 && a.B2 && gemm_ext_supported(a.M, a.N, a.K2) &&
                       (a.force_bn == 0 || a.force_bn == 256 || a.force_bn == 512),
                   "gemm(ext): the in-kernel LoRA intermediate needs the CTA-pair kernel (M > 128, N >= 256, K2 in {64, 128})");
  if (a_mn) {
    B200RL_REQUIRE(b_mn, "gemm: A MN-major with B K-major is not instantiated");
    B200RL_REQUIRE(a.M % 8 == 0, "gemm(dW form): M must be a multiple of 8");
    if (bn == 0) bn = a.N <= 64 ? 64 : 128;
    if (bn == 64) return launch<64, true, true>(a, stream);
    if (bn == 128) return launch<128, true, true>(a, stream);
    return set_error(B200RL_ERR_UNSUPPORTED, "gemm(dW form): BN=%d not instantiated", bn);
  }
  if (a.fuse) {
    B200RL_REQUIRE(!a_mn && a.splits <= 1 && gemm_fuse_supported(a.M, a.fuse == 1 ? a.N / 2 : a.N),
                   "gemm: fused SwiGLU epilogue needs the CTA-pair kernel (M > 128, I %% 128 == 0)");
    return gemm_pair_dispatch(a, 256, stream);
  }
  // CTA-pair kernel (cta_group::2, 256-row tiles) for the large activation x weight GEMMs
  if (gemm_pair_enabled() && a.splits <= 1 && a.M > BM && a.N >= 256 &&
      (a.force_bn == 0 || a.force_bn == 128 || a.force_bn == 192 || a.force_bn == 224 || a.force_bn == 256 ||
       a.force_bn == 512)) {
    // Pair tiles are 256 x BN.  256 x 256 is the most efficient per flop (the 128-wide pair tile is smem-bound
    // again, measured 630-750 TFLOP/s), but N = 3584 (14 tiles of 256) leaves the last of 3.4 waves 40 % full on
    // 74 CTA pairs: pick the width in {256, 224, 192} with the lowest waves x per-tile cost.
    int pbn = a.force_bn;
    // wide 256 x 512 tiles (25 % fewer operand bytes per flop) when every CTA pair gets at least one of them
    if (pbn == 0 && gemm_pair_wide_for(a.M, a.N, a.K1 + a.K2)) pbn = 512;
    if (pbn == 0 && (a.ext_B || a.nf4_packed)) pbn = 256;   // ext units / NF4 producers: 256-column (sub-)tiles only
    if (pbn == 0) {
      const int clusters = num_sms() / 2;
      const long long mb = (a.M + 2 * BM - 1) / (2 * BM);
      double best = 1e30;
      const int cands[3] = {256, 224, 192};
      const double cost[3] = {256.0, 224.0 * g_pair_cost224, 192.0 * g_pair_cost192};
      for (int i = 0; i < 3; ++i) {
        const long long tiles = mb * ((a.N + cands[i] - 1) / cands[i]);
        // waves, counting the K-split of a last wave that is at most half full (gemm2_tcgen05.cu)
        double waves = (double)(tiles / clusters);
        if (tiles % clusters) waves += 1.0 / pair_tail_split(tiles, clusters, (a.K1 + BK - 1) / BK + (a.K2 + BK - 1) / BK);
        const double t = waves * cost[i];
        if (t < best - 1e-9) {
          best = t;
          pbn = cands[i];
        }
      }
    }
    return gemm_pair_dispatch(a, pbn, stream);
  }
  if (bn == 0) {
    if (a.N <= 64) bn = 64;
    else if (a.N <= 128) bn = 128;
    else {
      // pick the tile width with the best wave efficiency on this GPU
      const int sms = num_sms();
      const int mb = (a.M + BM - 1) / BM;
      double best = -1;
      const int cands[3] = {256, 192, 128};
      for (int i = 0; i < 3; ++i) {
        const int c = cands[i];
        const long long tiles = (long long)mb * ((a.N + c - 1) / c);
        const long long waves = (tiles + sms - 1) / sms;
        // useful work / (waves * full-tile work); wider tiles get a small bonus for smem traffic
        double eff = ((double)a.M * a.N) / ((double)waves * sms * BM * c);
        eff *= (c == 256 ? 1.0 : (c == 192 ? 0.97 : 0.93));
        if (eff > best) {
          best = eff;
          bn = c;
        }
      }
    }
  }
  if (b_mn) {
    switch (bn) {
      case 64: return launch<64, false, true>(a, stream);
      case 128: return launch<128, false, true>(a, stream);
      case 192: return launch<192, false, true>(a, stream);
      case 256: return launch<256, false, true>(a, stream);
      default: return set_error(B200RL_ERR_UNSUPPORTED, "gemm: BN=%d not instantiated", bn);
    }
  }
  switch (bn) {
    case 64: return launch<64, false, false>(a, stream);
    case 128: return launch<128, false, false>(a, stream);
    case 192: return launch<192, false, false>(a, stream);
    case 256: return launch<256, false, false>(a, stream);
    default: return set_error(B200RL_ERR_UNSUPPORTED, "gemm: BN=%d not instantiated", bn);
  }
}

}  // namespace b200rl

using namespace b200rl;

// C ABI — declared in include/b200rl.h
extern "C" int b200rl_gemm(const void* A1, long long lda1, const void* B1, long long ldb1, int K1,
                           const void* A2, long long lda2, const void* B2, long long ldb2, int K2,
                           void* C, long long ldc, int c_fp32, const void* bias,
                           const void* residual, long long ldr, float alpha, int M, int N,
                           int mn_major, int splits, long long c_split_stride, int force_bn,
                           int max_ctas, void* stream) {
  GemmArgs a;
  a.A1 = A1; a.B1 = B1; a.A2 = A2; a.B2 = B2;
  a.lda1 = lda1; a.ldb1 = ldb1; a.lda2 = lda2; a.ldb2 = ldb2;
  a.K1 = K1; a.K2 = K2;
  a.C = C; a.ldc = ldc; a.c_fp32 = c_fp32;
  a.bias = bias; a.residual = residual; a.ldr = ldr;
  a.alpha = alpha; a.M = M; a.N = N;
  a.mn_major = mn_major; a.splits = splits; a.c_split_stride = c_split_stride;
  a.force_bn = force_bn; a.max_ctas = max_ctas;
  return gemm_dispatch(a, reinterpret_cast<cudaStream_t>(stream));
}

// Base + LoRA projection with the LoRA intermediate produced by the same launch (CTA-pair kernel, "ext units"):
//   U[M,K2] = ext_alpha * A1 . Bext^T   (bf16, written to `U`)     then     C = A1.B1^T + U.B2^T (+bias) (+residual)
// mn_major 0: B1 [N,K1], Bext [K2,K1], B2 [N,K2] (forward y = x W^T + s (x A^T) B^T, distributed_actor.py:241-243 through
// PEFT's LoRA formula); mn_major 2 (dX form): B1 [K1,N], Bext [K1,K2], B2 [K2,N] (dx = dy W + s (dy B) A).
extern "C" int b200rl_gemm_lora(const void* A1, long long lda1, const void* B1, long long ldb1, int K1,
                                const void* Bext, long long ld_ext, float ext_alpha, void* U, long long ldu,
                                const void* B2, long long ldb2, int K2, void* C, long long ldc, const void* bias,
                                const void* residual, long long ldr, int M, int N, int mn_major, int force_bn, void* stream) {
  B200RL_REQUIRE(mn_major == 0 || mn_major == 2, "gemm_lora: mn_major must be 0 (forward) or 2 (dX form)");
  B200RL_REQUIRE(Bext && U, "gemm_lora: null LoRA operand");
  GemmArgs a;
  a.A1 = A1; a.B1 = B1; a.A2 = U; a.B2 = B2;
  a.lda1 = lda1; a.ldb1 = ldb1; a.lda2 = ldu; a.ldb2 = ldb2;
  a.K1 = K1; a.K2 = K2;
  a.C = C; a.ldc = ldc; a.c_fp32 = 0;
  a.bias = bias; a.residual = residual; a.ldr = ldr;
  a.alpha = 1.f; a.M = M; a.N = N;
  a.mn_major = mn_major; a.splits = 1; a.c_split_stride = 0;
  a.force_bn = force_bn; a.max_ctas = 0;
  a.ext_B = Bext; a.ld_ext_b = ld_ext; a.ext_alpha = ext_alpha;
  return gemm_dispatch(a, reinterpret_cast<cudaStream_t>(stream));
}

// C = A1 . dequant(NF4)^T (+ A2.B2^T) (+bias) (+residual) with the 4-bit base weight expanded inside the GEMM mainloop
// (reference: load_in_4bit weights, distributed_actor.py:16-17, :58-66).  mn_major 0: codes of W [N, K1]; 2: of W [K1, N].
extern "C" int b200rl_gemm_nf4(const void* A1, long long lda1, const void* packed, const float* absmax, int K1,
                               const void* A2, long long lda2, const void* B2, long long ldb2, int K2, void* C,
                               long long ldc, const void* bias, const void* residual, long long ldr, int M, int N,
                               int mn_major, int force_bn, void* stream) {
  B200RL_REQUIRE(mn_major == 0 || mn_major == 2, "gemm_nf4: mn_major must be 0 or 2");
  GemmArgs a;
  a.A1 = A1; a.B1 = nullptr; a.A2 = A2; a.B2 = B2;
  a.lda1 = lda1; a.ldb1 = 0; a.lda2 = lda2; a.ldb2 = ldb2;
  a.K1 = K1; a.K2 = K2;
  a.C = C; a.ldc = ldc; a.c_fp32 = 0;
  a.bias = bias; a.residual = residual; a.ldr = ldr;
  a.alpha = 1.f; a.M = M; a.N = N;
  a.mn_major = mn_major; a.splits = 1; a.c_split_stride = 0;
  a.force_bn = force_bn; a.max_ctas = 0;
  a.nf4_packed = packed; a.nf4_absmax = absmax;
  return gemm_dispatch(a, reinterpret_cast<cudaStream_t>(stream));
}

// Fused SwiGLU GEMMs (CTA-pair kernel).  mode 1: gu[M,2I] = A1.B1^T + A2.B2^T and act[M,I] = silu(gate)*up;
// mode 2: dgu[M,2I] = swiglu_bwd(gu, dact = A1.B1 + A2.B2) with B stored [K, I].  Bit-identical to b200rl_gemm
// followed by b200rl_swiglu_fwd / b200rl_swiglu_bwd.
extern "C" int b200rl_gemm_swiglu(int mode, const void* A1, long long lda1, const void* B1, long long ldb1, int K1,
                                  const void* A2, long long lda2, const void* B2, long long ldb2, int K2,
                                  void* C, long long ldc, void* aux, long long ld_aux, int M, int I, void* stream) {
  B200RL_REQUIRE(mode == 1 || mode == 2, "gemm_swiglu: mode must be 1 (forward) or 2 (backward)");
  GemmArgs a;
  a.A1 = A1; a.B1 = B1; a.A2 = A2; a.B2 = B2;
  a.lda1 = lda1; a.ldb1 = ldb1; a.lda2 = lda2; a.ldb2 = ldb2;
  a.K1 = K1; a.K2 = K2;
  a.C = C; a.ldc = ldc; a.c_fp32 = 0;
  a.bias = nullptr; a.residual = nullptr; a.ldr = 0;
  a.alpha = 1.f; a.M = M; a.N = mode == 1 ? 2 * I : I;
  a.mn_major = mode == 1 ? 0 : 2; a.splits = 1; a.c_split_stride = 0;
  a.force_bn = 0; a.max_ctas = 0;
  a.fuse = mode; a.aux = aux; a.ld_aux = ld_aux;
  return gemm_dispatch(a, reinterpret_cast<cudaStream_t>(stream));
}
