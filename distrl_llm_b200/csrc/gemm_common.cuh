// Shared pieces of the tcgen05 GEMM kernels (single-CTA and CTA-pair variants): tile constants,
// kernel parameters, UMMA descriptors, tensor-map construction, host argument block.
#pragma once
#include "common.cuh"
#include <stdlib.h>

namespace b200rl {

static constexpr int BM = 128;
static constexpr int BK = 64;           // 64 bf16 = 128 B = one 128B-swizzle row
static constexpr int UMMA_K = 16;
static constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB

struct GemmParams {
  int M, N;
  int kb1, kb2;          // k-blocks (of 64) in segment 1 / 2
  int num_m_blocks, num_n_blocks, splits, kb_per_split;
  void* C;
  long long ldc;
  long long c_split_stride;  // elements between split-K slabs
  int c_fp32;
  const bf16* bias;
  const bf16* residual;
  long long ldr;
  float alpha;
  // CTA-pair kernel only: split-K of the LAST (partial) wave.  Work units 0..tail_first-1 are whole tiles; every
  // later tile is cut into tail_split K-ranges that run concurrently on different CTA pairs.  Parts 1.. spill their
  // fp32 accumulators to tail_ws and raise tail_flags (value = tail_epoch); part 0 adds them and runs the epilogue.
  int tail_first = 0x7fffffff;
  int tail_split = 1;
  float* tail_ws = nullptr;
  int* tail_flags = nullptr;
  int tail_epoch = 0;
  int gm = 8;  // m-blocks per rasterisation group (tile_coords); host: raster_group()
  // CTA-pair kernel only: fused SwiGLU epilogues (GemmArgs::fuse)
  const bf16* aux_in = nullptr;
  bf16* aux_out = nullptr;
  long long ld_aux = 0;
  int fuse_I = 0;
  // CTA-pair kernel only: LoRA intermediate computed INSIDE the launch (GemmArgs::ext_B).  The first n_ext work units
  // ("ext units", one per 256-row m-block) compute U[m rows, ext_n] = ext_alpha * A1[m rows] . Bext^T over the whole K1
  // into ext_out and raise ext_flags (value = ext_epoch); every ordinary tile waits for the flags of its m-block right
  // before it stages the K-extension segment (A2 == ext_out).
  int n_ext = 0;
  int ext_n = 0;               // 64 or 128 columns (= the group's K2)
  bf16* ext_out = nullptr;
  long long ld_ext = 0;
  float ext_alpha = 1.f;
  int* ext_flags = nullptr;    // [num_m_blocks][2 CTAs][4 warps]
  int ext_epoch = 0;
  // CTA-pair kernel, NF4 instantiations: the segment-1 B operand is NF4 storage (GemmArgs::nf4_packed), expanded to bf16
  // by four producer warps per CTA straight into the swizzled shared-memory stage the UMMA reads
  const uint8_t* nf4_q = nullptr;   // packed codes, 2 per byte, row-major like the dense matrix
  const float* nf4_am = nullptr;    // one fp32 absmax per 64 consecutive values
  int K1 = 0;                       // segment-1 depth in elements (row length of a K-major B / row count of an MN-major B)
};

// ---- descriptors ---------------------------------------------------------------------------
// Shared-memory matrix descriptor (tcgen05), 128B swizzle. Field layout checked against
// cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start[0,14) lbo[16,30) sbo[32,46) version[46,48)
// layout_type[61,64) (SWIZZLE_128B = 2); addresses/offsets in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// The same descriptor in two halves, for the MMA-issuing thread: every instruction it spends per MMA is serial latency
// (one thread issues all MMAs of a CTA pair; profiles/r2_run11, r2_run12: ~100 cycles per MMA with the descriptor rebuilt
// each time vs 128 cycles of execution for a 256 x 256 x 16 pair MMA).  The low word is built once per k-block and a byte
// offset is ADDED per MMA (the address field holds addr >> 4 in bits 0-13; shared memory is < 256 KB: no carry leaves it).
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t SMEM_DESC_HI_SBO1024 = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t smem_desc_at(uint32_t lo, uint32_t byte_off) {
  return ((uint64_t)SMEM_DESC_HI_SBO1024 << 32) | (uint64_t)(lo + (byte_off >> 4));
}
// Instruction descriptor (InstrDescriptor): c_format[4,6)=F32(1), a_format[7,10)=BF16(1),
// b_format[10,13)=BF16(1), a_major bit15, b_major bit16 (0 = K-major, 1 = MN-major),
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}


// Tile rasterisation: groups of GM m-blocks sweep all n-blocks before the next group, so the ~74-148 tiles in
// flight form a compact (GM x ~9..18) patch of C and share both A and B tiles through L2 (ncu on the m-fastest
// order: 3.2-4.4x DRAM read amplification on the MLP GEMMs, A = 261 MB does not fit the 126 MB L2).
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int GM, int& m_blk, int& n_blk) {
  const int per_group = GM * num_n;
  const int group = tile / per_group;
  const int first_m = group * GM;
  const int gm = min(num_m - first_m, GM);
  const int in_group = tile - group * per_group;
  m_blk = first_m + in_group % gm;
  n_blk = in_group / gm;
}

// epilogue shared by both kernels: 32 accumulator columns of one row -> alpha / bias / residual -> global
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, const uint32_t* r, int row, int col0,
                                                 int split) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int col = col0 + g * 8;
    if (col < p.N) {  // N is a multiple of 8 (checked on the host)
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * p.alpha;
      if (p.bias) {
        float b[8];
        unpack8(*reinterpret_cast<const bf16x8*>(p.bias + col), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += b[i];
      }
      if (p.residual) {
        float b[8];
        unpack8(*reinterpret_cast<const bf16x8*>(p.residual + (long long)row * p.ldr + col), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += b[i];
      }
      if (p.c_fp32) {
        float* dst = reinterpret_cast<float*>(p.C) + (long long)split * p.c_split_stride +
                     (long long)row * p.ldc + col;
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        bf16* dst = reinterpret_cast<bf16*>(p.C) + (long long)split * p.c_split_stride +
                    (long long)row * p.ldc + col;
        *reinterpret_cast<bf16x8*>(dst) = pack8(v);
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map over X[rows][inner] with leading dimension ld (elements), 128B swizzle.
inline int make_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, uint64_t ld,
                    uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * 2) & 15u) != 0)
    return set_error(B200RL_ERR_ARG, "gemm operand must be 16-byte aligned with ld %% 8 == 0 (ld=%llu)",
                     (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(B200RL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu ld=%llu)",
                     (int)r, (unsigned long long)inner, (unsigned long long)rows,
                     (unsigned long long)ld);
  return 0;
}

struct GemmArgs {
  const void *A1, *B1, *A2, *B2;
  long long lda1, ldb1, lda2, ldb2;
  int K1, K2;
  void* C;
  long long ldc;
  int c_fp32;
  const void* bias;
  const void* residual;
  long long ldr;
  float alpha;
  int M, N;
  int mn_major;   // bit0: A stored [K][M] (MN-major), bit1: B stored [K][N]. 0 = TN, 3 = dW form, 2 = dX form
  int splits;     // split-K factor (fp32 output slabs, c_split_stride apart)
  long long c_split_stride;
  int force_bn;   // 0 = heuristic
  int max_ctas;   // 0 = all SMs
  // Fused SwiGLU epilogues (CTA-pair kernel, 256-wide tiles; the caller checks gemm_fuse_supported()):
  //  1: gate|up projection, N = 2I.  Each 256-column tile holds gate columns j0..j0+127 (B rows j0.., staged by
  //     CTA 0) and up columns j0..j0+127 (B rows I+j0.., staged by CTA 1); the epilogue writes gate|up to C [M,2I]
  //     AND silu(gate)*up to aux [M, I] (ld_aux).
  //  2: down-projection input gradient, N = I (dX form).  The epilogue turns the accumulator dact into
  //     dgate|dup using gate|up read from aux [M,2I] (ld_aux) and writes them to C [M,2I].
  int fuse = 0;
  void* aux = nullptr;
  long long ld_aux = 0;
  // LoRA-in-kernel (CTA-pair kernel; the caller checks gemm_ext_supported()): the LoRA intermediate that the K-extension
  // segment consumes is produced by the same launch.  ext_B = Acat [ext_n, K1] K-major (forward: U = s x.Acat^T) or
  // Bcat [K1, ext_n] MN-major (dX form: dU = s dY.Bcat), following mn_major bit 1 like B1.  The result (bf16, scaled by
  // ext_alpha) is written to A2 [M, K2 == ext_n] (lda2), which the K-extension then reads back through TMA.
  const void* ext_B = nullptr;
  long long ld_ext_b = 0;
  float ext_alpha = 1.f;
  // NF4 base weight dequantised IN the mainloop (north_star: "4-bit base weights dequantised on the fly in-kernel";
  // CTA-pair kernel, the caller checks gemm_nf4_supported()): B1 is given as NF4 storage instead of a bf16 matrix —
  // packed codes [N, K1] (mn_major bit 1 clear) or [K1, N] (set), dense row-major, K1 % 64 == 0, N % 64 == 0.
  const void* nf4_packed = nullptr;
  const float* nf4_absmax = nullptr;
};
bool gemm_nf4_supported(int M, int N, int K1);
bool gemm_ext_supported(int M, int N, int K2);
bool gemm_fuse_supported(int M, int I);


// K-split factor of the last partial wave of the CTA-pair kernel (1 = none): only when that wave is at most half full
// and every K-range keeps >= 64 k-blocks -- measured on B200 (profiles/r1_run19*): with K = 3584 the spill / flag /
// reload chain (~15 us) costs more than the quarter wave it saves, with K >= 18944 it saves 7-13 % of the GEMM.
// Rasterisation group height.  The tiles in flight sweep n inside a group of gm m-blocks: B panels are shared by gm
// tiles through L2, A panels are re-read once per n step.  Every group streams all of B once, so fewer, taller groups
// cut the B traffic as long as a group's A panels stay L2-resident (ncu, packed config 2: with gm = 8 the gate|up GEMM
// read B 3x from DRAM at 18 m-blocks and 4.4x at 35).  K-long operands (panels > 5 MB) keep the square-ish 8 x ~9 patch
// that minimises (gm + gn) panels per wave.
extern int g_gemm_raster_forced;   // b200rl_gemm_set_raster (debug / sweeps): > 0 overrides the choice below
inline int raster_group(long long a_bytes, int num_m_blocks) {
  if (g_gemm_raster_forced > 0) return g_gemm_raster_forced;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("B200RL_GEMM_GM");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) return forced;
  if (num_m_blocks <= 8) return num_m_blocks > 0 ? num_m_blocks : 1;
  // m-blocks whose A panels stay L2-resident (40 MB budget: the 126 MB L2 is split over two dies and also carries the
  // B stream and the C write-allocate traffic); balanced groups, never thinner than the square-ish 8
  const long long panel = a_bytes / num_m_blocks;
  const long long fit = panel > 0 ? (40ll << 20) / panel : num_m_blocks;
  // K-long operands (a 256-row A panel above ~5 MB: gate|up dX, down fwd, lm_head dX): nothing stays resident anyway, and
  // n-fastest order (the tiles that share an A panel sit on neighbouring CTA pairs and start together) measured 10-17 %
  // less DRAM traffic than groups of 8 on all three shapes (profiles/r2_run13_raster_sweep.txt)
  if (fit < 8) return 1;
  const long long groups = (num_m_blocks + fit - 1) / fit;
  return (int)((num_m_blocks + groups - 1) / groups);
}

inline int pair_tail_split(long long tiles, int clusters, int kb_total) {
  const long long rem = tiles % clusters;
  if (tiles <= clusters || rem == 0 || rem * 2 > clusters || rem > 128) return 1;
  int S = (int)(clusters / rem);
  if (S > 4) S = 4;
  while (S > 1 && kb_total < 64 * S) --S;
  return S;
}

int gemm_dispatch(const GemmArgs& a, cudaStream_t stream);             // gemm_tcgen05.cu
int gemm_pair_dispatch(const GemmArgs& a, int bn, cudaStream_t stream);  // gemm2_tcgen05.cu
void gemm_pair_set_tail_split(int enable);
bool gemm_pair_enabled();
bool gemm_pair_wide_enabled();   // 256 x 512 "wide" pair tiles (gemm2_tcgen05.cu, PairCfg NT = 2)
bool gemm_pair_wide_for(int M, int n_cols, int k_total);

}  // namespace b200rl
