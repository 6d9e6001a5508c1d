// Numerics check of tcgen05.mma with the A operand in TENSOR MEMORY (TS form), the layout assumption the attention kernels
// use for P / dS: A[m][k] (bf16) of row m lives in TMEM lane m, two consecutive k per 32-bit column (low half = even k), a
// K = 16 MMA step reads 8 columns.  A is written with tcgen05.st (32x32b), B is a K-major SW128 tile in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I distrl_llm_b200/csrc -I include scripts/micro/umma_ts_numerics.cu -o /tmp/umma_ts
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace b200rl;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t id, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(id), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// a: [128][64] bf16 row-major, b: [N][64] bf16 row-major (K-major B), b_mn variant: bt [64][N] (MN-major B), d: [128][N] f32
template <int N, bool B_MN>
__global__ void __launch_bounds__(128, 1) ts_kernel(const bf16* a, const bf16* b, float* d) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int t = threadIdx.x, warp = t >> 5;
  if (!B_MN) {
    // K-major SW128 tile: row n at n*128 B, 16-byte chunk c at c ^ (n & 7)
    for (int i = t; i < N * 8; i += 128) {
      const int n = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(sm + n * 128 + ((c ^ (n & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + n * 64 + c * 8);
    }
  } else {
    // MN-major SW128: slabs of 64 n-columns (8 KB each: 64 k-rows x 128 B), row k at k*128 B, chunk c (8 n values) at c ^ (k & 7)
    for (int i = t; i < 64 * (N / 8); i += 128) {
      const int k = i / (N / 8), cn = i % (N / 8);
      const int slab = cn >> 3, c = cn & 7;
      *reinterpret_cast<uint4*>(sm + slab * 8192 + k * 128 + ((c ^ (k & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + k * N + cn * 8);
    }
  }
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_smem;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  // A -> TMEM columns [256, 288): thread = row
  {
    uint32_t r[32];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a + t * 64);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = src[j];
    tmem_st_32x32(tb + 256 + lane_addr, r);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (t == 0) {
    constexpr uint32_t id = idesc(N, B_MN);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t db = B_MN ? desc_sw128(base + kk * 2048, 8192, 1024) : desc_sw128(base + kk * 32, 16, 1024);
      umma_ts(tb, tb + 256 + kk * 8, db, id, kk > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(tb + lane_addr + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[t * N + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, bool B_MN>
int run() {
  std::vector<bf16> ha(128 * 64), hb(N * 64);
  std::vector<float> fa(128 * 64), fb(N * 64);
  srand(1234 + N + B_MN);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fb[i] = __bfloat162float(hb[i]); }
  // hb is [N][64] (K-major) or, for B_MN, interpreted as [64][N]
  bf16 *da, *db; float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const size_t smem = 64 * N * 2 + 2048;
  cudaFuncSetAttribute(ts_kernel<N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ts_kernel<N, B_MN><<<1, 128, smem>>>(da, db, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d b_mn=%d: %s\n", N, (int)B_MN, cudaGetErrorString(e)); return 1; }
  std::vector<float> hd(128 * N);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)fa[m * 64 + k] * (B_MN ? fb[k * N + n] : fb[n * 64 + k]);
      maxerr = fmax(maxerr, fabs(ref - hd[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("TS-mode A in TMEM, N=%3d, B %s: max |err| %.3e (max |ref| %.2f) %s\n", N, B_MN ? "MN-major" : "K-major ", maxerr, maxref,
         maxerr < 1e-3 ? "OK" : "MISMATCH");
  return maxerr < 1e-3 ? 0 : 2;
}

int main() {
  int rc = 0;
  rc |= run<64, false>();
  rc |= run<128, false>();
  rc |= run<64, true>();
  rc |= run<128, true>();
  return rc;
}
