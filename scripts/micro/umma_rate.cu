// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) by N, B-operand majorness and A source
// (shared-memory descriptor vs TMEM).  One CTA per SM, operands are whatever bytes sit in shared memory / TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I distrl_llm_b200/csrc -I include scripts/micro/umma_rate.cu -o gpurun_out/umma_rate
#include "common.cuh"
#include <cstdio>
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

// MODE 0: SS, B K-major; 1: SS, B MN-major; 2: TS (A in TMEM), B K-major; 3: TS, B MN-major
// NACC: number of distinct accumulators the chain alternates over; k-steps walk 4 x 32-byte offsets (like a real tile).
// Descriptors are built once; the unrolled loop only adds compile-time offsets, so the issuing thread is not the limit.
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u + (i * 2654435761u & 0x03ff03ffu);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_smem;
  const uint32_t sA = base, sB = base + 32768;
  if (threadIdx.x == 0) {
    constexpr uint32_t id = idesc(N, MODE & 1);
    const uint64_t da0 = desc_sw128(sA, 16, 1024);
    const uint64_t db0 = (MODE & 1) ? desc_sw128(sB, 8192, 1024) : desc_sw128(sB, 16, 1024);
    constexpr uint32_t bstep = (MODE & 1) ? 2048 / 16 : 32 / 16;
    long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int kk = u & 3;
        const uint32_t d = tb + (u % NACC) * N;
        const uint32_t acc = (r > 0 || u >= NACC) ? 1u : 0u;
        if (MODE >= 2) umma_ts(d, tb + 384 + kk * 8, db0 + kk * bstep, id, acc);
        else umma_bf16(d, da0 + kk * 2, db0 + kk * bstep, id, acc);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N, int MODE, int NACC>
int run(long long* out) {
  const char* names[4] = {"SS  B K-major ", "SS  B MN-major", "TS  B K-major ", "TS  B MN-major"};
  const int reps = 512;
  const size_t smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2];
  for (int it = 0; it < 2; ++it) {
    rate_kernel<N, MODE, NACC><<<148, 128, smem>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s N=%d: %s\n", names[MODE], N, cudaGetErrorString(e)); return 1; }
  }
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("%s N=%3d accumulators=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (tensor floor %d, smem-read floor %d)\n", names[MODE], N, NACC,
         (double)h[0] / reps, (double)h[1] / reps, N / 2, ((MODE >= 2 ? 0 : 4096) + N * 32) / 128);
  return 0;
}

template <int MODE>
int run_mode(long long* out) {
  return run<64, MODE, 1>(out) || run<64, MODE, 2>(out) || run<128, MODE, 1>(out) || run<128, MODE, 2>(out) || run<256, MODE, 1>(out);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  return run_mode<0>(out) || run_mode<1>(out) || run_mode<2>(out) || run_mode<3>(out);
}
