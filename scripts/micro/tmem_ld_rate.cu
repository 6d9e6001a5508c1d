// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (32x32b.x32, 4 or 8 warps, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I distrl_llm_b200/csrc -I include scripts/micro/tmem_ld_rate.cu -o /tmp/tmem_ld_rate
#include "common.cuh"
#include <cstdio>
using namespace b200rl;

template <int MODE>   // 0: ld only, 1: st only, 2: ld + 32 MUFU.EX2 per load (softmax-like)
__global__ void __launch_bounds__(256, 1) tmem_kernel(int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_smem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 32 + i;
  tmem_st_32x32(tb, v);
  tmem_st_32x32(tb + 32, v);
  tmem_st_wait();
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const uint32_t col = ((r & 7) * 32 + (warp >> 2) * 256) & 511;
    if (MODE == 1) {
      tmem_st_32x32(tb + col, v);
      tmem_st_wait();
    } else {
      uint32_t w[32];
      tmem_ld_32x32(tb + col, w);
      tmem_ld_wait();
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += ex2_approx(__uint_as_float(w[i]) * 1e-30f);
      } else {
        acc += __uint_as_float(w[r & 31]);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base_smem, 512); }
}

template <int MODE>
void run(const char* name, int threads, long long* out, float* sink) {
  const int reps = 2048;
  long long h;
  for (int it = 0; it < 2; ++it) {
    tmem_kernel<MODE><<<148, threads>>>(reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)reps * (threads / 32) * 4096.0;
  printf("%-34s %d warps: %.1f cycles per 4 KB warp access, %.0f B/clk/SM\n", name, threads / 32, (double)h / reps, bytes / h);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  run<0>("tcgen05.ld 32x32b.x32", 128, out, sink);
  run<0>("tcgen05.ld 32x32b.x32", 256, out, sink);
  run<1>("tcgen05.st 32x32b.x32", 128, out, sink);
  run<1>("tcgen05.st 32x32b.x32", 256, out, sink);
  run<2>("tcgen05.ld + 32 ex2 per thread", 128, out, sink);
  run<2>("tcgen05.ld + 32 ex2 per thread", 256, out, sink);
  return 0;
}
