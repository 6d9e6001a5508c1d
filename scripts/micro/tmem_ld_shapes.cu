// Micro-benchmark: tcgen05.ld throughput by shape (all 4 KB per warp instruction), 4 and 8 warps, one CTA per SM.
#include "common.cuh"
#include <cstdio>
using namespace b200rl;

#define REGS32(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), \
  "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), \
  "=r"(v[30]), "=r"(v[31])
#define LIST32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}"

template <int SHAPE>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v) {
  if (SHAPE == 0) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 1) asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 2) asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
  if (SHAPE == 3) asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " LIST32 ", [%32];" : REGS32(v) : "r"(taddr) : "memory");
}

template <int SHAPE, int BATCH>
__global__ void __launch_bounds__(256, 1) k(int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_smem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; r += BATCH) {
    uint32_t w[BATCH][32];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) ld<SHAPE>(tb + (((r + b) & 3) * 64 + (warp >> 2) * 256), w[b]);
    tmem_ld_wait();
#pragma unroll
    for (int b = 0; b < BATCH; ++b) acc += __uint_as_float(w[b][(r + b) & 31]);
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base_smem, 512); }
}

template <int SHAPE, int BATCH>
void run(const char* name, int threads, long long* out, float* sink) {
  const int reps = 2048;
  long long h;
  for (int it = 0; it < 2; ++it) {
    k<SHAPE, BATCH><<<148, threads>>>(reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-22s batch %d, %d warps: %6.1f cycles per 4 KB warp load, %5.0f B/clk/SM\n", name, BATCH, threads / 32, (double)h / reps,
         (double)reps * (threads / 32) * 4096.0 / h);
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8); cudaMalloc(&sink, 4);
  for (int threads : {32, 128, 256}) {
    run<0, 1>("32x32b.x32", threads, out, sink);
    run<0, 2>("32x32b.x32", threads, out, sink);
    run<1, 1>("16x256b.x8", threads, out, sink);
    run<1, 2>("16x256b.x8", threads, out, sink);
    run<2, 1>("16x128b.x16", threads, out, sink);
    run<3, 1>("16x64b.x32", threads, out, sink);
  }
  return 0;
}
