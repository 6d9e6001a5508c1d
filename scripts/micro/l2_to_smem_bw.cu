// Micro-benchmark: aggregate L2 -> shared-memory bandwidth of bulk async copies (the path TMA operand loads take), all SMs
// streaming a buffer that fits in L2 (48 MB, second pass onwards) vs one that does not (4 GB, DRAM-bound reference).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I distrl_llm_b200/csrc -I include scripts/micro/l2_to_smem_bw.cu -o /tmp/l2bw
#include "common.cuh"
#include <cstdio>
using namespace b200rl;

constexpr int CHUNK = 32768;   // bytes per bulk copy (one 256-row x 64-col bf16 operand stage)
constexpr int STAGES = 6;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t bytes, int passes, unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[STAGES];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t n_chunks = bytes / CHUNK;
    // CTA b streams chunks b, b + grid, ... (neighbouring CTAs read neighbouring chunks, like tiles sharing a panel do not)
    size_t issued = 0, done = 0;
    const size_t total = (n_chunks / gridDim.x) * passes;
    auto issue = [&](size_t i) {
      const size_t c = (i % (n_chunks / gridDim.x)) * gridDim.x + blockIdx.x;
      const int s = i % STAGES;
      mbar_arrive_expect_tx(&full[s], CHUNK);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * CHUNK),
                   "l"(src + c * (size_t)CHUNK), "r"(CHUNK), "r"(smem_u32(&full[s]))
                   : "memory");
    };
    for (; issued < STAGES && issued < total; ++issued) issue(issued);
    for (; done < total; ++done) {
      mbar_wait(&full[done % STAGES], (done / STAGES) & 1);
      if (issued < total) { issue(issued); ++issued; }
    }
    if (blockIdx.x == 0) sink[0] = done;
  }
}

int main() {
  uint8_t* buf;
  const size_t big = 4ull << 30;
  cudaMalloc(&buf, big);
  cudaMemset(buf, 1, big);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  const size_t smem = STAGES * CHUNK + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (size_t bytes : {(size_t)24 << 20, (size_t)48 << 20, (size_t)96 << 20, big}) {
    const int passes = bytes == big ? 2 : (int)((8ull << 30) / bytes);
    k<<<148, 128, smem>>>(buf, bytes, 2, sink);   // warm L2
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148, 128, smem>>>(buf, bytes, passes, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double moved = (double)(bytes / CHUNK / 148) * 148 * CHUNK * passes;
    printf("buffer %6zu MB, %3d passes: %.2f TB/s into shared memory (148 CTAs, %d x %d KB bulk copies in flight each)\n", bytes >> 20, passes,
           moved / ms / 1e9, STAGES, CHUNK / 1024);
  }
  return 0;
}
