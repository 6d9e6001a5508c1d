// NF4 (4-bit NormalFloat) base-weight storage: quantise (synthetic / random-init weights) and
// dequantise to bf16 tiles for the tcgen05 GEMMs.
//
// The reference loads `load_in_4bit=True` checkpoints through Unsloth/bitsandbytes
// (reference distributed_actor.py:16-17, :58-66); bitsandbytes itself is NOT in /root/reference
// ([3P], pinned 0.45.2 in requirements.txt:3).  Restated format (QLoRA paper / bitsandbytes
// `quantize_4bit(quant_type="nf4", blocksize=64)`): flat row-major weight, blocks of 64
// consecutive values share one fp32 absmax, each value is the index of the nearest of 16 codebook
// levels of value/absmax, two indices per byte with the EVEN element in the HIGH nibble.
// Double-quantisation of absmax is not restated (absmax stays fp32): "parity unpinned" for
// real bnb checkpoints, self-consistent for the random-init BASELINE configs.
#include "common.cuh"

namespace b200rl {

__constant__ float c_nf4[16] = {-1.0f,
                                -0.6961928009986877f,
                                -0.5250730514526367f,
                                -0.39491748809814453f,
                                -0.28444138169288635f,
                                -0.18477343022823334f,
                                -0.09105003625154495f,
                                0.0f,
                                0.07958029955625534f,
                                0.16093020141124725f,
                                0.24611230194568634f,
                                0.33791524171829224f,
                                0.44070982933044434f,
                                0.5626170039176941f,
                                0.7229568362236023f,
                                1.0f};

__device__ __forceinline__ int nf4_nearest(float x) {
  // midpoints between adjacent levels (decision thresholds)
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 15; ++i) idx += x > 0.5f * (c_nf4[i] + c_nf4[i + 1]);
  return idx;
}

// one warp per 64-value block: lane handles 2 values
__global__ void nf4_quant_kernel(const bf16* __restrict__ w, uint8_t* __restrict__ packed,
                                 float* __restrict__ absmax, long long nblocks) {
  const long long blk = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (blk >= nblocks) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat162 v2 = reinterpret_cast<const __nv_bfloat162*>(w + blk * 64)[lane];
  const float2 v = __bfloat1622float2(v2);
  float am = warp_max(fmaxf(fabsf(v.x), fabsf(v.y)));
  if (lane == 0) absmax[blk] = am;
  const float inv = am > 0.f ? 1.f / am : 0.f;
  const int hi = nf4_nearest(v.x * inv), lo = nf4_nearest(v.y * inv);
  packed[blk * 32 + lane] = (uint8_t)((hi << 4) | lo);
}

// row-major dequant: out[i] = bf16(code * absmax).  Each thread expands 4 packed bytes -> 8 values
// (one 16-byte store); consecutive lanes take consecutive words, so a warp reads 128 B and writes
// 512 B contiguously.  The 16-entry level table is expanded to a 256-entry byte -> (hi, lo) table
// in shared memory: a __constant__ lookup with a per-lane index serialises up to 16-way.
__global__ void __launch_bounds__(256)
nf4_dequant_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                   bf16* __restrict__ out, long long n8) {
  __shared__ float2 lut[256];
  lut[threadIdx.x] = make_float2(c_nf4[threadIdx.x >> 4], c_nf4[threadIdx.x & 15]);
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += stride * U) {
    uint32_t q[U];
    float am[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      q[u] = i < n8 ? __ldg(reinterpret_cast<const uint32_t*>(packed) + i) : 0u;
      am[u] = i < n8 ? __ldg(absmax + (i >> 3)) : 0.f;  // 8 values per word, 64 per block
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n8) {
        bf16x8 o;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float2 v = lut[(q[u] >> (8 * b)) & 0xFFu];
          o.w(b) = pack_bf16x2(v.x * am[u], v.y * am[u]);
        }
        reinterpret_cast<bf16x8*>(out)[i] = o;
      }
    }
  }
}

// transposed dequant: W is [rows, cols] row-major (cols % 64 == 0), out = W^T [cols, rows].
// 64x64 tile through shared memory; both global sides are 128-byte coalesced.
__global__ void nf4_dequant_t_kernel(const uint8_t* __restrict__ packed,
                                     const float* __restrict__ absmax, bf16* __restrict__ out,
                                     int rows, int cols) {
  __shared__ bf16 tile[64][66];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  // 256 threads: thread t loads row (t/8 + 32*p), 8 values at column (t%8)*8
  for (int p = 0; p < 2; ++p) {
    const int r = (threadIdx.x >> 3) + 32 * p;
    const int cv = (threadIdx.x & 7) * 8;
    if (r0 + r < rows) {
      const long long e = (long long)(r0 + r) * cols + c0 + cv;  // flat element index
      const uint32_t q = *reinterpret_cast<const uint32_t*>(packed + (e >> 1));
      const float am = absmax[e >> 6];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t byte = (q >> (8 * b)) & 0xFFu;
        tile[r][cv + 2 * b] = __float2bfloat16_rn(c_nf4[byte >> 4] * am);
        tile[r][cv + 2 * b + 1] = __float2bfloat16_rn(c_nf4[byte & 15u] * am);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[r][cv + j] = __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  // write: out row = c0 + c, 64 consecutive r values; thread t writes column c=(t/8)+32p, r=(t%8)*8..+8
  for (int p = 0; p < 2; ++p) {
    const int c = (threadIdx.x >> 3) + 32 * p;
    const int rv = (threadIdx.x & 7) * 8;
    if (r0 + rv < rows) {  // rows % 8 == 0 checked on host
      bf16x8 o;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h;
        h.x = tile[rv + 2 * j][c];
        h.y = tile[rv + 2 * j + 1][c];
        o.set(j, h);
      }
      *reinterpret_cast<bf16x8*>(out + (long long)(c0 + c) * rows + r0 + rv) = o;
    }
  }
}

}  // namespace b200rl

using namespace b200rl;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int b200rl_nf4_quantize(const void* w_bf16, void* packed, float* absmax, long long n,
                                   void* stream) {
  B200RL_REQUIRE(w_bf16 && packed && absmax && n > 0 && n % 64 == 0,
                 "nf4_quantize: n must be a positive multiple of 64 (n=%lld)", n);
  const long long nblocks = n / 64;
  const int wpb = 8;
  nf4_quant_kernel<<<(unsigned)((nblocks + wpb - 1) / wpb), wpb * 32, 0, STREAM>>>(
      (const bf16*)w_bf16, (uint8_t*)packed, absmax, nblocks);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_nf4_dequant(const void* packed, const float* absmax, void* out_bf16, int rows,
                                  int cols, int transpose, void* stream) {
  B200RL_REQUIRE(packed && absmax && out_bf16 && rows > 0 && cols > 0 && cols % 64 == 0,
                 "nf4_dequant: cols must be a multiple of 64 (rows=%d cols=%d)", rows, cols);
  if (!transpose) {
    const long long n8 = (long long)rows * cols / 8;
    long long blocks = (n8 + 256 * 4 - 1) / (256 * 4);
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    nf4_dequant_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>((const uint8_t*)packed, absmax,
                                                                (bf16*)out_bf16, n8);
  } else {
    B200RL_REQUIRE(rows % 8 == 0, "nf4_dequant(transpose): rows must be a multiple of 8");
    dim3 grid(cols / 64, (rows + 63) / 64);
    nf4_dequant_t_kernel<<<grid, 256, 0, STREAM>>>((const uint8_t*)packed, absmax, (bf16*)out_bf16,
                                                   rows, cols);
  }
  B200RL_LAUNCH_OK();
  return 0;
}
