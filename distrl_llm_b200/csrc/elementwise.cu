// HBM-bound row kernels of the causal-LM forward/backward:
// embedding gather, RMSNorm fwd/bwd, RoPE (rotate-half), SwiGLU fwd/bwd.
// These replace the Unsloth/Triton kernels the reference reaches through `policy(...)`
// (reference distributed_actor.py:241-243) and their autograd backward (:385, :483).
// All use 16-byte vector accesses, one row per CTA (rows are 7-37 KB, re-reads hit L1/L2).
#include "common.cuh"

namespace b200rl {

// ------------------------------------------------------------------------------------------
// embedding gather: out[m, :] = table[ids[m], :]
// ------------------------------------------------------------------------------------------
__global__ void embed_kernel(const int* __restrict__ ids, const bf16* __restrict__ table,
                             bf16* __restrict__ out, int H, int vocab) {
  pdl_enter();
  const int m = blockIdx.x;
  int id = ids[m];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * H);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)m * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------
// RMSNorm: y = w * bf16(x * rsqrt(mean(x^2) + eps))   (HF Qwen2RMSNorm rounding order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  __syncthreads();
  return t;
}

__global__ void rmsnorm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                   bf16* __restrict__ y, float* __restrict__ rstd, int H,
                                   float eps) {
  pdl_enter();
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * H);
  float ss = 0.f;
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    float f[8];
    unpack8(xr[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
  ss = block_sum(ss, red);
  const float r = rsqrtf(ss / (float)H + eps);
  if (threadIdx.x == 0 && rstd) rstd[row] = r;
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  bf16x8* yr = reinterpret_cast<bf16x8*>(y + row * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    float f[8], g[8];
    unpack8(xr[i], f);
    unpack8(wr[i], g);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = g[j] * __bfloat162float(__float2bfloat16_rn(f[j] * r));
    yr[i] = pack8(f);
  }
}

// dx = dres + rstd * ( w*dy - xhat * mean(w*dy*xhat) ),  xhat = x*rstd.  (norm weight frozen)
__global__ void rmsnorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                   const bf16* __restrict__ w, const float* __restrict__ rstd,
                                   const bf16* __restrict__ dres, bf16* __restrict__ dx, int H) {
  pdl_enter();
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const bf16x8* xr = reinterpret_cast<const bf16x8*>(x + row * H);
  const bf16x8* dyr = reinterpret_cast<const bf16x8*>(dy + row * H);
  const bf16x8* wr = reinterpret_cast<const bf16x8*>(w);
  const float r = rstd[row];
  float dot = 0.f;
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    float f[8], g[8], d[8];
    unpack8(xr[i], f);
    unpack8(wr[i], g);
    unpack8(dyr[i], d);
#pragma unroll
    for (int j = 0; j < 8; ++j) dot += g[j] * d[j] * f[j] * r;
  }
  dot = block_sum(dot, red) / (float)H;
  const bf16x8* rr = dres ? reinterpret_cast<const bf16x8*>(dres + row * H) : nullptr;
  bf16x8* dxr = reinterpret_cast<bf16x8*>(dx + row * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    float f[8], g[8], d[8], o[8];
    unpack8(xr[i], f);
    unpack8(wr[i], g);
    unpack8(dyr[i], d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = r * (g[j] * d[j] - f[j] * r * dot);
    if (rr) {
      float e[8];
      unpack8(rr[i], e);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += e[j];
    }
    dxr[i] = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------
// RoPE (HF rotate_half convention, head_dim D): in place on the first n_rot_heads heads of each
// row of qkv [M, row_stride]; position of row m is (m % L) (reference passes no position_ids, so
// HF uses arange(L) even under left padding: transformers Qwen2Model.forward).
//   out[i]      = x[i]*cos - x[i+D/2]*sin
//   out[i+D/2]  = x[i+D/2]*cos + x[i]*sin       (sign = -1 gives the transpose = backward)
// ------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(float* __restrict__ cs, int L, int half, float theta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L * half) return;
  const int pos = idx / half, i = idx % half;
  // inv_freq = theta^(-2i/D) computed like torch: 1 / (theta ** (arange(0, D, 2) / D)) in fp32
  const float inv_freq = 1.0f / powf(theta, (float)(2 * i) / (float)(2 * half));
  const float ang = (float)pos * inv_freq;
  float s, c;
  sincosf(ang, &s, &c);
  cs[2 * idx] = c;
  cs[2 * idx + 1] = s;
}

__global__ void rope_kernel(bf16* __restrict__ qkv, const float* __restrict__ cs, int L,
                            long long row_stride, int n_rot_heads, int D, float sign,
                            const int* __restrict__ pos_idx) {
  pdl_enter();
  const size_t m = blockIdx.x;
  const int pos = pos_idx ? pos_idx[m] : (int)(m % L);  // packed layout carries explicit positions
  const int half = D / 2;
  bf16* row = qkv + m * row_stride;
  const float2* tab = reinterpret_cast<const float2*>(cs) + (size_t)pos * half;
  // each thread handles 8 consecutive i of one head
  const int per_head = half / 8;
  for (int t = threadIdx.x; t < n_rot_heads * per_head; t += blockDim.x) {
    const int h = t / per_head, i0 = (t % per_head) * 8;
    bf16x8* plo = reinterpret_cast<bf16x8*>(row + h * D + i0);
    bf16x8* phi = reinterpret_cast<bf16x8*>(row + h * D + half + i0);
    float lo[8], hi[8], olo[8], ohi[8];
    unpack8(*plo, lo);
    unpack8(*phi, hi);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 c = tab[i0 + j];
      const float s = c.y * sign;
      olo[j] = lo[j] * c.x - hi[j] * s;
      ohi[j] = hi[j] * c.x + lo[j] * s;
    }
    *plo = pack8(olo);
    *phi = pack8(ohi);
  }
}

// ------------------------------------------------------------------------------------------
// SwiGLU on a fused [M, 2I] gate|up buffer
// ------------------------------------------------------------------------------------------
__global__ void swiglu_fwd_kernel(const bf16* __restrict__ gu, bf16* __restrict__ act, int I) {
  pdl_enter();
  const size_t row = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= I / 8) return;
  const bf16x8* g = reinterpret_cast<const bf16x8*>(gu + row * 2 * I);
  const bf16x8* u = reinterpret_cast<const bf16x8*>(gu + row * 2 * I + I);
  float a[8], b[8], o[8];
  unpack8(g[i], a);
  unpack8(u[i], b);
  swiglu_fwd8(a, b, o);
  reinterpret_cast<bf16x8*>(act + row * I)[i] = pack8(o);
}

// dgu[:, :I] = dact * up * silu'(gate) ; dgu[:, I:] = dact * silu(gate)
__global__ void swiglu_bwd_kernel(const bf16* __restrict__ gu, const bf16* __restrict__ dact,
                                  bf16* __restrict__ dgu, int I) {
  pdl_enter();
  const size_t row = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= I / 8) return;
  const bf16x8* g = reinterpret_cast<const bf16x8*>(gu + row * 2 * I);
  const bf16x8* u = reinterpret_cast<const bf16x8*>(gu + row * 2 * I + I);
  const bf16x8* d = reinterpret_cast<const bf16x8*>(dact + row * I);
  float a[8], b[8], c[8], og[8], ou[8];
  unpack8(g[i], a);
  unpack8(u[i], b);
  unpack8(d[i], c);
  swiglu_bwd8(a, b, c, og, ou);
  reinterpret_cast<bf16x8*>(dgu + row * 2 * I)[i] = pack8(og);
  reinterpret_cast<bf16x8*>(dgu + row * 2 * I + I)[i] = pack8(ou);
}

// select rows: out[b*T + t, :] = x[b*L + start + t, :]   (completion positions P-1 .. L-2)
__global__ void gather_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int L,
                                   int T, int start) {
  pdl_enter();
  const int r = blockIdx.x;
  const int b = r / T, t = r % T;
  const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * L + start + t) * H);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)r * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}
// packed layout: out[r, :] = x[src[r], :]
__global__ void gather_rows_idx_kernel(const bf16* __restrict__ x, const int* __restrict__ src,
                                       bf16* __restrict__ out, int H) {
  pdl_enter();
  const int r = blockIdx.x;
  const uint4* s4 = reinterpret_cast<const uint4*>(x + (size_t)src[r] * H);
  uint4* d4 = reinterpret_cast<uint4*>(out + (size_t)r * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) d4[i] = s4[i];
}
// packed layout: dx[m, :] = sum over the scored rows r whose hidden state is row m (CSR start/list, fixed order);
// rows without a scored position get zeros.  The shared last prompt row collects one term per completion.
__global__ void scatter_add_rows_kernel(const bf16* __restrict__ d, const int* __restrict__ start,
                                        const int* __restrict__ list, bf16* __restrict__ dx, int H) {
  pdl_enter();
  const int m = blockIdx.x;
  const int s0 = start[m], s1 = start[m + 1];
  bf16x8* dst = reinterpret_cast<bf16x8*>(dx + (size_t)m * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = s0; s < s1; ++s) {
      float f[8];
      unpack8(reinterpret_cast<const bf16x8*>(d + (size_t)list[s] * H)[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    dst[i] = pack8(acc);
  }
}
// scatter back (zero elsewhere): dx[b*L + start + t, :] = d[b*T + t, :], other rows 0
__global__ void scatter_rows_kernel(const bf16* __restrict__ d, bf16* __restrict__ dx, int H, int L,
                                    int T, int start) {
  pdl_enter();
  const int m = blockIdx.x;
  const int b = m / L, pos = m % L;
  uint4* dst = reinterpret_cast<uint4*>(dx + (size_t)m * H);
  const int t = pos - start;
  if (t >= 0 && t < T) {
    const uint4* src = reinterpret_cast<const uint4*>(d + ((size_t)b * T + t) * H);
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
  }
}

}  // namespace b200rl

using namespace b200rl;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int b200rl_embed(const int* ids, const void* table, void* out, int M, int H, int vocab,
                            void* stream) {
  B200RL_REQUIRE(ids && table && out && M > 0 && H % 8 == 0, "embed: bad args (M=%d H=%d)", M, H);
  B200RL_CUDA_OK(launch_pdl(embed_kernel, dim3(M), dim3(128), 0, STREAM, ids, (const bf16*)table, (bf16*)out, H, vocab));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int M, int H,
                                  float eps, void* stream) {
  B200RL_REQUIRE(x && w && y && M > 0 && H % 8 == 0, "rmsnorm_fwd: bad args (M=%d H=%d)", M, H);
  B200RL_CUDA_OK(launch_pdl(rmsnorm_fwd_kernel, dim3(M), dim3(256), 0, STREAM, (const bf16*)x, (const bf16*)w, (bf16*)y, rstd, H, eps));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd,
                                  const void* dres, void* dx, int M, int H, void* stream) {
  B200RL_REQUIRE(dy && x && w && rstd && dx && M > 0 && H % 8 == 0,
                 "rmsnorm_bwd: bad args (M=%d H=%d)", M, H);
  B200RL_CUDA_OK(launch_pdl(rmsnorm_bwd_kernel, dim3(M), dim3(256), 0, STREAM, (const bf16*)dy, (const bf16*)x, (const bf16*)w, rstd,
                                            (const bf16*)dres, (bf16*)dx, H));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_rope_table(float* cs, int L, int head_dim, float theta, void* stream) {
  B200RL_REQUIRE(cs && L > 0 && head_dim % 16 == 0, "rope_table: bad args");
  const int n = L * head_dim / 2;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, STREAM>>>(cs, L, head_dim / 2, theta);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_rope(void* qkv, const float* cs, int M, int L, long long row_stride,
                           int n_rot_heads, int head_dim, int backward, void* stream) {
  B200RL_REQUIRE(qkv && cs && M > 0 && L > 0 && head_dim % 16 == 0 && row_stride % 8 == 0,
                 "rope: bad args");
  B200RL_CUDA_OK(launch_pdl(rope_kernel, dim3(M), dim3(128), 0, STREAM, (bf16*)qkv, cs, L, row_stride, n_rot_heads, head_dim,
                                     backward ? -1.f : 1.f, nullptr));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_rope_pos(void* qkv, const float* cs, const int* pos, int M, long long row_stride,
                               int n_rot_heads, int head_dim, int backward, void* stream) {
  B200RL_REQUIRE(qkv && cs && pos && M > 0 && head_dim % 16 == 0 && row_stride % 8 == 0, "rope_pos: bad args");
  B200RL_CUDA_OK(launch_pdl(rope_kernel, dim3(M), dim3(128), 0, STREAM, (bf16*)qkv, cs, 1, row_stride, n_rot_heads, head_dim, backward ? -1.f : 1.f, pos));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_gather_rows_idx(const void* x, const int* src, void* out, int R, int H, void* stream) {
  B200RL_REQUIRE(x && src && out && R > 0 && H % 8 == 0, "gather_rows_idx: bad args");
  B200RL_CUDA_OK(launch_pdl(gather_rows_idx_kernel, dim3(R), dim3(128), 0, STREAM, (const bf16*)x, src, (bf16*)out, H));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_scatter_add_rows(const void* d, const int* start, const int* list, void* dx, int M, int H,
                                       void* stream) {
  B200RL_REQUIRE(d && start && list && dx && M > 0 && H % 8 == 0, "scatter_add_rows: bad args");
  B200RL_CUDA_OK(launch_pdl(scatter_add_rows_kernel, dim3(M), dim3(128), 0, STREAM, (const bf16*)d, start, list, (bf16*)dx, H));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_swiglu_fwd(const void* gu, void* act, int M, int I, void* stream) {
  B200RL_REQUIRE(gu && act && M > 0 && I % 8 == 0, "swiglu_fwd: bad args");
  dim3 grid((I / 8 + 255) / 256, M);
  B200RL_CUDA_OK(launch_pdl(swiglu_fwd_kernel, dim3(grid), dim3(256), 0, STREAM, (const bf16*)gu, (bf16*)act, I));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_swiglu_bwd(const void* gu, const void* dact, void* dgu, int M, int I,
                                 void* stream) {
  B200RL_REQUIRE(gu && dact && dgu && M > 0 && I % 8 == 0, "swiglu_bwd: bad args");
  dim3 grid((I / 8 + 255) / 256, M);
  B200RL_CUDA_OK(launch_pdl(swiglu_bwd_kernel, dim3(grid), dim3(256), 0, STREAM, (const bf16*)gu, (const bf16*)dact, (bf16*)dgu, I));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_gather_rows(const void* x, void* out, int B, int L, int T, int start, int H,
                                  void* stream) {
  B200RL_REQUIRE(x && out && B > 0 && T > 0 && start >= 0 && start + T <= L && H % 8 == 0,
                 "gather_rows: bad args");
  B200RL_CUDA_OK(launch_pdl(gather_rows_kernel, dim3(B * T), dim3(128), 0, STREAM, (const bf16*)x, (bf16*)out, H, L, T, start));
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_scatter_rows(const void* d, void* dx, int B, int L, int T, int start, int H,
                                   void* stream) {
  B200RL_REQUIRE(d && dx && B > 0 && T > 0 && start >= 0 && start + T <= L && H % 8 == 0,
                 "scatter_rows: bad args");
  B200RL_CUDA_OK(launch_pdl(scatter_rows_kernel, dim3(B * L), dim3(128), 0, STREAM, (const bf16*)d, (bf16*)dx, H, L, T, start));
  B200RL_LAUNCH_OK();
  return 0;
}
