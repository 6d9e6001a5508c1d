// G4: causal GQA attention with key-padding mask, forward + backward (flash-style, no [L,L] tensor
// in HBM).  Replaces the xformers / SDPA attention the reference reaches inside `policy(...)`
// (distributed_actor.py:241-243) with attention_mask = cat(prompt_mask, answer_mask) (:236-239),
// and its autograd backward (:385, :483).
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate), 64-row tiles,
// online softmax in the log2 domain.  Attention is ~2-6 % of the step's FLOPs, so
// the tcgen05/TMEM rewrite of this kernel is scheduled after the GEMMs.
// Backward is split in two deterministic kernels (dQ by query block; dK/dV by key block, summed
// over the GQA group inside the CTA) instead of one kernel with fp32 atomics.
//
// Layout: qkv [B*L, (nq+2nkv)*hd] bf16 (q heads | k heads | v heads per row, RoPE already applied),
// out/dout [B*L, nq*hd] bf16, lse2 [B, nq, L] fp32 = log2-sum-exp2 of the scaled scores
// (+inf for rows with no visible key, which makes every probability of that row exactly 0).
#include "common.cuh"
#include "b200rl.h"
#include <stdlib.h>

namespace b200rl {

static constexpr float LOG2E_F = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                          uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Load a [ROWS x HD] bf16 tile (global row stride gstride elements) into padded smem [ROWS][HD+8];
// rows >= nrows_valid are zero-filled.
template <int HD, int ROWS>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, long long gstride, int nrows_valid) {
  constexpr int LD = HD + 8;
  constexpr int VPR = HD / 8;
  for (int v = threadIdx.x; v < ROWS * VPR; v += blockDim.x) {
    const int r = v / VPR, c = (v % VPR) * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < nrows_valid) val = *reinterpret_cast<const uint4*>(g + (long long)r * gstride + c);
    *reinterpret_cast<uint4*>(s + r * LD + c) = val;
  }
}

// Same tile, asynchronously (cp.async 16 B, zero-fill for rows >= nrows_valid); pair with cp_async_commit /
// cp_async_wait<N> + __syncthreads.  Streams the K/V (or Q/dO) tiles one iteration ahead of the math.
template <int HD, int ROWS>
__device__ __forceinline__ void load_tile_async(bf16* s, const bf16* g, long long gstride, int nrows_valid) {
  constexpr int LD = HD + 8;
  constexpr int VPR = HD / 8;
  for (int v = threadIdx.x; v < ROWS * VPR; v += blockDim.x) {
    const int r = v / VPR, c = (v % VPR) * 8;
    const bool ok = r < nrows_valid;
    const bf16* src = ok ? g + (long long)r * gstride + c : g;
    const uint32_t dst = smem_u32(s + r * LD + c);
    const int nbytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// A-operand fragment (16 rows x 16 k) from a row-major padded tile: rows r0.., k columns k0..
template <int LD>
__device__ __forceinline__ void load_a_frag(const bf16* s, int r0, int k0, uint32_t* a) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr = smem_u32(s + (r0 + (lane & 15)) * LD + k0 + (lane >> 4) * 8);
  ldsm_x4(addr, a[0], a[1], a[2], a[3]);
}
// B fragments for C += A . T^T where T is a row-major tile [n][k]: two adjacent 8-wide n-tiles
// (n0..n0+15) at k columns k0..k0+15.  b[0],b[1] -> n-tile 0, b[2],b[3] -> n-tile 1.
template <int LD>
__device__ __forceinline__ void load_b_frag_nt(const bf16* s, int n0, int k0, uint32_t* b) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr =
      smem_u32(s + (n0 + (lane & 7) + ((lane >> 4) << 3)) * LD + k0 + ((lane >> 3) & 1) * 8);
  ldsm_x4(addr, b[0], b[1], b[2], b[3]);
}
// B fragments for C += A . T where T is a row-major tile [k][n]: k rows k0..k0+15, two adjacent
// 8-wide n-tiles n0..n0+15 (transposing ldmatrix).
template <int LD>
__device__ __forceinline__ void load_b_frag_nn(const bf16* s, int k0, int n0, uint32_t* b) {
  const int lane = threadIdx.x & 31;
  const uint32_t addr =
      smem_u32(s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + n0 + (lane >> 4) * 8);
  ldsm_x4_t(addr, b[0], b[1], b[2], b[3]);
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const bf16* __restrict__ qkv, const int* __restrict__ key_mask,
                bf16* __restrict__ out, float* __restrict__ lse2, int L, int nq, int nkv,
                float scale_log2) {
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sKb = sQ + 64 * LD;            // [2][64][LD]
  bf16* sVb = sKb + 2 * 64 * LD;       // [2][64][LD]
  int* sMaskb = reinterpret_cast<int*>(sVb + 2 * 64 * LD);  // [2][64]

  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (nq / nkv);
  const long long stride = (long long)(nq + 2 * nkv) * HD;
  const bf16* base = qkv + (long long)b * L * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;

  load_tile<HD, 64>(sQ, base + (long long)q0 * stride + h * HD, stride, min(64, L - q0));
  __syncthreads();
  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) load_a_frag<LD>(sQ, warp * 16, kk * 16, qf[kk]);

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int row_a = q0 + warp * 16 + gq, row_b = row_a + 8;

  const int kb_end = min((q0 + 63) / 64, (L - 1) / 64);
  auto prefetch = [&](int kb) {
    const int k0p = kb * 64, buf = kb & 1;
    load_tile_async<HD, 64>(sKb + buf * 64 * LD, base + (long long)k0p * stride + (nq + g) * HD, stride, min(64, L - k0p));
    load_tile_async<HD, 64>(sVb + buf * 64 * LD, base + (long long)k0p * stride + (nq + nkv + g) * HD, stride, min(64, L - k0p));
    if (threadIdx.x < 64)
      sMaskb[buf * 64 + threadIdx.x] = (k0p + threadIdx.x < L) ? key_mask[(long long)b * L + k0p + threadIdx.x] : 0;
    cp_async_commit();
  };
  prefetch(0);
  for (int kb = 0; kb <= kb_end; ++kb) {
    const int k0 = kb * 64;
    if (kb < kb_end) {
      prefetch(kb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const bf16* sK = sKb + (kb & 1) * 64 * LD;
    const bf16* sV = sVb + (kb & 1) * 64 * LD;
    const int* sMask = sMaskb + (kb & 1) * 64;

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bfr[4];
        load_b_frag_nt<LD>(sK, np * 16, kk * 16, bfr);
        mma16816(s[2 * np], qf[kk], bfr[0], bfr[1]);
        mma16816(s[2 * np + 1], qf[kk], bfr[2], bfr[3]);
      }
    }
    // mask + online softmax (log2 domain)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = k0 + nt * 8 + 2 * tq + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        const bool ok = (col <= row) && sMask[nt * 8 + 2 * tq + (e & 1)] != 0;
        const float v = ok ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], mu[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      mu[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = exp2f(m_run[r] - mu[r]);  // m_run = -inf -> 0
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[nt][e] - mu[e >> 1]);
        s[nt][e] = p;
        rs[e >> 1] += p;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    // O += P . V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        uint32_t bfr[4];
        load_b_frag_nn<LD>(sV, ks * 16, dp * 16, bfr);
        mma16816(o[2 * dp], pa, bfr[0], bfr[1]);
        mma16816(o[2 * dp + 1], pa, bfr[2], bfr[3]);
      }
    }
    __syncthreads();  // everyone is done with this buffer before the next prefetch overwrites it
  }
  // finalize
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    inv[r] = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
  }
  if (tq == 0) {
    if (row_a < L)
      lse2[((long long)b * nq + h) * L + row_a] = l_run[0] > 0.f ? m_run[0] + log2f(l_run[0]) : INFINITY;
    if (row_b < L)
      lse2[((long long)b * nq + h) * L + row_b] = l_run[1] > 0.f ? m_run[1] + log2f(l_run[1]) : INFINITY;
  }
  // stage O through this warp's rows of sQ, then coalesced 16-byte stores
  __syncwarp();
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + gq) * LD + i * 8 + 2 * tq) =
        pack_bf16(o[i][0] * inv[0], o[i][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + gq + 8) * LD + i * 8 + 2 * tq) =
        pack_bf16(o[i][2] * inv[1], o[i][3] * inv[1]);
  }
  __syncthreads();
  constexpr int VPR = HD / 8;
  for (int v = threadIdx.x; v < 64 * VPR; v += blockDim.x) {
    const int r = v / VPR, c = (v % VPR) * 8;
    if (q0 + r < L)
      *reinterpret_cast<uint4*>(out + ((long long)b * L + q0 + r) * nq * HD + h * HD + c) =
          *reinterpret_cast<const uint4*>(sQ + r * LD + c);
  }
}

// ------------------------------------------------------------------------------------------
// backward: delta[b,h,i] = sum_d dO[i,d] * O[i,d]
// ------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout,
                                  float* __restrict__ delta, int L, int nq, int hd, long long rows) {
  pdl_enter();
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (row, head)
  if (w >= rows * nq) return;
  const long long row = w / nq;
  const int h = (int)(w % nq);
  const int lane = threadIdx.x & 31;
  const bf16* po = out + row * nq * hd + h * hd;
  const bf16* pd = dout + row * nq * hd + h * hd;
  float acc = 0.f;
  for (int c = lane * 4; c < hd; c += 128) {  // 8-byte loads: every lane is active for hd = 128
    const uint2 a2 = *reinterpret_cast<const uint2*>(po + c);
    const uint2 d2 = *reinterpret_cast<const uint2*>(pd + c);
    acc += __uint_as_float(a2.x << 16) * __uint_as_float(d2.x << 16) +
           __uint_as_float(a2.x & 0xFFFF0000u) * __uint_as_float(d2.x & 0xFFFF0000u) +
           __uint_as_float(a2.y << 16) * __uint_as_float(d2.y << 16) +
           __uint_as_float(a2.y & 0xFFFF0000u) * __uint_as_float(d2.y & 0xFFFF0000u);
  }
  acc = warp_sum(acc);
  const long long b = row / L, i = row % L;
  if (lane == 0) delta[(b * nq + h) * L + i] = acc;
}

// ------------------------------------------------------------------------------------------
// backward: dQ (one CTA per 64-query block and q head)
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const int* __restrict__ key_mask,
                   const bf16* __restrict__ dout, const float* __restrict__ lse2,
                   const float* __restrict__ delta, bf16* __restrict__ dqkv, int L, int nq, int nkv,
                   float scale, float scale_log2) {
  constexpr int LD = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sdO = sQ + 64 * LD;
  bf16* sKb = sdO + 64 * LD;           // [2][64][LD]
  bf16* sVb = sKb + 2 * 64 * LD;       // [2][64][LD]
  int* sMaskb = reinterpret_cast<int*>(sVb + 2 * 64 * LD);

  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int g = h / (nq / nkv);
  const long long stride = (long long)(nq + 2 * nkv) * HD;
  const bf16* base = qkv + (long long)b * L * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int nvalid_q = min(64, L - q0);

  load_tile<HD, 64>(sQ, base + (long long)q0 * stride + h * HD, stride, nvalid_q);
  load_tile<HD, 64>(sdO, dout + ((long long)b * L + q0) * nq * HD + h * HD, (long long)nq * HD, nvalid_q);

  const int row_a = q0 + warp * 16 + gq, row_b = row_a + 8;
  const long long sidx = ((long long)b * nq + h) * L;
  const float lse_a = row_a < L ? lse2[sidx + row_a] : INFINITY;
  const float lse_b = row_b < L ? lse2[sidx + row_b] : INFINITY;
  const float del_a = row_a < L ? delta[sidx + row_a] : 0.f;
  const float del_b = row_b < L ? delta[sidx + row_b] : 0.f;

  float dq[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  const int kb_end = min((q0 + 63) / 64, (L - 1) / 64);
  auto prefetch = [&](int kb) {
    const int k0p = kb * 64, buf = kb & 1;
    load_tile_async<HD, 64>(sKb + buf * 64 * LD, base + (long long)k0p * stride + (nq + g) * HD, stride, min(64, L - k0p));
    load_tile_async<HD, 64>(sVb + buf * 64 * LD, base + (long long)k0p * stride + (nq + nkv + g) * HD, stride, min(64, L - k0p));
    if (threadIdx.x < 64)
      sMaskb[buf * 64 + threadIdx.x] = (k0p + threadIdx.x < L) ? key_mask[(long long)b * L + k0p + threadIdx.x] : 0;
    cp_async_commit();
  };
  prefetch(0);
  for (int kb = 0; kb <= kb_end; ++kb) {
    const int k0 = kb * 64;
    if (kb < kb_end) {
      prefetch(kb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();  // also publishes the sQ / sdO tiles on the first iteration
    const bf16* sK = sKb + (kb & 1) * 64 * LD;
    const bf16* sV = sVb + (kb & 1) * 64 * LD;
    const int* sMask = sMaskb + (kb & 1) * 64;

    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t qa[4], da[4];
      load_a_frag<LD>(sQ, warp * 16, kk * 16, qa);
      load_a_frag<LD>(sdO, warp * 16, kk * 16, da);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bk[4], bv[4];
        load_b_frag_nt<LD>(sK, np * 16, kk * 16, bk);
        load_b_frag_nt<LD>(sV, np * 16, kk * 16, bv);
        mma16816(s[2 * np], qa, bk[0], bk[1]);
        mma16816(s[2 * np + 1], qa, bk[2], bk[3]);
        mma16816(dp[2 * np], da, bv[0], bv[1]);
        mma16816(dp[2 * np + 1], da, bv[2], bv[3]);
      }
    }
    // dS = scale * P * (dP - delta)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = k0 + nt * 8 + 2 * tq + (e & 1);
        const int row = (e < 2) ? row_a : row_b;
        const bool ok = (col <= row) && sMask[nt * 8 + 2 * tq + (e & 1)] != 0;
        const float p = ok ? exp2f(s[nt][e] * scale_log2 - ((e < 2) ? lse_a : lse_b)) : 0.f;
        s[nt][e] = scale * p * (dp[nt][e] - ((e < 2) ? del_a : del_b));
      }
    }
    // dQ += dS . K
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
      pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
      pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < HD / 16; ++d2) {
        uint32_t bfr[4];
        load_b_frag_nn<LD>(sK, ks * 16, d2 * 16, bfr);
        mma16816(dq[2 * d2], pa, bfr[0], bfr[1]);
        mma16816(dq[2 * d2 + 1], pa, bfr[2], bfr[3]);
      }
    }
    __syncthreads();  // buffer free for the next prefetch
  }
  __syncthreads();  // everyone is done with sQ as an operand
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + gq) * LD + i * 8 + 2 * tq) = pack_bf16(dq[i][0], dq[i][1]);
    *reinterpret_cast<uint32_t*>(sQ + (warp * 16 + gq + 8) * LD + i * 8 + 2 * tq) = pack_bf16(dq[i][2], dq[i][3]);
  }
  __syncthreads();
  constexpr int VPR = HD / 8;
  for (int v = threadIdx.x; v < 64 * VPR; v += blockDim.x) {
    const int r = v / VPR, c = (v % VPR) * 8;
    if (q0 + r < L)
      *reinterpret_cast<uint4*>(dqkv + ((long long)b * L + q0 + r) * stride + h * HD + c) =
          *reinterpret_cast<const uint4*>(sQ + r * LD + c);
  }
}

// ------------------------------------------------------------------------------------------
// backward: dK, dV (one CTA per 64-key block and kv head; loops over the GQA group's q heads and
// over 32-query blocks at or after the key block)
// ------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const int* __restrict__ key_mask,
                    const bf16* __restrict__ dout, const float* __restrict__ lse2,
                    const float* __restrict__ delta, bf16* __restrict__ dqkv, int L, int nq, int nkv,
                    float scale, float scale_log2) {
  constexpr int LD = HD + 8;
  constexpr int BQ = 32;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sK = reinterpret_cast<bf16*>(smem_attn);
  bf16* sV = sK + 64 * LD;
  bf16* sQb = sV + 64 * LD;            // [2][BQ][LD]
  bf16* sdOb = sQb + 2 * BQ * LD;      // [2][BQ][LD]
  float* sLseb = reinterpret_cast<float*>(sdOb + 2 * BQ * LD);  // [2][BQ]
  float* sDelb = sLseb + 2 * BQ;                                 // [2][BQ]

  const int k0 = blockIdx.x * 64, g = blockIdx.y, b = blockIdx.z;
  const int group = nq / nkv;
  const long long stride = (long long)(nq + 2 * nkv) * HD;
  const bf16* base = qkv + (long long)b * L * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int nvalid_k = min(64, L - k0);

  load_tile<HD, 64>(sK, base + (long long)k0 * stride + (nq + g) * HD, stride, nvalid_k);
  load_tile<HD, 64>(sV, base + (long long)k0 * stride + (nq + nkv + g) * HD, stride, nvalid_k);
  const int key_a = k0 + warp * 16 + gq, key_b = key_a + 8;
  const bool km_a = key_a < L && key_mask[(long long)b * L + key_a] != 0;
  const bool km_b = key_b < L && key_mask[(long long)b * L + key_b] != 0;

  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  // flattened (q head of the group, 32-query block at or after the key block) iteration space,
  // Q / dO / lse / delta tiles streamed one iteration ahead with cp.async
  const int qb0 = k0 / BQ;
  const int nqb = (L - qb0 * BQ + BQ - 1) / BQ;
  const int n_it = group * nqb;
  auto prefetch = [&](int it) {
    const int h = g * group + it / nqb;
    const int q0p = (qb0 + it % nqb) * BQ;
    const int buf = it & 1;
    const int nv = min(BQ, L - q0p);
    load_tile_async<HD, BQ>(sQb + buf * BQ * LD, base + (long long)q0p * stride + h * HD, stride, nv);
    load_tile_async<HD, BQ>(sdOb + buf * BQ * LD, dout + ((long long)b * L + q0p) * nq * HD + h * HD, (long long)nq * HD, nv);
    if (threadIdx.x < BQ) {
      const int qi = q0p + threadIdx.x;
      const long long sidx = ((long long)b * nq + h) * L;
      sLseb[buf * BQ + threadIdx.x] = qi < L ? lse2[sidx + qi] : INFINITY;
      sDelb[buf * BQ + threadIdx.x] = qi < L ? delta[sidx + qi] : 0.f;
    }
    cp_async_commit();
  };
  prefetch(0);
  for (int it = 0; it < n_it; ++it) {
    {
      const int q0 = (qb0 + it % nqb) * BQ;
      if (it + 1 < n_it) {
        prefetch(it + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();  // also publishes sK / sV on the first iteration
      const bf16* sQ = sQb + (it & 1) * BQ * LD;
      const bf16* sdO = sdOb + (it & 1) * BQ * LD;
      const float* sLse = sLseb + (it & 1) * BQ;
      const float* sDel = sDelb + (it & 1) * BQ;

      // S^T = K . Q^T and dP^T = V . dO^T : [16 keys x 32 queries] per warp
      float st[4][4], dpt[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t ka[4], va[4];
        load_a_frag<LD>(sK, warp * 16, kk * 16, ka);
        load_a_frag<LD>(sV, warp * 16, kk * 16, va);
#pragma unroll
        for (int np = 0; np < BQ / 16; ++np) {
          uint32_t bq[4], bd[4];
          load_b_frag_nt<LD>(sQ, np * 16, kk * 16, bq);
          load_b_frag_nt<LD>(sdO, np * 16, kk * 16, bd);
          mma16816(st[2 * np], ka, bq[0], bq[1]);
          mma16816(st[2 * np + 1], ka, bq[2], bq[3]);
          mma16816(dpt[2 * np], va, bd[0], bd[1]);
          mma16816(dpt[2 * np + 1], va, bd[2], bd[3]);
        }
      }
      // P^T and dS^T
      float pt[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ql = nt * 8 + 2 * tq + (e & 1);
          const int qi = q0 + ql;
          const int key = (e < 2) ? key_a : key_b;
          const bool ok = ((e < 2) ? km_a : km_b) && key <= qi && qi < L;
          const float p = ok ? exp2f(st[nt][e] * scale_log2 - sLse[ql]) : 0.f;
          pt[nt][e] = p;
          st[nt][e] = scale * p * (dpt[nt][e] - sDel[ql]);
        }
      }
      // dV += P^T . dO ; dK += dS^T . Q     (reduction over the 32 queries = 2 k-steps)
#pragma unroll
      for (int ks = 0; ks < BQ / 16; ++ks) {
        uint32_t pa[4], sa[4];
        pa[0] = pack_bf16(pt[2 * ks][0], pt[2 * ks][1]);
        pa[1] = pack_bf16(pt[2 * ks][2], pt[2 * ks][3]);
        pa[2] = pack_bf16(pt[2 * ks + 1][0], pt[2 * ks + 1][1]);
        pa[3] = pack_bf16(pt[2 * ks + 1][2], pt[2 * ks + 1][3]);
        sa[0] = pack_bf16(st[2 * ks][0], st[2 * ks][1]);
        sa[1] = pack_bf16(st[2 * ks][2], st[2 * ks][3]);
        sa[2] = pack_bf16(st[2 * ks + 1][0], st[2 * ks + 1][1]);
        sa[3] = pack_bf16(st[2 * ks + 1][2], st[2 * ks + 1][3]);
#pragma unroll
        for (int d2 = 0; d2 < HD / 16; ++d2) {
          uint32_t bd[4], bq[4];
          load_b_frag_nn<LD>(sdO, ks * 16, d2 * 16, bd);
          load_b_frag_nn<LD>(sQ, ks * 16, d2 * 16, bq);
          mma16816(dv[2 * d2], pa, bd[0], bd[1]);
          mma16816(dv[2 * d2 + 1], pa, bd[2], bd[3]);
          mma16816(dk[2 * d2], sa, bq[0], bq[1]);
          mma16816(dk[2 * d2 + 1], sa, bq[2], bq[3]);
        }
      }
      __syncthreads();  // buffer free for the next prefetch
    }
  }
  // write dK, dV through smem (reuse sK / sV: all operand reads are finished after this barrier)
  __syncthreads();
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + gq) * LD + i * 8 + 2 * tq) = pack_bf16(dk[i][0], dk[i][1]);
    *reinterpret_cast<uint32_t*>(sK + (warp * 16 + gq + 8) * LD + i * 8 + 2 * tq) = pack_bf16(dk[i][2], dk[i][3]);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + gq) * LD + i * 8 + 2 * tq) = pack_bf16(dv[i][0], dv[i][1]);
    *reinterpret_cast<uint32_t*>(sV + (warp * 16 + gq + 8) * LD + i * 8 + 2 * tq) = pack_bf16(dv[i][2], dv[i][3]);
  }
  __syncthreads();
  constexpr int VPR = HD / 8;
  for (int v = threadIdx.x; v < 64 * VPR; v += blockDim.x) {
    const int r = v / VPR, c = (v % VPR) * 8;
    if (k0 + r < L) {
      bf16* drow = dqkv + ((long long)b * L + k0 + r) * stride;
      *reinterpret_cast<uint4*>(drow + (nq + g) * HD + c) = *reinterpret_cast<const uint4*>(sK + r * LD + c);
      *reinterpret_cast<uint4*>(drow + (nq + nkv + g) * HD + c) = *reinterpret_cast<const uint4*>(sV + r * LD + c);
    }
  }
}

template <int HD>
static int attn_fwd_launch(const void* qkv, const int* key_mask, void* out, float* lse, int B, int L,
                           int nq, int nkv, float scale, cudaStream_t stream) {
  constexpr int LD = HD + 8;
  const int smem = 5 * 64 * LD * 2 + 2 * 64 * 4;
  auto kern = attn_fwd_kernel<HD>;
  B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((L + 63) / 64, nq, B);
  kern<<<grid, 128, smem, stream>>>((const bf16*)qkv, key_mask, (bf16*)out, lse, L, nq, nkv,
                                    scale * LOG2E_F);
  B200RL_LAUNCH_OK();
  return 0;
}

template <int HD>
static int attn_bwd_launch(const void* qkv, const int* key_mask, const void* out, const void* dout,
                           const float* lse, float* delta, void* dqkv, int B, int L, int nq, int nkv,
                           float scale, cudaStream_t stream) {
  constexpr int LD = HD + 8;
  const long long rows = (long long)B * L;
  {
    const long long warps = rows * nq;
    B200RL_CUDA_OK(launch_pdl(attn_delta_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, (const bf16*)out, (const bf16*)dout,
                                                                        delta, L, nq, HD, rows));
    B200RL_LAUNCH_OK();
  }
  {
    const int smem = 6 * 64 * LD * 2 + 2 * 64 * 4;
    auto kern = attn_bwd_dq_kernel<HD>;
    B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((L + 63) / 64, nq, B);
    kern<<<grid, 128, smem, stream>>>((const bf16*)qkv, key_mask, (const bf16*)dout, lse, delta,
                                      (bf16*)dqkv, L, nq, nkv, scale, scale * LOG2E_F);
    B200RL_LAUNCH_OK();
  }
  {
    const int smem = (2 * 64 + 4 * 32) * LD * 2 + 4 * 32 * 4;
    auto kern = attn_bwd_dkv_kernel<HD>;
    B200RL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((L + 63) / 64, nkv, B);
    kern<<<grid, 128, smem, stream>>>((const bf16*)qkv, key_mask, (const bf16*)dout, lse, delta,
                                      (bf16*)dqkv, L, nq, nkv, scale, scale * LOG2E_F);
    B200RL_LAUNCH_OK();
  }
  return 0;
}

// tcgen05 forward (attention_tc.cu), head_dim 128
int attn_fwd_tc_launch(const void* qkv, const int* key_mask, void* out, float* lse, int B, int L, int nq, int nkv,
                       float scale, cudaStream_t stream);
int attn_bwd_tc_launch(const void* qkv, const int* key_mask, const void* dout, const float* lse, const float* delta,
                       void* dqkv, int B, int L, int nq, int nkv, float scale, cudaStream_t stream);
static int attn_bwd_tc(const void* qkv, const int* key_mask, const void* out, const void* dout, const float* lse,
                       float* delta, void* dqkv, int B, int L, int nq, int nkv, float scale, cudaStream_t stream) {
  const long long rows = (long long)B * L;
  const long long warps = rows * nq;
  B200RL_CUDA_OK(launch_pdl(attn_delta_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, stream, (const bf16*)out, (const bf16*)dout, delta, L, nq,
                                                                      128, rows));
  B200RL_LAUNCH_OK();
  return attn_bwd_tc_launch(qkv, key_mask, dout, lse, delta, dqkv, B, L, nq, nkv, scale, stream);
}
int attn_fwd_seg_launch(const void* qkv, const int* key_mask, void* out, float* lse, long long rows, int nq, int nkv,
                        float scale, const b200rl_attn_qblock* qblocks_dev, int n_qblocks, cudaStream_t stream);
int attn_bwd_seg_launch(const void* qkv, const int* key_mask, const void* dout, const float* lse, const float* delta,
                        void* dqkv, float* kv_part, long long rows, int nq, int nkv, float scale,
                        const b200rl_attn_qblock* qblocks_dev, int n_qblocks, const b200rl_attn_kblock* kblocks_dev,
                        int n_kblocks, const int* red_start_dev, const int* red_list_dev, cudaStream_t stream);
static int g_attn_tc = -1;
static bool attn_tc_enabled() {
  if (g_attn_tc < 0) {
    const char* e = getenv("B200RL_ATTN_TC");
    g_attn_tc = (e && e[0] == '0') ? 0 : 1;
  }
  return g_attn_tc != 0;
}

}  // namespace b200rl

using namespace b200rl;

// bisection switch: 1 (default) = tcgen05 attention kernels where available (head_dim 128), 0 = mma.sync kernels
extern "C" int b200rl_attn_set_tc(int enable) {
  b200rl::g_attn_tc = enable ? 1 : 0;
  return 0;
}

extern "C" int b200rl_attn_fwd(const void* qkv, const int* key_mask, void* out, float* lse, int B,
                               int L, int n_q_heads, int n_kv_heads, int head_dim, float scale,
                               void* stream) {
  B200RL_REQUIRE(qkv && key_mask && out && lse && B > 0 && L > 0, "attn_fwd: bad args");
  B200RL_REQUIRE(n_kv_heads > 0 && n_q_heads % n_kv_heads == 0, "attn_fwd: nq %% nkv != 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 32: return attn_fwd_launch<32>(qkv, key_mask, out, lse, B, L, n_q_heads, n_kv_heads, scale, st);
    case 64: return attn_fwd_launch<64>(qkv, key_mask, out, lse, B, L, n_q_heads, n_kv_heads, scale, st);
    case 128:
      if (attn_tc_enabled()) return attn_fwd_tc_launch(qkv, key_mask, out, lse, B, L, n_q_heads, n_kv_heads, scale, st);
      return attn_fwd_launch<128>(qkv, key_mask, out, lse, B, L, n_q_heads, n_kv_heads, scale, st);
    default: return set_error(B200RL_ERR_UNSUPPORTED, "attn: head_dim %d not in {32,64,128}", head_dim);
  }
}

extern "C" int b200rl_attn_bwd(const void* qkv, const int* key_mask, const void* out, const void* dout,
                               const float* lse, float* delta, void* dqkv, int B, int L,
                               int n_q_heads, int n_kv_heads, int head_dim, float scale, void* stream) {
  B200RL_REQUIRE(qkv && key_mask && out && dout && lse && delta && dqkv && B > 0 && L > 0,
                 "attn_bwd: bad args");
  B200RL_REQUIRE(n_kv_heads > 0 && n_q_heads % n_kv_heads == 0, "attn_bwd: nq %% nkv != 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 32: return attn_bwd_launch<32>(qkv, key_mask, out, dout, lse, delta, dqkv, B, L, n_q_heads, n_kv_heads, scale, st);
    case 64: return attn_bwd_launch<64>(qkv, key_mask, out, dout, lse, delta, dqkv, B, L, n_q_heads, n_kv_heads, scale, st);
    case 128:
      if (attn_tc_enabled()) return attn_bwd_tc(qkv, key_mask, out, dout, lse, delta, dqkv, B, L, n_q_heads, n_kv_heads, scale, st);
      return attn_bwd_launch<128>(qkv, key_mask, out, dout, lse, delta, dqkv, B, L, n_q_heads, n_kv_heads, scale, st);
    default: return set_error(B200RL_ERR_UNSUPPORTED, "attn: head_dim %d not in {32,64,128}", head_dim);
  }
}

// ---- packed (shared-prompt) layout, head_dim 128 only ----------------------------------------------------------
extern "C" int b200rl_attn_seg_fwd(const void* qkv, const int* key_mask, void* out, float* lse, long long rows,
                                   int n_q_heads, int n_kv_heads, float scale,
                                   const b200rl_attn_qblock* qblocks_dev, int n_qblocks, void* stream) {
  B200RL_REQUIRE(qkv && key_mask && out && lse && qblocks_dev && rows > 0 && n_qblocks > 0, "attn_seg_fwd: bad args");
  B200RL_REQUIRE(n_kv_heads > 0 && n_q_heads % n_kv_heads == 0, "attn_seg_fwd: nq %% nkv != 0");
  return attn_fwd_seg_launch(qkv, key_mask, out, lse, rows, n_q_heads, n_kv_heads, scale, qblocks_dev, n_qblocks,
                             reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int b200rl_attn_seg_bwd(const void* qkv, const int* key_mask, const void* out, const void* dout,
                                   const float* lse, float* delta, void* dqkv, float* kv_part, long long rows,
                                   int n_q_heads, int n_kv_heads, float scale,
                                   const b200rl_attn_qblock* qblocks_dev, int n_qblocks,
                                   const b200rl_attn_kblock* kblocks_dev, int n_kblocks, const int* red_start_dev,
                                   const int* red_list_dev, void* stream) {
  B200RL_REQUIRE(qkv && key_mask && out && dout && lse && delta && dqkv && kv_part && qblocks_dev && kblocks_dev &&
                     red_start_dev && red_list_dev && rows > 0, "attn_seg_bwd: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // delta[h][row]: the classic kernel with B = 1, L = rows uses exactly that index
  const long long warps = rows * n_q_heads;
  B200RL_CUDA_OK(launch_pdl(attn_delta_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, (const bf16*)out, (const bf16*)dout, delta, (int)rows,
                                                                  n_q_heads, 128, rows));
  B200RL_LAUNCH_OK();
  return attn_bwd_seg_launch(qkv, key_mask, dout, lse, delta, dqkv, kv_part, rows, n_q_heads, n_kv_heads, scale,
                             qblocks_dev, n_qblocks, kblocks_dev, n_kblocks, red_start_dev, red_list_dev, st);
}
