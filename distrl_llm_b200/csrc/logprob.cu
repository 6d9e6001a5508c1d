// G7: token-gather / log-softmax / advantage-scale in ONE coalesced pass over the logits.
//
// Reference: BaseLearner.compute_current_policy_probs per-row log_softmax + gather
// (distributed_actor.py:252-260), PG loss (:375), GRPO loss (:467-470), 1/num_batches scaling
// (:382, :479) and the autograd backward of all of it (:385, :483).
//
// Because the loss is linear in the per-token log-probs with coefficients that depend only on
// (advantage, answer mask, length, batch sizes),
//     d loss / d lp[i,t] = coef[i,t] = -A_i * mask[i,t] / (len_i * B_m * nb),
// forward and backward fuse: pass 1 streams a logits row once (online max / sum-exp, fp32 like the
// reference's autocast log_softmax), pass 2 re-reads the row (L2-resident: 304 KB per row at
// V=152064) and overwrites it in place with dlogits = coef * (onehot(y) - softmax(z)).
// HBM traffic = read T*V*2 + write T*V*2 bytes; the reference saves an fp32 [T,V] log-prob tensor
// per row for autograd instead.
#include "common.cuh"

namespace b200rl {

static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

// one CTA per row
__global__ void __launch_bounds__(1024)
logprob_kernel(bf16* __restrict__ logits, long long ld, const int* __restrict__ targets,
               const float* __restrict__ coef, const float* __restrict__ klw,
               const float* __restrict__ ref_lp, const float* __restrict__ old_lp, float clip_eps,
               float* __restrict__ lp_out, int V, int write_grad, const int* __restrict__ slot) {
  __shared__ float red_m[32], red_s[32];
  __shared__ float s_lse2, s_max2;
  const size_t row = blockIdx.x;
  bf16* z = logits + row * ld;
  const bf16x8* zv = reinterpret_cast<const bf16x8*>(z);
  const int nvec = V / 8;  // V % 8 == 0 checked on host
  // ---- pass 1: online softmax statistics in the log2 domain ----
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8];
    unpack8(zv[i], f);
    float lm = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
    lm *= LOG2E;
    if (lm > m) {
      s *= exp2f(m - lm);
      m = lm;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += exp2f(f[j] * LOG2E - m);
  }
  // warp then block combine of (m, s)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (nm == -INFINITY) ? 0.f : s * exp2f(m - nm) + os * exp2f(om - nm);
    m = nm;
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    red_m[w] = m;
    red_s[w] = s;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    m = l < nw ? red_m[l] : -INFINITY;
    s = l < nw ? red_s[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const float os = __shfl_xor_sync(0xffffffffu, s, o);
      const float nm = fmaxf(m, om);
      s = (nm == -INFINITY) ? 0.f : s * exp2f(m - nm) + os * exp2f(om - nm);
      m = nm;
    }
    if (l == 0) {
      s_max2 = m;
      s_lse2 = m + log2f(s);  // log2-sum-exp2
    }
  }
  __syncthreads();
  const float lse2 = s_lse2;
  // compacted scoring (packed layout, live rows only): logits / targets are per compacted row, the per-token arrays
  // (coef, klw, ref_lp, old_lp, lp_out) stay [B*T] and are addressed through the row's slot
  const size_t sl = slot ? (size_t)slot[row] : row;
  const int y = targets[row];
  const float zy = (y >= 0 && y < V) ? __bfloat162float(z[y]) : 0.f;
  const float lp = (y >= 0 && y < V) ? (zy * LOG2E - lse2) * LN2 : 0.f;
  float c = coef ? coef[sl] : 0.f;
  // optional clipped-ratio surrogate (not in the reference, whose ratio exp(lp - lp.detach()) is identically 1,
  // distributed_actor.py:467): loss_t = -min(rho A, clip(rho, 1-eps, 1+eps) A) with rho = exp(lp - old_lp), so
  // d loss_t / d lp = -A rho while the unclipped branch is the active minimum and 0 once rho has left the trust region in
  // the direction the advantage pushes.  coef carries -A mask / (len Bm nb): its sign is that of -A.
  if (old_lp && clip_eps > 0.f && c != 0.f) {
    const float rho = __expf(lp - old_lp[sl]);
    const bool active = c < 0.f ? rho <= 1.f + clip_eps : rho >= 1.f - clip_eps;
    c = active ? c * rho : 0.f;
  }
  // optional KL(pi || pi_ref) term, k3 estimator exp(q-p) - (q-p) - 1 per token:
  // d k3 / d lp = 1 - exp(q - p); klw carries beta * mask / (len * Bm * nb)   (not in the reference: beta = 0)
  if (klw && ref_lp && klw[sl] != 0.f) c += klw[sl] * (1.f - __expf(ref_lp[sl] - lp));
  if (threadIdx.x == 0 && lp_out) lp_out[sl] = lp;
  if (!write_grad) return;
  __syncthreads();  // z[y] read above must precede the in-place overwrite
  // ---- pass 2: dz = coef * (onehot - softmax), in place ----
  bf16x8* zw = reinterpret_cast<bf16x8*>(z);
  if (c == 0.f) {
    bf16x8 zero;
    zero.u = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) zw[i] = zero;
    return;
  }
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float f[8];
    unpack8(zv[i], f);
    const int base = i * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p = exp2f(f[j] * LOG2E - lse2);
      f[j] = c * (((base + j) == y ? 1.f : 0.f) - p);
    }
    zw[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------
// loss coefficients (reference :375 / :470 with :382 / :479 scaling), fp64 like the reference's
// float64 rewards tensor (:350, :441), emitted as fp32 per scored token:
//   len_i = sum_t mask[i,t];  coef[i,t] = -A_i * mask[i,t] / (len_i * Bm * nb)   (0 when len_i = 0)
// ------------------------------------------------------------------------------------------
__global__ void loss_coef_kernel(const int* __restrict__ mask, const double* __restrict__ adv,
                                 float* __restrict__ coef, float* __restrict__ klw, double beta,
                                 int* __restrict__ lens, int T, int Bm, int nb) {
  __shared__ int red[32];
  const int i = blockIdx.x;
  int cnt = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) cnt += mask[(size_t)i * T + t] != 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    int t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  const int len = red[0];
  if (threadIdx.x == 0 && lens) lens[i] = len;
  const double base = len > 0 ? -adv[i] / ((double)len * (double)Bm * (double)nb) : 0.0;
  const double kbase = len > 0 ? beta / ((double)len * (double)Bm * (double)nb) : 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const bool on = mask[(size_t)i * T + t] != 0;
    coef[(size_t)i * T + t] = on ? (float)base : 0.f;
    if (klw) klw[(size_t)i * T + t] = on ? (float)kbase : 0.f;
  }
}

// loss value of one micro-batch, accumulated on the device (no per-micro-batch .item() sync):
//   PG   (:375): loss_m = -(1/Bm) sum_i A_i * (sum_t mask*lp) / len_i
//   GRPO (:467-470): importance = exp(lp - lp.detach()) == 1  =>  loss_m = -(1/Bm) sum_i A_i * [len_i>0]
// The reference's returned scalar is the SUM of loss_m over micro-batches (quirk Q2) -> *accum += loss_m.
__global__ void loss_value_kernel(const float* __restrict__ lp, const int* __restrict__ mask,
                                  const double* __restrict__ adv, const float* __restrict__ ref_lp,
                                  double beta, const float* __restrict__ old_lp, double clip_eps,
                                  double* __restrict__ accum, int Bm, int T, int grpo) {
  __shared__ double red[32];
  double total = 0.0;
  for (int i = 0; i < Bm; ++i) {
    double s = 0.0, kl = 0.0;
    int cnt = 0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      if (mask[(size_t)i * T + t] != 0) {
        const double p = (double)lp[(size_t)i * T + t];
        if (old_lp && clip_eps > 0.0) {   // clipped surrogate: s accumulates min(rho A, clip(rho) A) instead of lp
          const double rho = exp(p - (double)old_lp[(size_t)i * T + t]);
          const double a = adv[i];
          const double cl = fmin(fmax(rho, 1.0 - clip_eps), 1.0 + clip_eps);
          s += fmin(rho * a, cl * a);
        } else {
          s += p;
        }
        cnt += 1;
        if (ref_lp) {
          const double d = (double)ref_lp[(size_t)i * T + t] - p;
          kl += exp(d) - d - 1.0;
        }
      }
    }
    s = warp_sum_d(s);
    kl = warp_sum_d(kl);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    __shared__ double redk[32];
    if ((threadIdx.x & 31) == 0) {
      red[threadIdx.x >> 5] = s;
      redk[threadIdx.x >> 5] = kl;
    }
    __shared__ int redc[32];
    if ((threadIdx.x & 31) == 0) redc[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      double ss = 0.0, kk = 0.0;
      int cc = 0;
      for (int w = 0; w < (blockDim.x >> 5); ++w) {
        ss += red[w];
        kk += redk[w];
        cc += redc[w];
      }
      const bool clipped = old_lp && clip_eps > 0.0;
      if (cc > 0) total += (clipped ? ss / (double)cc : (grpo ? adv[i] : adv[i] * (ss / (double)cc))) - beta * (kk / (double)cc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *accum += -total / (double)Bm;
}

}  // namespace b200rl

using namespace b200rl;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

extern "C" int b200rl_logprob_kl(void* logits, long long ld, const int* targets, const float* coef,
                                 const float* klw, const float* ref_lp, float* lp_out, int rows, int V,
                                 int write_grad, void* stream) {
  B200RL_REQUIRE(logits && targets && rows > 0 && V > 0 && V % 8 == 0 && ld % 8 == 0,
                 "logprob: bad args (rows=%d V=%d ld=%lld)", rows, V, ld);
  B200RL_REQUIRE(!write_grad || coef, "logprob: write_grad needs coef");
  const int threads = V >= 8192 ? 1024 : 256;
  logprob_kernel<<<rows, threads, 0, STREAM>>>((bf16*)logits, ld, targets, coef, klw, ref_lp, nullptr, 0.f, lp_out, V,
                                               write_grad, nullptr);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_logprob(void* logits, long long ld, const int* targets, const float* coef,
                              float* lp_out, int rows, int V, int write_grad, void* stream) {
  B200RL_REQUIRE(logits && targets && rows > 0 && V > 0 && V % 8 == 0 && ld % 8 == 0,
                 "logprob: bad args (rows=%d V=%d ld=%lld)", rows, V, ld);
  B200RL_REQUIRE(!write_grad || coef, "logprob: write_grad needs coef");
  const int threads = V >= 8192 ? 1024 : 256;
  logprob_kernel<<<rows, threads, 0, STREAM>>>((bf16*)logits, ld, targets, coef, nullptr, nullptr, nullptr, 0.f, lp_out,
                                               V, write_grad, nullptr);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_loss_coef_kl(const int* mask, const double* adv, float* coef, float* klw, double beta,
                                   int* lens, int Bm, int T, int nb, void* stream) {
  B200RL_REQUIRE(mask && adv && coef && Bm > 0 && T > 0 && nb > 0, "loss_coef: bad args");
  loss_coef_kernel<<<Bm, 256, 0, STREAM>>>(mask, adv, coef, klw, beta, lens, T, Bm, nb);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_loss_coef(const int* mask, const double* adv, float* coef, int* lens, int Bm,
                                int T, int nb, void* stream) {
  B200RL_REQUIRE(mask && adv && coef && Bm > 0 && T > 0 && nb > 0, "loss_coef: bad args");
  loss_coef_kernel<<<Bm, 256, 0, STREAM>>>(mask, adv, coef, nullptr, 0.0, lens, T, Bm, nb);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_loss_value_kl(const float* lp, const int* mask, const double* adv, const float* ref_lp,
                                    double beta, double* accum, int Bm, int T, int grpo, void* stream) {
  B200RL_REQUIRE(lp && mask && adv && accum && Bm > 0 && T > 0, "loss_value: bad args");
  loss_value_kernel<<<1, 256, 0, STREAM>>>(lp, mask, adv, ref_lp, beta, nullptr, 0.0, accum, Bm, T, grpo);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_loss_value(const float* lp, const int* mask, const double* adv, double* accum,
                                 int Bm, int T, int grpo, void* stream) {
  B200RL_REQUIRE(lp && mask && adv && accum && Bm > 0 && T > 0, "loss_value: bad args");
  loss_value_kernel<<<1, 256, 0, STREAM>>>(lp, mask, adv, nullptr, 0.0, nullptr, 0.0, accum, Bm, T, grpo);
  B200RL_LAUNCH_OK();
  return 0;
}

// Clipped-ratio variants (not in the reference): old_lp [rows] = log-probs of the policy that
// generated the batch, clip_eps = trust-region half width.  old_lp == NULL or clip_eps == 0 is the plain form above.
extern "C" int b200rl_logprob_slots(void* logits, long long ld, const int* targets, const float* coef, const float* klw,
                                    const float* ref_lp, const float* old_lp, double clip_eps, float* lp_out, int rows,
                                    int V, int write_grad, const int* slot, void* stream) {
  B200RL_REQUIRE(logits && targets && rows > 0 && V > 0 && V % 8 == 0 && ld % 8 == 0,
                 "logprob: bad args (rows=%d V=%d ld=%lld)", rows, V, ld);
  B200RL_REQUIRE(!write_grad || coef, "logprob: write_grad needs coef");
  B200RL_REQUIRE(clip_eps >= 0.0, "logprob: clip_eps must be >= 0");
  const int threads = V >= 8192 ? 1024 : 256;
  logprob_kernel<<<rows, threads, 0, STREAM>>>((bf16*)logits, ld, targets, coef, klw, ref_lp, old_lp, (float)clip_eps,
                                               lp_out, V, write_grad, slot);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_logprob_clip(void* logits, long long ld, const int* targets, const float* coef, const float* klw,
                                   const float* ref_lp, const float* old_lp, double clip_eps, float* lp_out, int rows,
                                   int V, int write_grad, void* stream) {
  return b200rl_logprob_slots(logits, ld, targets, coef, klw, ref_lp, old_lp, clip_eps, lp_out, rows, V, write_grad,
                              nullptr, stream);
}

extern "C" int b200rl_loss_value_clip(const float* lp, const int* mask, const double* adv, const float* ref_lp,
                                      double beta, const float* old_lp, double clip_eps, double* accum, int Bm, int T,
                                      int grpo, void* stream) {
  B200RL_REQUIRE(lp && mask && adv && accum && Bm > 0 && T > 0, "loss_value: bad args");
  loss_value_kernel<<<1, 256, 0, STREAM>>>(lp, mask, adv, ref_lp, beta, old_lp, clip_eps, accum, Bm, T, grpo);
  B200RL_LAUNCH_OK();
  return 0;
}
