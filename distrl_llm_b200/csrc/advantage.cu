// G9: group-relative advantage normalisation + top-k subselect, fp64, one CTA per problem.
//
// Reference: Trainer.train advantage block (distributed_trainer.py:262-279) and top-k filter
// (:281-294), which run in numpy float64 on the driver:
//   s_c      = rewards[c,0] + rewards[c,1]                         (:273 batch_reward.sum(axis=1))
//   mean     = np.mean(s)        (pairwise summation)              (:267, :273)
//   std      = np.std(s)         (ddof=0: sqrt(mean((s-mean)^2)))  (:273)
//   GRPO adv = (s - mean) / (std + 1e-8)                           (:273, :276)
//   PG       : rewards = s, baseline = mean                        (:267, :274, :278-279)
//   top-k    : idx = argsort(values)[-k:]  (ascending)             (:287)
// np.add.reduce's pairwise summation is restated exactly (block of <=128 with 8 interleaved
// partial sums, recursive halving above) and every fp64 op uses an explicit rounding intrinsic so
// no FMA contraction changes bits: results are bit-identical to numpy for distinct values.
// Ties in argsort: numpy's introsort is unstable; this kernel uses the stable order (equal values
// keep candidate order), the oracle does the same with kind="stable" (documented divergence: only
// the ORDER of equal-advantage candidates can differ, never the selected value multiset).
#include "common.cuh"

namespace b200rl {

__device__ double np_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  } else if (n <= 128) {
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
  }
}

// rewards: [G, C, 2] f64.  values_out: [G, C] (GRPO advantages or PG summed rewards),
// baseline_out: [G] (mean of summed rewards), topk_idx: [G, k] candidate indices in ascending
// value order (the reference's argsort(...)[-k:]), topk_val: [G, k].
__global__ void group_advantage_topk_kernel(const double* __restrict__ rewards,
                                            double* __restrict__ values_out,
                                            double* __restrict__ baseline_out,
                                            int* __restrict__ topk_idx, double* __restrict__ topk_val,
                                            int C, int k, int grpo) {
  extern __shared__ double sh[];  // s[C], tmp[C]
  double* s = sh;
  double* tmp = sh + C;
  __shared__ double s_mean, s_std;
  const int g = blockIdx.x;
  const double* r = rewards + (size_t)g * C * 2;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s[c] = __dadd_rn(r[2 * c], r[2 * c + 1]);
  __syncthreads();
  if (threadIdx.x == 0) s_mean = __ddiv_rn(np_pairwise_sum(s, C), (double)C);
  __syncthreads();
  const double mean = s_mean;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double d = __dsub_rn(s[c], mean);
    tmp[c] = __dmul_rn(d, d);
  }
  __syncthreads();
  if (threadIdx.x == 0) s_std = __dsqrt_rn(__ddiv_rn(np_pairwise_sum(tmp, C), (double)C));
  __syncthreads();
  const double denom = __dadd_rn(s_std, 1e-8);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double val = grpo ? __ddiv_rn(__dsub_rn(s[c], mean), denom) : s[c];
    tmp[c] = val;
    values_out[(size_t)g * C + c] = val;
  }
  if (threadIdx.x == 0 && baseline_out) baseline_out[g] = mean;
  __syncthreads();
  if (!topk_idx) return;
  // stable ascending rank by counting; keep ranks >= C-k
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double v = tmp[c];
    int rank = 0;
    for (int j = 0; j < C; ++j) rank += (tmp[j] < v) || (tmp[j] == v && j < c);
    const int pos = rank - (C - k);
    if (pos >= 0) {
      topk_idx[(size_t)g * k + pos] = c;
      if (topk_val) topk_val[(size_t)g * k + pos] = v;
    }
  }
}

}  // namespace b200rl

using namespace b200rl;

extern "C" int b200rl_group_advantage_topk(const double* rewards, double* values, double* baselines,
                                           int* topk_idx, double* topk_val, int G, int C, int k,
                                           int grpo, void* stream) {
  B200RL_REQUIRE(rewards && values && G > 0 && C > 0, "group_advantage_topk: bad args");
  B200RL_REQUIRE(k >= 1, "group_advantage_topk: k must be >= 1");
  if (k > C) k = C;  // argsort(...)[-k:] with k > C returns all C
  B200RL_REQUIRE(C <= 2048, "group_advantage_topk: C=%d too large", C);
  const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  group_advantage_topk_kernel<<<G, threads, 2 * C * sizeof(double),
                                reinterpret_cast<cudaStream_t>(stream)>>>(
      rewards, values, baselines, topk_idx, topk_val, C, k, grpo);
  B200RL_LAUNCH_OK();
  return 0;
}
