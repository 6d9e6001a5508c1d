// G8: multi-learner gradient average as a one-shot P2P reduce over NVSwitch, fused with the
// Adam/AdamW update of the flat fp32 LoRA parameter buffer, then P2P write-back of the updated
// slice to every learner.
//
// Replaces (reference): `param.grad.clone().cpu()` export (distributed_actor.py:289-293), the
// driver hop (distributed_trainer.py:325-342), the CPU python-loop mean and H2D re-assignment
// (distributed_actor.py:311-328) and `bnb.optim.Adam8bit.step()` (:209-211, :332, :414, :512).
// Reference semantics kept: merged grad = mean over learners (:323); Adam betas (0.9, 0.999),
// eps 1e-8, weight_decay 0 ([3P] bitsandbytes defaults) — the 8-bit state quantisation of
// Adam8bit is NOT restated (fp32 m/v): post-step weights are "parity unpinned" and compared with
// torch.optim.Adam instead.  Unlike the reference (quirk Q4: only learner 0 steps), every learner
// ends the step with identical parameters.
//
// Update formulas follow torch.optim.Adam's single-tensor path operation by operation
// (lerp for m, mul+addcmul for v, sqrt/bias_correction2_sqrt + eps, addcdiv) so the fp32 result
// matches torch to ~1 ulp.
#include "common.cuh"
#include <string.h>
#include <math.h>

namespace b200rl {

static constexpr int MAX_PEERS = 8;

struct AdamArgs {
  float* p;        // local params (fp32)
  float* m;
  float* v;
  const float* g[MAX_PEERS];  // gradient buffers of every learner (peer-mapped), rank order
  float* p_peer[MAX_PEERS];   // parameter buffers of every learner (peer-mapped), rank order
  int world;                  // number of learners (1 = local only)
  long long lo, hi;           // owned slice [lo, hi), multiples of 4
  float inv_world;
  float beta1_w;              // 1 - beta1 (lerp weight)
  float beta2, one_minus_beta2;
  float inv_bc2_sqrt;         // 1 / sqrt(1 - beta2^t)
  float eps;
  float neg_step;             // -(lr / (1 - beta1^t))
  float decay;                // 1 - lr * weight_decay (AdamW), 1.0 for Adam
  int zero_local_grad;        // world == 1: clear the gradient in the same pass
  const volatile int* status; // world > 1: the group's status word; non-zero (a barrier timed out) = do nothing
};

// Device-resident copy of "a barrier of this process gave up" (the host-mapped status word), refreshed by every barrier
// kernel.  The reduce kernel used to read the host-mapped word itself, once per block: 1184 PCIe reads made a 424 MB
// reduce take 935 us (453 GB/s; profiles/r2_run29_reduce_adam_alone.txt) — now it reads this word from L2.
__device__ int g_p2p_gave_up = 0;

// WORLD > 0: compile-time learner count, so that all WORLD peer loads of an element are in flight together (a remote
// load is ~2 us of NVLink latency; issued one after the other they serialise).  WORLD == 0: generic loop.
template <int WORLD>
__global__ void __launch_bounds__(256) reduce_adam_kernel(const AdamArgs a) {
  const int world = WORLD > 0 ? WORLD : a.world;
  if (world > 1 && a.status && g_p2p_gave_up != 0) return;   // a peer never arrived at the barrier: leave every buffer untouched
  const long long n4 = (a.hi - a.lo) / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long e = a.lo + i * 4;
    // one-shot reduce: read this slice from every learner's gradient buffer (NVLink loads),
    // summed in rank order so every run gives the same bits
    float4 g;
    if constexpr (WORLD > 0) {
      float4 t[WORLD];
#pragma unroll
      for (int r = 0; r < WORLD; ++r) t[r] = *reinterpret_cast<const float4*>(a.g[r] + e);
      g = t[0];
#pragma unroll
      for (int r = 1; r < WORLD; ++r) { g.x += t[r].x; g.y += t[r].y; g.z += t[r].z; g.w += t[r].w; }
    } else {
      g = *reinterpret_cast<const float4*>(a.g[0] + e);
      for (int r = 1; r < world; ++r) {
        const float4 t = *reinterpret_cast<const float4*>(a.g[r] + e);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    if (world > 1) { g.x *= a.inv_world; g.y *= a.inv_world; g.z *= a.inv_world; g.w *= a.inv_world; }
    float4 p = *reinterpret_cast<float4*>(a.p + e);
    float4 m = *reinterpret_cast<float4*>(a.m + e);
    float4 v = *reinterpret_cast<float4*>(a.v + e);
    float* gp = &g.x; float* pp = &p.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pp[j] = pp[j] * a.decay;
      mp[j] = __fadd_rn(mp[j], __fmul_rn(a.beta1_w, __fsub_rn(gp[j], mp[j])));           // lerp
      vp[j] = __fadd_rn(__fmul_rn(vp[j], a.beta2),
                        __fmul_rn(__fmul_rn(a.one_minus_beta2, gp[j]), gp[j]));           // addcmul
      const float denom = __fadd_rn(__fmul_rn(sqrtf(vp[j]), a.inv_bc2_sqrt), a.eps);
      pp[j] = __fadd_rn(pp[j], __fmul_rn(a.neg_step, __fdiv_rn(mp[j], denom)));           // addcdiv
    }
    *reinterpret_cast<float4*>(a.m + e) = m;
    *reinterpret_cast<float4*>(a.v + e) = v;
    if (world > 1) {
      // all-gather by store: push the updated slice into every learner's parameter buffer
#pragma unroll
      for (int r = 0; r < (WORLD > 0 ? WORLD : MAX_PEERS); ++r)
        if (r < world) *reinterpret_cast<float4*>(a.p_peer[r] + e) = p;
    } else {
      *reinterpret_cast<float4*>(a.p + e) = p;
      if (a.zero_local_grad)
        *reinterpret_cast<float4*>(const_cast<float*>(a.g[0]) + e) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// Cross-GPU barrier on peer-mapped flag arrays. flags_peer[r] points at learner r's flag array
// (uint32[MAX_PEERS]); learner `rank` writes `epoch` into slot `rank` of every peer, then waits
// until all slots of its own array reached `epoch`.  A peer that does not arrive within timeout_ns does NOT take the
// CUDA context down: the waiter records (1 + the missing rank) in the group's status word (host-visible) and returns;
// the following reduce kernel sees the word and leaves every buffer untouched, and the host turns the word into an
// error at its next check (b200rl_p2p_status).
struct BarrierArgs {
  unsigned int* flags_peer[MAX_PEERS];
  int world, rank;
  unsigned int epoch;
  unsigned long long timeout_ns;
  volatile int* status;
};
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void p2p_barrier_kernel(const BarrierArgs a) {
  const int t = threadIdx.x;
  if (t < a.world) {
    __threadfence_system();
    volatile unsigned int* dst = a.flags_peer[t] + a.rank;
    *dst = a.epoch;
    __threadfence_system();
    volatile unsigned int* mine = a.flags_peer[a.rank] + t;
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while ((int)(*mine - a.epoch) < 0) {
      if ((++spins & 1023u) == 0) {
        if (a.status && *a.status != 0) break;   // another lane / an earlier barrier already gave up
        if (global_ns() - t0 > a.timeout_ns) {
          if (a.status) *a.status = 1 + t;
          break;
        }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (t == 0) g_p2p_gave_up = (a.status && *a.status != 0) ? 1 : 0;   // one PCIe read per barrier, for the kernels behind it
}

// fp32 master -> bf16 operand copies in the padded layouts the GEMM consumes.
// One descriptor per LoRA tensor; dst element (i, j) = src element (i, j) (optionally transposed),
// written at dst + (row_off + i') * ld + col_off + j'.
struct PackDesc {
  long long src_off;   // offset into the flat fp32 buffer
  int rows, cols;      // source shape
  long long dst_off;   // element offset into the bf16 arena
  int dst_ld;
  int transpose;       // dst[j][i] = src[i][j]
};
__global__ void lora_pack_kernel(const float* __restrict__ flat, bf16* __restrict__ arena,
                                 const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  const long long n = (long long)d.rows * d.cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / d.cols), j = (int)(idx % d.cols);
    const float val = flat[d.src_off + idx];
    const long long o = d.transpose ? (long long)j * d.dst_ld + i : (long long)i * d.dst_ld + j;
    arena[d.dst_off + o] = __float2bfloat16_rn(val);
  }
}

// Accumulate LoRA gradients from split-K fp32 slabs (output of the dW GEMM) into the flat buffer:
//   flat[dst_off + i*cols + j] += scale * sum_s slab[s][ (row_off+i)*ld + col_off + j ]   (or transposed)
struct UnpackDesc {
  long long dst_off;
  int rows, cols;       // destination tensor shape
  const float* slabs;   // [splits][slab_stride]
  long long slab_stride;
  int splits;
  int ld;               // leading dimension inside a slab
  int row_off, col_off; // where the block sits inside the slab
  int transpose;        // dst[i][j] = slab[(row_off+j)*ld + col_off + i]
};
__global__ void lora_grad_accum_kernel(float* __restrict__ flat, const UnpackDesc* __restrict__ descs) {
  const UnpackDesc d = descs[blockIdx.y];
  const long long n = (long long)d.rows * d.cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / d.cols), j = (int)(idx % d.cols);
    const long long o = d.transpose ? (long long)(d.row_off + j) * d.ld + d.col_off + i
                                    : (long long)(d.row_off + i) * d.ld + d.col_off + j;
    float acc = 0.f;
    for (int s = 0; s < d.splits; ++s) acc += d.slabs[(long long)s * d.slab_stride + o];
    flat[d.dst_off + idx] += acc;
  }
}

}  // namespace b200rl

using namespace b200rl;
#define STREAM reinterpret_cast<cudaStream_t>(stream)

// Status word of the multi-learner exchange: one per process, in mapped pinned host memory so that a timed-out barrier
// kernel can report without killing the context.  0 = fine, k > 0 = learner k-1 never reached a barrier.
static volatile int* g_p2p_status_host = nullptr;
static int* g_p2p_status_dev = nullptr;
static int p2p_status_init() {
  if (g_p2p_status_host) return 0;
  void* h = nullptr;
  B200RL_CUDA_OK(cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(h, 0, 64);
  void* d = nullptr;
  B200RL_CUDA_OK(cudaHostGetDevicePointer(&d, h, 0));
  g_p2p_status_host = (volatile int*)h;
  g_p2p_status_dev = (int*)d;
  return 0;
}
// With CUDA's lazy module loading the FIRST launch of a kernel loads its code, which synchronises with the device: if a
// barrier kernel of this context is spinning at that moment (several learners in one process, or one stream per "rank" in
// the single-device tests) the load waits for the barrier and the barrier waits for the kernel behind the load.  Force the
// kernels of the exchange to be resident before the first barrier is ever launched.
static int p2p_preload_kernels() {
  static bool done[64] = {false};
  int dev = 0;
  B200RL_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
  cudaFuncAttributes fa;
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, p2p_barrier_kernel));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, reduce_adam_kernel<0>));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, reduce_adam_kernel<1>));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, reduce_adam_kernel<2>));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, reduce_adam_kernel<4>));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, reduce_adam_kernel<8>));
  B200RL_CUDA_OK(cudaFuncGetAttributes(&fa, lora_pack_kernel));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
}
static double p2p_timeout_s() {
  static double t = -1.0;
  if (t < 0) {
    const char* e = getenv("B200RL_P2P_TIMEOUT_S");
    t = e ? atof(e) : 600.0;   // long enough for a peer that is still loading weights or running a long step
    if (!(t > 0)) t = 600.0;
  }
  return t;
}

// Single- or multi-learner fused (reduce +) Adam(W). `grads` / `params_peer` are arrays of `world`
// device pointers (host memory) in rank order; for world == 1 pass the local buffers.
extern "C" int b200rl_lora_reduce_adamw(float* p, float* m, float* v, const float* const* grads,
                                        float* const* params_peer, int world, int rank,
                                        long long n, int step, float lr, float beta1, float beta2,
                                        float eps, float weight_decay, int zero_local_grad,
                                        void* stream) {
  B200RL_REQUIRE(p && m && v && grads && n > 0 && n % 4 == 0, "reduce_adamw: bad args (n=%lld)", n);
  B200RL_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world,
                 "reduce_adamw: world=%d rank=%d", world, rank);
  B200RL_REQUIRE(step >= 1, "reduce_adamw: step counts from 1");
  AdamArgs a;
  a.p = p; a.m = m; a.v = v;
  for (int r = 0; r < MAX_PEERS; ++r) {
    a.g[r] = r < world ? grads[r] : nullptr;
    a.p_peer[r] = (r < world && params_peer) ? params_peer[r] : nullptr;
  }
  if (world > 1) B200RL_REQUIRE(params_peer != nullptr, "reduce_adamw: world > 1 needs params_peer");
  a.status = nullptr;
  if (world > 1) {
    int rc = p2p_status_init();
    if (rc) return rc;
    a.status = g_p2p_status_dev;
  }
  a.world = world;
  // owned slice: contiguous, 4-element aligned
  const long long n4 = n / 4;
  const long long per = (n4 + world - 1) / world;
  a.lo = 4 * (per * rank < n4 ? per * rank : n4);
  a.hi = 4 * (per * (rank + 1) < n4 ? per * (rank + 1) : n4);
  a.inv_world = 1.0f / (float)world;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.beta1_w = (float)(1.0 - (double)beta1);
  a.beta2 = beta2;
  a.one_minus_beta2 = (float)(1.0 - (double)beta2);
  a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  a.eps = eps;
  a.neg_step = (float)(-((double)lr / bc1));
  a.decay = (float)(1.0 - (double)lr * (double)weight_decay);
  a.zero_local_grad = zero_local_grad;
  if (a.hi > a.lo) {
    const long long work = (a.hi - a.lo) / 4;
    long long blocks = (work + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
    switch (world) {
      case 1: reduce_adam_kernel<1><<<g, 256, 0, STREAM>>>(a); break;
      case 2: reduce_adam_kernel<2><<<g, 256, 0, STREAM>>>(a); break;
      case 4: reduce_adam_kernel<4><<<g, 256, 0, STREAM>>>(a); break;
      case 8: reduce_adam_kernel<8><<<g, 256, 0, STREAM>>>(a); break;
      default: reduce_adam_kernel<0><<<g, 256, 0, STREAM>>>(a); break;
    }
    B200RL_LAUNCH_OK();
  }
  return 0;
}

extern "C" int b200rl_p2p_barrier_timeout(unsigned int* const* flags_peer, int world, int rank,
                                          unsigned int epoch, double timeout_s, void* stream) {
  B200RL_REQUIRE(flags_peer && world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world,
                 "p2p_barrier: bad args");
  int rc = p2p_status_init();
  if (!rc) rc = p2p_preload_kernels();
  if (rc) return rc;
  BarrierArgs a;
  for (int r = 0; r < MAX_PEERS; ++r) a.flags_peer[r] = r < world ? flags_peer[r] : nullptr;
  a.world = world; a.rank = rank; a.epoch = epoch;
  if (!(timeout_s > 0)) timeout_s = p2p_timeout_s();
  a.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  a.status = g_p2p_status_dev;
  p2p_barrier_kernel<<<1, 32, 0, STREAM>>>(a);
  B200RL_LAUNCH_OK();
  return 0;
}
extern "C" int b200rl_p2p_barrier(unsigned int* const* flags_peer, int world, int rank,
                                  unsigned int epoch, void* stream) {
  return b200rl_p2p_barrier_timeout(flags_peer, world, rank, epoch, 0.0, stream);
}
// 0 = every barrier so far completed; k > 0 = learner k-1 did not arrive in time (an error is recorded and returned as
// B200RL_ERR_STATE).  Reads a host-mapped word: call after a stream / event synchronisation.  reset != 0 clears it.
extern "C" int b200rl_p2p_status(int reset) {
  if (!g_p2p_status_host) return 0;
  const int s = *g_p2p_status_host;
  if (reset) *g_p2p_status_host = 0;
  if (s != 0) return set_error(B200RL_ERR_STATE, "p2p barrier timed out waiting for learner %d", s - 1);
  return 0;
}

// ---- CUDA IPC plumbing for the peer-mapped buffers (library-owned allocations) -------------------
extern "C" int b200rl_p2p_alloc(long long bytes, void** ptr, void* handle64) {
  B200RL_REQUIRE(bytes > 0 && ptr && handle64, "p2p_alloc: bad args");
  {
    int rc = p2p_status_init();
    if (!rc) rc = p2p_preload_kernels();
    if (rc) return rc;
  }
  B200RL_CUDA_OK(cudaMalloc(ptr, (size_t)bytes));
  B200RL_CUDA_OK(cudaMemset(*ptr, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  B200RL_CUDA_OK(cudaIpcGetMemHandle(&h, *ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int b200rl_p2p_open(const void* handle64, void** ptr) {
  B200RL_REQUIRE(handle64 && ptr, "p2p_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200RL_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int b200rl_p2p_close(void* ptr) {
  B200RL_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int b200rl_p2p_free(void* ptr) {
  B200RL_CUDA_OK(cudaFree(ptr));
  return 0;
}
// Learners that live in ONE process on different devices (the single-process trainer harness) exchange plain device
// pointers instead of IPC handles; the current device then needs explicit peer access to `peer_device`.
extern "C" int b200rl_p2p_enable_peer_access(int peer_device) {
  int dev = 0;
  B200RL_CUDA_OK(cudaGetDevice(&dev));
  if (dev == peer_device) return 0;
  int can = 0;
  B200RL_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  B200RL_REQUIRE(can, "p2p_enable_peer_access: device %d cannot access device %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    (void)cudaGetLastError();
    return 0;
  }
  B200RL_CUDA_OK(e);
  return 0;
}

extern "C" int b200rl_lora_pack(const float* flat, void* arena_bf16, const void* descs_dev,
                                int n_desc, int max_elems, void* stream) {
  B200RL_REQUIRE(flat && arena_bf16 && descs_dev && n_desc > 0, "lora_pack: bad args");
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_desc);
  lora_pack_kernel<<<grid, 256, 0, STREAM>>>(flat, (bf16*)arena_bf16, (const PackDesc*)descs_dev);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_lora_grad_accum(float* flat, const void* descs_dev, int n_desc, int max_elems,
                                      void* stream) {
  B200RL_REQUIRE(flat && descs_dev && n_desc > 0, "lora_grad_accum: bad args");
  int bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_desc);
  lora_grad_accum_kernel<<<grid, 256, 0, STREAM>>>(flat, (const UnpackDesc*)descs_dev);
  B200RL_LAUNCH_OK();
  return 0;
}

extern "C" int b200rl_sizeof_pack_desc(void) { return (int)sizeof(PackDesc); }
extern "C" int b200rl_sizeof_unpack_desc(void) { return (int)sizeof(UnpackDesc); }
