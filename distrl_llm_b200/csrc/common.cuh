// Shared helpers for libb200rl (sm_100a only): error convention, PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM, small math utilities.
//
// Error convention (SURVEY.md §8b "C-ABI the replacement must export"): every entry point
// returns 0 on success or a negative code; b200rl_last_error() returns a thread-local message.
#pragma once
#include <utility>
#include <string.h>
#include <cuda.h>          // CUtensorMap types only; the driver symbol is resolved at run time
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace b200rl {

enum : int {
  B200RL_OK = 0,
  B200RL_ERR_ARG = -1,       // bad argument (shape / alignment / null pointer)
  B200RL_ERR_CUDA = -2,      // a CUDA runtime / driver call failed
  B200RL_ERR_UNSUPPORTED = -3,
  B200RL_ERR_STATE = -4,
};

int set_error(int code, const char* fmt, ...);

#define B200RL_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::b200rl::set_error(::b200rl::B200RL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
  } while (0)

#define B200RL_REQUIRE(cond, ...)                                                         \
  do {                                                                                    \
    if (!(cond)) return ::b200rl::set_error(::b200rl::B200RL_ERR_ARG, __VA_ARGS__);       \
  } while (0)

#define B200RL_LAUNCH_OK()                                                                \
  do {                                                                                    \
    ++::b200rl::g_launch_count;                                                           \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::b200rl::set_error(::b200rl::B200RL_ERR_CUDA, "kernel launch failed: %s (%s:%d)", \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
  } while (0)

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device
// Kernel attributes (max dynamic shared memory ...) belong to a (function, DEVICE) pair: a process that drives several
// devices (one learner thread per GPU) has to set them once per device, not once per process.  Usage:
//   static DeviceOnce once;  if (once.first()) cudaFuncSetAttribute(...);
struct DeviceOnce {
  bool done[64] = {false};
  bool first() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
  }
};
extern long long g_launch_count;  // kernels launched by this library (bench.py's gpu_launches)
extern int g_pdl_enabled;         // programmatic dependent launch on the hot-path kernels (b200rl_set_pdl / B200RL_PDL)

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 8 x bf16 <-> float[8] through ONE 16-byte transaction.  The payload is a uint4 so that copying a bf16x8
// (load from / store to global or shared memory) is a single 128-bit LDG/STG/LDS/STS; a struct of four
// __nv_bfloat162 is copied member-wise and compiles to four 32-bit accesses (seen in SASS / ncu source view).
struct __align__(16) bf16x8 {
  uint4 u;
  __device__ __forceinline__ uint32_t& w(int i) { return reinterpret_cast<uint32_t*>(&u)[i]; }
  __device__ __forceinline__ const uint32_t& w(int i) const { return reinterpret_cast<const uint32_t*>(&u)[i]; }
  __device__ __forceinline__ void set(int i, __nv_bfloat162 h) { w(i) = *reinterpret_cast<uint32_t*>(&h); }
  __device__ __forceinline__ __nv_bfloat162 get(int i) const {
    const uint32_t t = w(i);
    return *reinterpret_cast<const __nv_bfloat162*>(&t);
  }
};
__device__ __forceinline__ void unpack8(const bf16x8& in, float* f) {
  const uint32_t r[4] = {in.u.x, in.u.y, in.u.z, in.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // bf16 -> fp32 is a 16-bit shift
    f[2 * i] = __uint_as_float(r[i] << 16);
    f[2 * i + 1] = __uint_as_float(r[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 o;
  o.u = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                   pack_bf16x2(f[6], f[7]));
  return o;
}

// single-instruction 2^x (MUFU.EX2, rel. error ~2^-22, flushes denormals); exp2f() adds ~4 range-handling instrs
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// one lane of the (converged) warp, chosen by the hardware: the compiler knows the region below it runs on a single
// elected lane, which keeps the tcgen05.mma operands in uniform registers (with `lane == 0` it re-elects and re-broadcasts
// them around every MMA)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must become a trap (reported as a CUDA error by the next
// runtime call), never a hang of the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s at 2 GHz
      printf("b200rl: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// ---- TMA -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit: arrive on an mbarrier when all previously issued MMAs have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A (M = 128 rows x K = 16 bf16) sits in TMEM lane = row, two consecutive k per
// 32-bit column (low half = even k), 8 columns per K = 16 step (scripts/micro/umma_ts_numerics.cu checks the layout).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 (or 16) columns of 32 bits: thread i of the warp writes row (lane base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

#endif  // __CUDACC__

// ---- SwiGLU element math, shared by the row kernels (elementwise.cu) and the fused GEMM epilogues (gemm2_tcgen05.cu)
// so that both paths are bit-identical.  a = gate, b = up (already bf16-rounded values), HF Qwen2MLP:
// act_fn(gate) is rounded to bf16 before the product.
__device__ __forceinline__ void swiglu_fwd8(const float* a, const float* b, float* o) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float s = a[j] / (1.f + __expf(-a[j]));
    o[j] = __bfloat162float(__float2bfloat16_rn(s)) * b[j];
  }
}
// og = dact * up * silu'(gate), ou = dact * silu(gate)
__device__ __forceinline__ void swiglu_bwd8(const float* a, const float* b, const float* c, float* og, float* ou) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sg = 1.f / (1.f + __expf(-a[j]));
    const float silu = a[j] * sg;
    og[j] = c[j] * b[j] * sg * (1.f + a[j] * (1.f - sg));
    ou[j] = c[j] * silu;
  }
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The learner step is ~12k dependent launches on one stream.  Kernels launched through launch_pdl() may start while
// the previous kernel drains: they run their prologue (barrier init, TMEM allocation, tensor-map prefetch) and then
// block in pdl_enter() until the previous grid has completed and its writes are visible.  Rules for a kernel:
// nothing before pdl_enter() may touch global memory that an earlier kernel of the stream writes or reads.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the NEXT kernel may start its prologue
  asm volatile("griddepcontrol.wait;" ::: "memory");               // the PREVIOUS kernel is complete and flushed
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace b200rl
