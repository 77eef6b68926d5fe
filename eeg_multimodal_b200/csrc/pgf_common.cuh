// Shared device/host helpers for the pgfuse kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pgfuse kernels are written for sm_100a (B200) only"
#endif

#define PGF_OK 0
#define PGF_ERR_ARG 1
#define PGF_ERR_CUDA 2
#define PGF_ERR_UNSUPPORTED 3
#define PGF_ERR_WORKSPACE 4

#define PGF_DT_F32 0
#define PGF_DT_BF16 1

// noise modes of the perturb/gate kernel
#define PGF_NOISE_INJECTED 0  // Laplace / Gumbel tensors supplied by the caller (parity tests)
#define PGF_NOISE_PHILOX 1    // counter-based Philox4x32-10 (production)
#define PGF_NOISE_NONE 2      // non-private ConcatModel path (reference model.py:53-64)

namespace pgf {

void set_error(const char* fmt, ...);

#define PGF_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      pgf::set_error(__VA_ARGS__);        \
      return PGF_ERR_ARG;                 \
    }                                     \
  } while (0)

#define PGF_CUDA_LAUNCH_CHECK(name)                                              \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      pgf::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
      return PGF_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define PGF_CUDA_CALL(call)                                                      \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      pgf::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return PGF_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

int num_sms();
// Per-(device, kernel) caches of cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and of the occupancy query: both are
// driver calls of several microseconds, which is a visible share of a launch in the launch-bound regimes.
void ensure_dynamic_smem(const void* kernel, size_t smem);
int cached_occupancy(const void* kernel, int threads, size_t smem, int fallback);

// ---- programmatic dependent launch (the fused B<=8 sweep step, sweep_step.cu) ---------------------------------------
// Every kernel of that chain is written as  [prologue that touches nothing its predecessor writes]  griddep_wait()
// griddep_launch()  [body].  Launched with the stream-serialization attribute (pdl_scope active) the next kernel's CTAs
// become resident while the predecessor's last wave drains, run their prologue (weight prefetch, Philox bits, smem
// carve-up) and block in griddep_wait() until the predecessor has completed and flushed.  launch_dependents is issued
// only AFTER the kernel's own wait, so completion is transitive: a prologue may read anything except what the
// IMMEDIATE predecessor writes.  Without the attribute both instructions are no-ops (ordinary stream order).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// thread-local switch read by launch(): set by the sweep-step plan while it enqueues its chain
bool pdl_enabled();
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on);
  ~PdlScope();
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Device-resident step state of a training loop (one per engine / plan): what changes from step to step lives here
// instead of in kernel arguments, so a whole step is a constant launch sequence (CUDA graph, no host in the loop).
// Kernels ADD these to their ordinary arguments (NULL state = plain arguments).  Written only by step_state_set /
// step_advance (sweep_step.cu); the Adam coefficients are those of step t+1, i.e. of the NEXT update.
struct StepState {
  long long noise_offset;   // + Philox offset argument
  long long t_dp;           // Adam steps already taken by the DP group
  long long t_model;        // ... by the weight group
  long long cursor;         // position of the current batch in the resident dataset / permutation
  float dp_step_size, dp_bc2_sqrt;        // lr/(1-b1^t), sqrt(1-b2^t) at t = t_dp + 1
  float model_step_size, model_bc2_sqrt;  // same at t = t_model + 1
  float lr, b1, b2, pad;                  // what step_advance needs to form the next coefficients
};
static_assert(sizeof(StepState) == 64, "StepState is part of the C ABI (pgf_step_state_*): 64 bytes");

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit global accesses (data touched once: keep it out of L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_u2(uint2* p, const uint2& v) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

}  // namespace pgf
