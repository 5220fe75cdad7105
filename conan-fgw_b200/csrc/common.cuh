// Shared host/device helpers for libconanmp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

#include "conanmp.h"

namespace cmp {

void set_error(const char* fmt, ...);
void count_launch();

#define CMP_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::cmp::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// Every launch is followed by this: catches bad launch configurations without synchronising.
#define CMP_LAUNCH_CHECK(what)                                                         \
  do {                                                                                 \
    cudaError_t e__ = cudaPeekAtLastError();                                           \
    if (e__ != cudaSuccess) {                                                          \
      ::cmp::set_error("%s: CUDA launch failed: %s", (what), cudaGetErrorString(e__)); \
      (void)cudaGetLastError();                                                        \
      return CMP_ECUDA;                                                                \
    }                                                                                  \
    ::cmp::count_launch();                                                             \
  } while (0)

inline cudaStream_t as_stream(cmp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// A kernel launched through launch_pdl() may be scheduled while its predecessor in the stream is still draining: its
// prologue (barrier init, TMEM allocation, TMA of weight images that were packed many launches earlier) overlaps the
// predecessor's tail.  Contract of a PDL-aware kernel: pdl_wait() before the first access to anything the predecessor
// may have written (it returns once the predecessor has completed and its writes are visible), nothing but private
// set-up before it, and pdl_launch_dependents() once per CTA so that ITS successor can be scheduled early in turn.
// Works under stream capture (the edge becomes a programmatic dependency of the CUDA graph).
bool pdl_enabled();   // CMP_NO_PDL=1 in the environment turns every launch_pdl() into a plain launch

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();

constexpr float kLn2 = 0.69314718055994530942f;
constexpr float kPi = 3.14159265358979323846f;

// softplus(x) - ln2 with torch's threshold (F.softplus: beta=1, threshold=20).
__device__ __forceinline__ float ssp(float x) {
  float sp = (x > 20.0f) ? x : log1pf(expf(x));
  return sp - kLn2;
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float silu(float x) { return x * sigmoidf(x); }

// d/dx silu(x)
__device__ __forceinline__ float silu_grad(float x) {
  float s = sigmoidf(x);
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == CMP_ACT_SSP) return ssp(x);
  if (act == CMP_ACT_SILU) return silu(x);
  return x;
}

// SchNet cosine cutoff (PyG CFConv): no d < cutoff mask.
__device__ __forceinline__ float cos_cutoff_nomask(float d, float pi_over_cutoff) {
  return 0.5f * (cosf(d * pi_over_cutoff) + 1.0f);
}

}  // namespace cmp
