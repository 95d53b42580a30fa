// Shared device helpers for the b200audio kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include "../../include/b200audio.h"

void b2a_set_error(const char* fmt, ...);

#define B2A_CHECK_ARG(cond, msg)                                   \
  do { if (!(cond)) { b2a_set_error("%s: %s", __func__, msg); return B2A_E_INVALID; } } while (0)

#define B2A_CHECK_LAUNCH()                                         \
  do { cudaError_t e_ = cudaGetLastError();                        \
       if (e_ != cudaSuccess) { b2a_set_error("%s: %s", __func__, cudaGetErrorString(e_)); return B2A_E_CUDA; } } while (0)

// sin for the Snake activations: two-constant Cody-Waite reduction to [-pi, pi] + the SFU sine (abs error < 5e-7 there), ~6
// instructions instead of libm's ~40 -- the prologue kernels that apply Snake are otherwise instruction-bound, not HBM-bound.
__device__ __forceinline__ float b2a_sin(float x) {
  if (fabsf(x) > 8192.f) return sinf(x);
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.2831854820251465f, x);
  r = fmaf(-k, -1.7484555314695172e-07f, r);
  return __sinf(r);
}

__device__ __forceinline__ float b2a_act(float v, int act, float p0, float a, float b) {
  switch (act) {
    case B2A_ACT_LRELU: return v > 0.f ? v : v * p0;
    case B2A_ACT_SNAKE: { float s = b2a_sin(a * v); return fmaf(b, s * s, v); }
    case B2A_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case B2A_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case B2A_ACT_GELU_TANH: { float u = 0.7978845608028654f * (v + 0.044715f * v * v * v); return 0.5f * v * (1.f + tanhf(u)); }
    case B2A_ACT_TANH: return tanhf(v);
    case B2A_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case B2A_ACT_SILU: return v / (1.f + expf(-v));
    case B2A_ACT_CLIP1: return fminf(fmaxf(v, -1.f), 1.f);
    default: return v;
  }
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
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): a kernel launched with the attribute may start while its predecessor drains; it must
// not touch the predecessor's outputs (or write anything) before pdl_wait().  Everything independent of the predecessor --
// weight loads, L2 prefetches, address math -- goes before it.  Inside a CUDA graph the edges become programmatic dependencies.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool b2a_pdl_enabled();      // env B2A_PDL != "0" (api.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t b2a_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = b2a_pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
