// Shared helpers for the panonerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/panonerf_b200.h"

namespace pnb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Error bookkeeping shared by all translation units (defined in api.cu).
void set_error(const char* where, cudaError_t e);
void set_error_msg(const char* msg);
void count_launch(int n = 1);

inline int finish(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(where, e);
    return (int)e;
  }
  count_launch();
  return 0;
}

#define PNB_REQUIRE(cond, msg)            \
  do {                                    \
    if (!(cond)) {                        \
      pnb::set_error_msg(msg);            \
      return PNB_ERR_ARG;                 \
    }                                     \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Grid for grid-stride element-wise kernels: enough CTAs to fill 148 SMs a few times over, never more than needed.
inline int grid_for(long long work_items, int threads, int max_ctas_per_sm = 8) {
  long long need = (work_items + threads - 1) / threads;
  long long cap = (long long)kNumSMs * max_ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- dtype-generic element access (fp32 / bf16 activations) -------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// torch.nn.Softplus(beta=1, threshold=20) and its derivatives.
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_d1(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus_d2(float x) {
  if (x > 20.f) return 0.f;
  float s = 1.f / (1.f + expf(-x));
  return s * (1.f - s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive prefix sum across the 32 lanes
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

}  // namespace pnb
