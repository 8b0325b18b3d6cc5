// Conical-frustum -> Gaussian conversion shared by the sampling (rays.cu) and resampling (render.cu) kernels.
// Compiled with -fmad=false: the fp32 operation order reproduces the reference's un-fused PyTorch arithmetic.
#pragma once
#include "common.cuh"

namespace pnb {

// ---- conical frustum -> Gaussian (models/mip.py:51-58) + diagonal lift (:10-22) -------------------------------
__device__ __forceinline__ void frustum_gaussian(float t0, float t1, float radius, const float* o, const float* d,
                                                 float* mean, float* cov) {
  float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
  float mu2 = mu * mu, hw2 = hw * hw;
  float hw4 = hw2 * hw2;
  float den = 3.f * mu2 + hw2;
  float t_mean = mu + (2.f * mu * hw2) / den;
  float t_var = hw2 / 3.f - (float)(4.0 / 15.0) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
  float r_var = (radius * radius) * (mu2 / 4.f + (float)(5.0 / 12.0) * hw2 - (float)(4.0 / 15.0) * hw4 / den);
  float d0 = d[0] * d[0], d1 = d[1] * d[1], d2 = d[2] * d[2];
  float dn = d0 + d1 + d2 + 1e-10f;
  float dd[3] = {d0, d1, d2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mean[k] = d[k] * t_mean + o[k];
    cov[k] = t_var * dd[k] + r_var * (1.f - dd[k] / dn);
  }
}

}  // namespace pnb
