// Conical-frustum -> Gaussian conversion shared by the sampling (rays.cu) and resampling (render.cu) kernels.
// Compiled with -fmad=false: the fp32 operation order reproduces the reference's un-fused PyTorch arithmetic.
#pragma once
#include "common.cuh"

namespace pnb {

// ---- conical frustum -> Gaussian (models/mip.py:51-58) + diagonal lift (:10-22) -------------------------------
// Per-ray invariants (direction squares, the `1 - d^2 / |d|^2` factors of lift_gaussian, radius^2) are formed once;
// per sample one IEEE division (1 / (3 mu^2 + hw^2)) serves the three quotients of conical_frustum_to_gaussian and
// hw^2 / 3 is a multiplication by fl32(1/3): each quotient may differ from upstream's by one ulp of a correction term,
// far inside the 1e-6 (means) / 1e-5 (covariances) parity bounds, for a third of the instructions.
struct RayGeom {
  float o[3], d[3], dd[3], perp[3], rad2;
};
__device__ __forceinline__ RayGeom ray_geom(const float* o, const float* d, float radius) {
  RayGeom g;
  const float d0 = d[0] * d[0], d1 = d[1] * d[1], d2 = d[2] * d[2];
  const float dn = d0 + d1 + d2 + 1e-10f;
  g.dd[0] = d0, g.dd[1] = d1, g.dd[2] = d2;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g.o[k] = o[k], g.d[k] = d[k];
    g.perp[k] = 1.f - g.dd[k] / dn;
  }
  g.rad2 = radius * radius;
  return g;
}
__device__ __forceinline__ void frustum_gaussian(float t0, float t1, const RayGeom& g, float* mean, float* cov) {
  const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
  const float mu2 = mu * mu, hw2 = hw * hw;
  const float hw4 = hw2 * hw2;
  const float den = 3.f * mu2 + hw2;
  const float inv = 1.f / den;
  const float t_mean = mu + (2.f * mu * hw2) * inv;
  const float t_var = hw2 * (float)(1.0 / 3.0) - (float)(4.0 / 15.0) * ((hw4 * (12.f * mu2 - hw2)) * (inv * inv));
  const float r_var = g.rad2 * (mu2 / 4.f + (float)(5.0 / 12.0) * hw2 - (float)(4.0 / 15.0) * (hw4 * inv));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    mean[k] = g.d[k] * t_mean + g.o[k];
    cov[k] = t_var * g.dd[k] + r_var * g.perp[k];
  }
}
__device__ __forceinline__ void frustum_gaussian(float t0, float t1, float radius, const float* o, const float* d,
                                                 float* mean, float* cov) {
  frustum_gaussian(t0, t1, ray_geom(o, d, radius), mean, cov);
}

}  // namespace pnb
