// Per-sample activations, density-gradient normals, the Lambertian surface branch, tone-mapped losses,
// fused Adam and the small element-wise helpers of the MLP backward.  All HBM-bound maps / per-ray reductions.
#include "common.cuh"

namespace pnb {

constexpr float kPiF = 3.14159265358979323846f;
constexpr int kWarpsPerBlock = 8;

#define PNB_GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

// ---- activations (models/pano_mip_nerf.py:264-278) --------------------------------------------------------------
__global__ void act_fwd_kernel(long long M, int C, const float* __restrict__ raw_rgb, const float* __restrict__ raw_den,
                               float bias, float pad, float* __restrict__ rgb, float* __restrict__ density,
                               float* __restrict__ albedo) {
  PNB_GRID_STRIDE(m, M) {
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[3 * m + k] = softplus_f(raw_rgb[3 * m + k]) * (1.f + 2.f * pad) - pad;
    density[m] = softplus_f(raw_den[m * C] + bias);
    if (albedo != nullptr) {
#pragma unroll
      for (int k = 0; k < 3; ++k) albedo[3 * m + k] = (1.f / (1.f + expf(-raw_den[m * C + 1 + k]))) * 0.77f + 0.03f;
    }
  }
}

__global__ void act_bwd_kernel(long long M, int C, const float* __restrict__ raw_rgb, const float* __restrict__ raw_den,
                               float bias, float pad, const float* __restrict__ d_rgb,
                               const float* __restrict__ d_density, const float* __restrict__ d_albedo,
                               float* __restrict__ d_raw_rgb, float* __restrict__ d_raw_den) {
  PNB_GRID_STRIDE(m, M) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      d_raw_rgb[3 * m + k] = d_rgb ? d_rgb[3 * m + k] * (1.f + 2.f * pad) * softplus_d1(raw_rgb[3 * m + k]) : 0.f;
    d_raw_den[m * C] = d_density ? d_density[m] * softplus_d1(raw_den[m * C] + bias) : 0.f;
    for (int k = 1; k < C; ++k) {
      float g = 0.f;
      if (d_albedo != nullptr && k <= 3) {
        float s = 1.f / (1.f + expf(-raw_den[m * C + k]));
        g = d_albedo[3 * m + k - 1] * 0.77f * s * (1.f - s);
      }
      d_raw_den[m * C + k] = g;  // the roughness channel never reaches a loss (pano_mip_nerf.py:293-294)
    }
  }
}

__global__ void density_grad_fwd_kernel(long long M, int C, const float* __restrict__ raw_den, float bias,
                                        const float* __restrict__ v, float* __restrict__ n_raw) {
  PNB_GRID_STRIDE(m, M) {
    float s1 = softplus_d1(raw_den[m * C] + bias);
#pragma unroll
    for (int k = 0; k < 3; ++k) n_raw[3 * m + k] = -(s1 * v[3 * m + k]);
  }
}

__global__ void density_grad_bwd_kernel(long long M, int C, const float* __restrict__ raw_den, float bias,
                                        const float* __restrict__ v, const float* __restrict__ d_n,
                                        float* __restrict__ d_raw0, float* __restrict__ d_v) {
  PNB_GRID_STRIDE(m, M) {
    float x = raw_den[m * C] + bias;
    float s1 = softplus_d1(x), s2 = softplus_d2(x);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dot += d_n[3 * m + k] * v[3 * m + k];
      d_v[3 * m + k] = -(s1 * d_n[3 * m + k]);
    }
    d_raw0[m] = -(s2 * dot);
  }
}

// ---- normals, orientation loss, albedo compositing (models/pano_mip_nerf.py:296-317) -------------------------------
__device__ __forceinline__ void unit3(const float* x, float* out, float* inv) {
  float n = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  float d = fmaxf(n, 1e-12f);  // torch.nn.functional.normalize eps
  out[0] = x[0] / d, out[1] = x[1] / d, out[2] = x[2] / d;
  *inv = 1.f / d;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normals_fwd_kernel(long long R, int N, const float* __restrict__ n_raw, const float* __restrict__ weights,
                   const float* __restrict__ dirs, const float* __restrict__ albedos, float* __restrict__ normal,
                   float* __restrict__ ort, float* __restrict__ albedo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + (threadIdx.x >> 5);
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    float ws = 0.f;
    for (int i = lane; i < N; i += 32) ws += weights[r * N + i];
    ws = warp_sum(ws);
    float d[3] = {dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]};
    float s[3] = {0.f, 0.f, 0.f}, al[3] = {0.f, 0.f, 0.f}, o = 0.f;
    for (int i = lane; i < N; i += 32) {
      float nh[3], inv;
      unit3(n_raw + 3 * (r * N + i), nh, &inv);
      float wn = weights[r * N + i] / ws;
      float dot = fmaxf(nh[0] * d[0] + nh[1] * d[1] + nh[2] * d[2], 0.f);
      o += wn * (dot * dot);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s[k] += wn * nh[k];
        if (albedos) al[k] += wn * albedos[3 * (r * N + i) + k];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = warp_sum(s[k]), al[k] = warp_sum(al[k]);
    o = warp_sum(o);
    if (lane == 0) {
      float nh[3], inv;
      unit3(s, nh, &inv);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        normal[3 * r + k] = nh[k];
        if (albedo) albedo[3 * r + k] = al[k];
      }
      if (ort) ort[r] = o;
    }
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
normals_bwd_kernel(long long R, int N, const float* __restrict__ n_raw, const float* __restrict__ weights,
                   const float* __restrict__ dirs, const float* __restrict__ albedos, const float* __restrict__ g_normal,
                   const float* __restrict__ g_ort, const float* __restrict__ g_albedo, float* __restrict__ d_n_raw,
                   float* __restrict__ d_weights, float* __restrict__ d_albedos) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + (threadIdx.x >> 5);
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    float ws = 0.f;
    for (int i = lane; i < N; i += 32) ws += weights[r * N + i];
    ws = warp_sum(ws);
    float d[3] = {dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]};
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = lane; i < N; i += 32) {
      float nh[3], inv;
      unit3(n_raw + 3 * (r * N + i), nh, &inv);
      float wn = weights[r * N + i] / ws;
#pragma unroll
      for (int k = 0; k < 3; ++k) s[k] += wn * nh[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = warp_sum(s[k]);
    // through normal = s / max(|s|, eps)
    float nrm[3], sinv;
    unit3(s, nrm, &sinv);
    float gn[3] = {g_normal ? g_normal[3 * r] : 0.f, g_normal ? g_normal[3 * r + 1] : 0.f,
                   g_normal ? g_normal[3 * r + 2] : 0.f};
    float gdot = gn[0] * nrm[0] + gn[1] * nrm[1] + gn[2] * nrm[2];
    bool s_big = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) > 1e-12f;
    float ds[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ds[k] = (gn[k] - (s_big ? nrm[k] * gdot : 0.f)) * sinv;
    float go = g_ort ? g_ort[r] : 0.f;
    float ga[3] = {g_albedo ? g_albedo[3 * r] : 0.f, g_albedo ? g_albedo[3 * r + 1] : 0.f,
                   g_albedo ? g_albedo[3 * r + 2] : 0.f};
    // pass A: sum_j (dL/dwn_j) wn_j
    float corr = 0.f;
    for (int i = lane; i < N; i += 32) {
      float nh[3], inv;
      unit3(n_raw + 3 * (r * N + i), nh, &inv);
      float wn = weights[r * N + i] / ws;
      float dot = fmaxf(nh[0] * d[0] + nh[1] * d[1] + nh[2] * d[2], 0.f);
      float dwn = nh[0] * ds[0] + nh[1] * ds[1] + nh[2] * ds[2] + go * dot * dot;
      if (albedos) {
        const float* a = albedos + 3 * (r * N + i);
        dwn += a[0] * ga[0] + a[1] * ga[1] + a[2] * ga[2];
      }
      corr += dwn * wn;
    }
    corr = warp_sum(corr);
    // pass B: outputs
    for (int i = lane; i < N; i += 32) {
      const float* nr = n_raw + 3 * (r * N + i);
      float nh[3], inv;
      unit3(nr, nh, &inv);
      float wn = weights[r * N + i] / ws;
      float dotr = nh[0] * d[0] + nh[1] * d[1] + nh[2] * d[2];
      float dot = fmaxf(dotr, 0.f);
      float dwn = nh[0] * ds[0] + nh[1] * ds[1] + nh[2] * ds[2] + go * dot * dot;
      if (albedos) {
        const float* a = albedos + 3 * (r * N + i);
        dwn += a[0] * ga[0] + a[1] * ga[1] + a[2] * ga[2];
        if (d_albedos) {
#pragma unroll
          for (int k = 0; k < 3; ++k) d_albedos[3 * (r * N + i) + k] = wn * ga[k];
        }
      }
      d_weights[r * N + i] = (dwn - corr) / ws;
      float dnh[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) dnh[k] = wn * ds[k] + go * wn * 2.f * dot * d[k];
      bool big = sqrtf(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]) > 1e-12f;
      float pd = dnh[0] * nh[0] + dnh[1] * nh[1] + dnh[2] * nh[2];
#pragma unroll
      for (int k = 0; k < 3; ++k) d_n_raw[3 * (r * N + i) + k] = (dnh[k] - (big ? nh[k] * pd : 0.f)) * inv;
    }
  }
}

// ---- surface point and Lambertian shading ------------------------------------------------------------------------
__global__ void surface_point_fwd_kernel(long long R, const float* __restrict__ o, const float* __restrict__ d,
                                         const float* __restrict__ dist, float* __restrict__ pts) {
  PNB_GRID_STRIDE(i, R * 3) { pts[i] = o[i] + d[i] * dist[i / 3]; }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
surface_point_bwd_kernel(long long R, int K, const float* __restrict__ dirs, const float* __restrict__ d_means,
                         float* __restrict__ d_dist) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + (threadIdx.x >> 5);
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    float s[3] = {0.f, 0.f, 0.f};
    const float* g = d_means + r * K * 3;
    for (int i = lane; i < K; i += 32) {
      s[0] += g[3 * i], s[1] += g[3 * i + 1], s[2] += g[3 * i + 2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) d_dist[r] = s[0] * dirs[3 * r] + s[1] * dirs[3 * r + 1] + s[2] * dirs[3 * r + 2];
  }
}

// utils/surface_rendering.py:104-165, roughness=None branch
__global__ void shade_fwd_kernel(long long R, int D, const float* __restrict__ env, const float* __restrict__ albedo,
                                 const float* __restrict__ normal, const float* __restrict__ l,
                                 const float* __restrict__ omega, float* __restrict__ rgb, float* __restrict__ shading) {
  PNB_GRID_STRIDE(r, R) {
    float n[3] = {normal[3 * r], normal[3 * r + 1], normal[3 * r + 2]};
    float sh[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < D; ++k) {
      float nol = fmaxf(n[0] * l[3 * k] + n[1] * l[3 * k + 1] + n[2] * l[3 * k + 2], 0.f);
      const float* e = env + 3 * (r * D + k);
#pragma unroll
      for (int c = 0; c < 3; ++c) sh[c] += e[c] * nol * omega[k];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      shading[3 * r + c] = sh[c];
      rgb[3 * r + c] = albedo[3 * r + c] / kPiF * sh[c];
    }
  }
}

__global__ void shade_bwd_kernel(long long R, int D, const float* __restrict__ env, const float* __restrict__ albedo,
                                 const float* __restrict__ normal, const float* __restrict__ l,
                                 const float* __restrict__ omega, const float* __restrict__ g_rgb,
                                 const float* __restrict__ g_shading, float* __restrict__ d_env,
                                 float* __restrict__ d_albedo, float* __restrict__ d_normal) {
  PNB_GRID_STRIDE(r, R) {
    float n[3] = {normal[3 * r], normal[3 * r + 1], normal[3 * r + 2]};
    float sh[3] = {0.f, 0.f, 0.f}, gs[3], dn[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < D; ++k) {
      float nol = fmaxf(n[0] * l[3 * k] + n[1] * l[3 * k + 1] + n[2] * l[3 * k + 2], 0.f);
      const float* e = env + 3 * (r * D + k);
#pragma unroll
      for (int c = 0; c < 3; ++c) sh[c] += e[c] * nol * omega[k];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gr = g_rgb ? g_rgb[3 * r + c] : 0.f;
      gs[c] = (g_shading ? g_shading[3 * r + c] : 0.f) + gr * (albedo[3 * r + c] / kPiF);
      d_albedo[3 * r + c] = gr * sh[c] / kPiF;
    }
    for (int k = 0; k < D; ++k) {
      float raw = n[0] * l[3 * k] + n[1] * l[3 * k + 1] + n[2] * l[3 * k + 2];
      float nol = fmaxf(raw, 0.f);
      const float* e = env + 3 * (r * D + k);
      float dnol = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        d_env[3 * (r * D + k) + c] = gs[c] * nol * omega[k];
        dnol += gs[c] * e[c] * omega[k];
      }
      if (raw > 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] += dnol * l[3 * k + c];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) d_normal[3 * r + c] = dn[c];
  }
}

// ---- tone mapping + losses ------------------------------------------------------------------------------------
__device__ __forceinline__ float aces(float x) { return (x * (2.51f * x + 0.03f)) / (x * (2.43f * x + 0.59f) + 0.14f); }
__device__ __forceinline__ float ldr_of(float x) {
  float f = fminf(fmaxf(aces(x), 0.f), 1.f);
  return powf(f, (float)(1.0 / 2.2));
}
__device__ __forceinline__ float ldr_grad(float x) {
  float num = x * (2.51f * x + 0.03f), den = x * (2.43f * x + 0.59f) + 0.14f;
  float f = num / den;
  if (!(f >= 0.f && f <= 1.f)) return 0.f;  // torch.clamp backward mask (closed interval)
  float df = ((5.02f * x + 0.03f) * den - num * (4.86f * x + 0.59f)) / (den * den);
  const float p = (float)(1.0 / 2.2);
  return p * powf(f, p - 1.f) * df;
}

__global__ void hdr_to_ldr_kernel(long long n, const float* __restrict__ x, int quant, float* __restrict__ out) {
  PNB_GRID_STRIDE(i, n) {
    float f = fminf(fmaxf(aces(x[i]), 0.f), 1.f);
    if (quant) f = (float)(unsigned char)(f * 255.f) / 255.f;  // .to(torch.uint8) truncates
    out[i] = powf(f, (float)(1.0 / 2.2));
  }
}

__global__ void tonemap_se_fwd_kernel(long long R, const float* __restrict__ pred, const float* __restrict__ gt,
                                      const float* __restrict__ mask, float* __restrict__ partial) {
  PNB_GRID_STRIDE(r, R) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float e = ldr_of(pred[3 * r + c]) - gt[3 * r + c];
      s += mask[r] * (e * e);
    }
    partial[r] = s;
  }
}

__global__ void tonemap_se_bwd_kernel(long long R, const float* __restrict__ pred, const float* __restrict__ gt,
                                      const float* __restrict__ mask, const float* __restrict__ g_scale,
                                      float* __restrict__ d_pred) {
  const float g = g_scale[0];
  PNB_GRID_STRIDE(i, R * 3) {
    float x = pred[i];
    d_pred[i] = g * mask[i / 3] * 2.f * (ldr_of(x) - gt[i]) * ldr_grad(x);
  }
}

__global__ void chroma_fwd_kernel(long long R, const float* __restrict__ gt, const float* __restrict__ alb,
                                  float* __restrict__ partial) {
  PNB_GRID_STRIDE(r, R) {
    float a[3], b[3], ia, ib;
    unit3(gt + 3 * r, a, &ia);
    unit3(alb + 3 * r, b, &ib);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) s += (a[c] - b[c]) * (a[c] - b[c]);
    partial[r] = s;
  }
}

__global__ void chroma_bwd_kernel(long long R, const float* __restrict__ gt, const float* __restrict__ alb,
                                  const float* __restrict__ g_scale, float* __restrict__ d_alb) {
  const float g = g_scale[0];
  PNB_GRID_STRIDE(r, R) {
    float a[3], b[3], ia, ib;
    unit3(gt + 3 * r, a, &ia);
    unit3(alb + 3 * r, b, &ib);
    float db[3], pd = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      db[c] = -2.f * g * (a[c] - b[c]);
      pd += db[c] * b[c];
    }
    const float* x = alb + 3 * r;
    bool big = sqrtf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) > 1e-12f;
#pragma unroll
    for (int c = 0; c < 3; ++c) d_alb[3 * r + c] = (db[c] - (big ? b[c] * pd : 0.f)) * ib;
  }
}

// ---- deterministic sum ------------------------------------------------------------------------------------------
__global__ void sum_pass1_kernel(long long n, const float* __restrict__ x, float* __restrict__ ws) {
  __shared__ float sh[32];
  float s = 0.f;
  PNB_GRID_STRIDE(i, n) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) ws[blockIdx.x] = v;
  }
}
__global__ void sum_pass2_kernel(int nb, const float* __restrict__ ws, float scale, float* __restrict__ out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += ws[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v * scale;
  }
}

// ---- Adam (torch.optim.Adam defaults, systems/base_system.py:82) ---------------------------------------------------
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            float gscale) {
  PNB_GRID_STRIDE(i, n) {
    float gi = g[i] * gscale;
    float mi = m[i] * b1 + (1.f - b1) * gi;
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi, v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

// Same update with the step-dependent scalars read from device memory (hyper = {lr, 1-beta1^t, sqrt(1-beta2^t)}), so
// that the launch can live in a CUDA graph and still follow the learning-rate schedule.
__global__ void adam_dev_kernel(long long n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, const float* __restrict__ hyper, float b1, float b2, float eps,
                                float gscale) {
  const float lr = hyper[0], bc1 = hyper[1], bc2_sqrt = hyper[2];
  PNB_GRID_STRIDE(i, n) {
    float gi = g[i] * gscale;
    float mi = m[i] * b1 + (1.f - b1) * gi;
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi, v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
  }
}

// ---- helpers of the MLP backward -----------------------------------------------------------------------------------
template <typename TS, typename TO>
__global__ void mask_scale_kernel(long long M, int N, const TS* __restrict__ src, int ld_src, const float* __restrict__ w,
                                  const float* __restrict__ g, TO* __restrict__ out, int ld_out) {
  PNB_GRID_STRIDE(idx, M * N) {
    long long m = idx / N;
    int n = (int)(idx - m * N);
    float v = to_f32<TS>(src[m * ld_src + n]) > 0.f ? w[n] * (g ? g[m] : 1.f) : 0.f;
    out[m * ld_out + n] = from_f32<TO>(v);
  }
}

template <typename T>
__global__ void mask_mul_kernel(long long M, int N, const T* __restrict__ x, int ldx, const T* __restrict__ src,
                                int ld_src, T* __restrict__ out, int ld_out) {
  PNB_GRID_STRIDE(idx, M * N) {
    long long m = idx / N;
    int n = (int)(idx - m * N);
    out[m * ld_out + n] = to_f32<T>(src[m * ld_src + n]) > 0.f ? x[m * ldx + n] : from_f32<T>(0.f);
  }
}

// column sums: each block reduces a slab of rows for all N columns (N <= blockDim), then one atomic per column.
template <typename T>
__global__ void colsum_kernel(long long M, int N, const T* __restrict__ x, int ldx, float* __restrict__ out) {
  const int rows_per_iter = blockDim.x / N;  // threads laid out [row-in-slab][column]
  const int col = threadIdx.x % N, rsub = threadIdx.x / N;
  if (rsub >= rows_per_iter) return;
  float s = 0.f;
  for (long long m = (long long)blockIdx.x * rows_per_iter + rsub; m < M; m += (long long)gridDim.x * rows_per_iter)
    s += to_f32<T>(x[m * ldx + col]);
  atomicAdd(out + col, s);
}

template <typename TS, typename TD>
__global__ void convert_kernel(long long M, int N, const TS* __restrict__ src, int ld_src, TD* __restrict__ dst,
                               int ld_dst) {
  PNB_GRID_STRIDE(idx, M * N) {
    long long m = idx / N;
    int n = (int)(idx - m * N);
    dst[m * ld_dst + n] = from_f32<TD>(to_f32<TS>(src[m * ld_src + n]));
  }
}

}  // namespace pnb

using namespace pnb;
#define LAUNCH_1D(kernel, n, ...)                                                            \
  do {                                                                                       \
    kernel<<<grid_for((long long)(n), 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__);       \
  } while (0)
#define LAUNCH_WARP_PER_ROW(kernel, rows, ...)                                                                  \
  do {                                                                                                          \
    kernel<<<grid_for((long long)(rows) * 32, kWarpsPerBlock * 32, 8), kWarpsPerBlock * 32, 0,                  \
             as_stream(stream)>>>(__VA_ARGS__);                                                                 \
  } while (0)

extern "C" int pnb_act_fwd(int M, int C, const float* raw_rgb, const float* raw_den, float density_bias,
                           float rgb_padding, float* rgb, float* density, float* albedo, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1 && (albedo == nullptr || C >= 5), "act_fwd: albedo needs C >= 5");
  if (M == 0) return 0;
  LAUNCH_1D(act_fwd_kernel, M, M, C, raw_rgb, raw_den, density_bias, rgb_padding, rgb, density, albedo);
  return finish("act_fwd");
}

extern "C" int pnb_act_bwd(int M, int C, const float* raw_rgb, const float* raw_den, float density_bias,
                           float rgb_padding, const float* d_rgb, const float* d_density, const float* d_albedo,
                           float* d_raw_rgb, float* d_raw_den, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1 && (d_albedo == nullptr || C >= 5), "act_bwd: albedo needs C >= 5");
  if (M == 0) return 0;
  LAUNCH_1D(act_bwd_kernel, M, M, C, raw_rgb, raw_den, density_bias, rgb_padding, d_rgb, d_density, d_albedo,
            d_raw_rgb, d_raw_den);
  return finish("act_bwd");
}

extern "C" int pnb_density_grad_fwd(int M, int C, const float* raw_den, float density_bias, const float* v,
                                    float* n_raw, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1, "density_grad_fwd: bad sizes");
  if (M == 0) return 0;
  LAUNCH_1D(density_grad_fwd_kernel, M, M, C, raw_den, density_bias, v, n_raw);
  return finish("density_grad_fwd");
}

extern "C" int pnb_density_grad_bwd(int M, int C, const float* raw_den, float density_bias, const float* v,
                                    const float* d_n_raw, float* d_raw0, float* d_v, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1, "density_grad_bwd: bad sizes");
  if (M == 0) return 0;
  LAUNCH_1D(density_grad_bwd_kernel, M, M, C, raw_den, density_bias, v, d_n_raw, d_raw0, d_v);
  return finish("density_grad_bwd");
}

extern "C" int pnb_normals_fwd(int R, int N, const float* n_raw, const float* weights, const float* dirs,
                               const float* albedos, float* normal, float* ort, float* albedo, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0, "normals_fwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_WARP_PER_ROW(normals_fwd_kernel, R, R, N, n_raw, weights, dirs, albedos, normal, ort, albedo);
  return finish("normals_fwd");
}

extern "C" int pnb_normals_bwd(int R, int N, const float* n_raw, const float* weights, const float* dirs,
                               const float* albedos, const float* g_normal, const float* g_ort, const float* g_albedo,
                               float* d_n_raw, float* d_weights, float* d_albedos, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0, "normals_bwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_WARP_PER_ROW(normals_bwd_kernel, R, R, N, n_raw, weights, dirs, albedos, g_normal, g_ort, g_albedo, d_n_raw,
                      d_weights, d_albedos);
  return finish("normals_bwd");
}

extern "C" int pnb_surface_point_fwd(int R, const float* origins, const float* dirs, const float* dist, float* pts,
                                     void* stream) {
  PNB_REQUIRE(R >= 0, "surface_point_fwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_1D(surface_point_fwd_kernel, (long long)R * 3, R, origins, dirs, dist, pts);
  return finish("surface_point_fwd");
}

extern "C" int pnb_surface_point_bwd(int R, int K, const float* dirs, const float* d_means, float* d_dist,
                                     void* stream) {
  PNB_REQUIRE(R >= 0 && K > 0, "surface_point_bwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_WARP_PER_ROW(surface_point_bwd_kernel, R, R, K, dirs, d_means, d_dist);
  return finish("surface_point_bwd");
}

extern "C" int pnb_shade_fwd(int R, int D, const float* env_rgb, const float* albedo, const float* normal,
                             const float* light_dirs, const float* solid_angle, float* surface_rgb, float* shading,
                             void* stream) {
  PNB_REQUIRE(R >= 0 && D > 0, "shade_fwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_1D(shade_fwd_kernel, R, R, D, env_rgb, albedo, normal, light_dirs, solid_angle, surface_rgb, shading);
  return finish("shade_fwd");
}

extern "C" int pnb_shade_bwd(int R, int D, const float* env_rgb, const float* albedo, const float* normal,
                             const float* light_dirs, const float* solid_angle, const float* g_rgb,
                             const float* g_shading, float* d_env, float* d_albedo, float* d_normal, void* stream) {
  PNB_REQUIRE(R >= 0 && D > 0, "shade_bwd: bad sizes");
  if (R == 0) return 0;
  LAUNCH_1D(shade_bwd_kernel, R, R, D, env_rgb, albedo, normal, light_dirs, solid_angle, g_rgb, g_shading, d_env,
            d_albedo, d_normal);
  return finish("shade_bwd");
}

extern "C" int pnb_hdr_to_ldr(long long n, const float* x, int quantize_u8, float* out, void* stream) {
  PNB_REQUIRE(n >= 0, "hdr_to_ldr: bad size");
  if (n == 0) return 0;
  LAUNCH_1D(hdr_to_ldr_kernel, n, n, x, quantize_u8, out);
  return finish("hdr_to_ldr");
}

extern "C" int pnb_tonemap_se_fwd(int R, const float* pred, const float* gt_ldr, const float* mask, float* partial,
                                  void* stream) {
  PNB_REQUIRE(R >= 0, "tonemap_se_fwd: bad size");
  if (R == 0) return 0;
  LAUNCH_1D(tonemap_se_fwd_kernel, R, R, pred, gt_ldr, mask, partial);
  return finish("tonemap_se_fwd");
}

extern "C" int pnb_tonemap_se_bwd(int R, const float* pred, const float* gt_ldr, const float* mask,
                                  const float* g_scale, float* d_pred, void* stream) {
  PNB_REQUIRE(R >= 0, "tonemap_se_bwd: bad size");
  if (R == 0) return 0;
  LAUNCH_1D(tonemap_se_bwd_kernel, (long long)R * 3, R, pred, gt_ldr, mask, g_scale, d_pred);
  return finish("tonemap_se_bwd");
}

extern "C" int pnb_chroma_fwd(int R, const float* gt_ldr, const float* albedo, float* partial, void* stream) {
  PNB_REQUIRE(R >= 0, "chroma_fwd: bad size");
  if (R == 0) return 0;
  LAUNCH_1D(chroma_fwd_kernel, R, R, gt_ldr, albedo, partial);
  return finish("chroma_fwd");
}

extern "C" int pnb_chroma_bwd(int R, const float* gt_ldr, const float* albedo, const float* g_scale, float* d_albedo,
                              void* stream) {
  PNB_REQUIRE(R >= 0, "chroma_bwd: bad size");
  if (R == 0) return 0;
  LAUNCH_1D(chroma_bwd_kernel, R, R, gt_ldr, albedo, g_scale, d_albedo);
  return finish("chroma_bwd");
}

extern "C" int pnb_sum(long long n, const float* x, float scale, float* out, float* ws, void* stream) {
  PNB_REQUIRE(n >= 0 && ws != nullptr, "sum: workspace required");
  int nb = grid_for(n, 256, 4);
  if (nb > 1024) nb = 1024;
  sum_pass1_kernel<<<nb, 256, 0, as_stream(stream)>>>(n, x, ws);
  sum_pass2_kernel<<<1, 256, 0, as_stream(stream)>>>(nb, ws, scale, out);
  count_launch();
  return finish("sum");
}

extern "C" int pnb_adam_step(long long n, float* p, const float* g, float* m, float* v, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  PNB_REQUIRE(n >= 0 && step >= 1, "adam_step: step is 1-based");
  if (n == 0) return 0;
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  LAUNCH_1D(adam_kernel, n, n, p, g, m, v, lr, beta1, beta2, eps, bc1, bc2s, grad_scale);
  return finish("adam_step");
}

extern "C" int pnb_adam_step_dev(long long n, float* p, const float* g, float* m, float* v, const float* hyper,
                                 float beta1, float beta2, float eps, float grad_scale, void* stream) {
  PNB_REQUIRE(n >= 0 && hyper != nullptr, "adam_step_dev: hyper (device {lr, 1-b1^t, sqrt(1-b2^t)}) required");
  if (n == 0) return 0;
  LAUNCH_1D(adam_dev_kernel, n, n, p, g, m, v, hyper, beta1, beta2, eps, grad_scale);
  return finish("adam_step_dev");
}

extern "C" int pnb_mask_scale(long long M, int N, const void* src, int ld_src, const float* w, const float* g,
                              void* out, int ld_out, int dtype, void* stream) {
  PNB_REQUIRE(M >= 0 && N > 0, "mask_scale: bad sizes");
  if (M == 0) return 0;
  if (dtype == PNB_BF16)
    LAUNCH_1D((mask_scale_kernel<__nv_bfloat16, __nv_bfloat16>), M * N, M, N, (const __nv_bfloat16*)src, ld_src, w, g,
              (__nv_bfloat16*)out, ld_out);
  else
    LAUNCH_1D((mask_scale_kernel<float, float>), M * N, M, N, (const float*)src, ld_src, w, g, (float*)out, ld_out);
  return finish("mask_scale");
}

extern "C" int pnb_mask_mul(long long M, int N, const void* x, int ldx, const void* src, int ld_src, void* out,
                            int ld_out, int dtype, void* stream) {
  PNB_REQUIRE(M >= 0 && N > 0, "mask_mul: bad sizes");
  if (M == 0) return 0;
  if (dtype == PNB_BF16)
    LAUNCH_1D(mask_mul_kernel<__nv_bfloat16>, M * N, M, N, (const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)src,
              ld_src, (__nv_bfloat16*)out, ld_out);
  else
    LAUNCH_1D(mask_mul_kernel<float>, M * N, M, N, (const float*)x, ldx, (const float*)src, ld_src, (float*)out,
              ld_out);
  return finish("mask_mul");
}

extern "C" int pnb_colsum(long long M, int N, const void* x, int ldx, int dtype, float* out, void* stream) {
  PNB_REQUIRE(M >= 0 && N > 0 && N <= 512, "colsum: N must be <= 512");
  if (M == 0) return 0;
  int threads = N <= 256 ? 256 / N * N : N;  // whole rows per block: threads = rows_per_iter * N
  int rows_per_iter = threads / N;
  long long slabs = (M + rows_per_iter - 1) / rows_per_iter;
  long long want = (slabs + 63) / 64;  // >= 64 rows per thread before the atomic
  int grid = (int)(want < 1 ? 1 : (want > kNumSMs * 4 ? kNumSMs * 4 : want));
  if (dtype == PNB_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, threads, 0, as_stream(stream)>>>(M, N, (const __nv_bfloat16*)x, ldx, out);
  else
    colsum_kernel<float><<<grid, threads, 0, as_stream(stream)>>>(M, N, (const float*)x, ldx, out);
  return finish("colsum");
}

extern "C" int pnb_convert(long long M, int N, const void* src, int ld_src, int src_dtype, void* dst, int ld_dst,
                           int dst_dtype, void* stream) {
  PNB_REQUIRE(M >= 0 && N > 0, "convert: bad sizes");
  if (M == 0) return 0;
  if (src_dtype == PNB_F32 && dst_dtype == PNB_BF16)
    LAUNCH_1D((convert_kernel<float, __nv_bfloat16>), M * N, M, N, (const float*)src, ld_src, (__nv_bfloat16*)dst,
              ld_dst);
  else if (src_dtype == PNB_BF16 && dst_dtype == PNB_F32)
    LAUNCH_1D((convert_kernel<__nv_bfloat16, float>), M * N, M, N, (const __nv_bfloat16*)src, ld_src, (float*)dst,
              ld_dst);
  else if (src_dtype == PNB_F32 && dst_dtype == PNB_F32)
    LAUNCH_1D((convert_kernel<float, float>), M * N, M, N, (const float*)src, ld_src, (float*)dst, ld_dst);
  else
    LAUNCH_1D((convert_kernel<__nv_bfloat16, __nv_bfloat16>), M * N, M, N, (const __nv_bfloat16*)src, ld_src,
              (__nv_bfloat16*)dst, ld_dst);
  return finish("convert");
}

namespace pnb {
// per-ray sums over the `group` consecutive samples of a ray: one thread per (group, column)  [generic path]
template <typename T>
__global__ void group_sum_kernel(long long G, int N, int group, const T* __restrict__ x, int ldx,
                                 float* __restrict__ out) {
  const long long total = G * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long g = idx / N;
    int n = (int)(idx - g * N);
    float s = 0.f;
    const T* base = x + g * group * (long long)ldx + n;
    for (int i = 0; i < group; ++i) s += to_f32<T>(base[(long long)i * ldx]);
    out[idx] = s;
  }
}
// bf16 fast path: one warp per (group, 64-column slab); eight lanes cover the 128 bytes of a row with 16-byte loads, the
// four lane groups take rows i = q, q + 4, ... (four rows per instruction, 4x the bytes in flight of a 4-byte-per-lane
// walk), then two shuffle steps add the four partial rows.  Row order inside a column differs from a sequential sum
// (4 interleaved partials), as with any parallel reduction; float4 stores.
__global__ void group_sum_bf16x2_kernel(long long G, int N, int group, const __nv_bfloat16* __restrict__ x, int ldx,
                                        float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int q = lane >> 3, c8 = lane & 7;  // row phase, 16-byte column chunk
  const int slabs = N / 64;
  const long long total = G * slabs;
  for (long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); w < total;
       w += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long g = (long long)((unsigned)w / (unsigned)slabs);  // (total < 2^31, checked by the launcher)
    const int col = (int)(w - g * slabs) * 64 + 8 * c8;
    const __nv_bfloat16* base = x + g * group * (long long)ldx + col;
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
#pragma unroll 4
    for (int i = q; i < group; i += 4) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(base + (long long)i * ldx));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s[2 * k] += __uint_as_float(u[k] << 16);
        s[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], 8);
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], 16);
    }
    if (q == 0) {
      float4* o = reinterpret_cast<float4*>(out + g * N + col);
      o[0] = make_float4(s[0], s[1], s[2], s[3]);
      o[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
  }
}

// Head gradients (d raw_rgb [M,3], d raw_sigma.. [M,C]) as tensor-core operands: bf16 [M,64], zero-padded, plus their
// column sums (= the head's bias gradient) in the same pass.  Eight consecutive lanes own one 128-byte output row
// (16 bytes each), so a warp stores 512 contiguous bytes per instruction.
__global__ void pad_head_grad_kernel(long long M, int C, const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     float* __restrict__ colsum) {
  const int chunk = threadIdx.x & 7;  // columns [8*chunk, 8*chunk + 8)
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const long long rows_per_pass = (long long)gridDim.x * (blockDim.x >> 3);
  for (long long m = blockIdx.x * (long long)(blockDim.x >> 3) + (threadIdx.x >> 3); m < M; m += rows_per_pass) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      v[k] = c < C ? src[m * C + c] : 0.f;
      acc[k] += v[k];
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    reinterpret_cast<uint4*>(dst + m * 64)[chunk] = *reinterpret_cast<uint4*>(h);
  }
  if (colsum != nullptr) {
    // lanes l, l+8, l+16, l+24 share a chunk (all lanes take part in the shuffles); then one shared-memory pass over
    // the block's warps, so that only 16 atomics per BLOCK reach the few hot global addresses
    __shared__ float part[8][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = acc[k];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane < 2) part[warp][lane * 8 + k] = s;   // lane 0: columns 0..7, lane 1: columns 8..15
    }
    __syncthreads();
    if (threadIdx.x < 16 && (int)threadIdx.x < C) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w][threadIdx.x];
      atomicAdd(colsum + threadIdx.x, s);
    }
  }
}
}  // namespace pnb

extern "C" int pnb_group_sum(long long M, int N, int group, const void* x, int ldx, int dtype, float* out,
                             void* stream) {
  PNB_REQUIRE(M >= 0 && N > 0 && group > 0 && M % group == 0, "group_sum: M must be a multiple of group");
  if (M == 0) return 0;
  long long G = M / group;
  if (dtype == PNB_BF16 && N % 64 == 0 && ldx % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
      G * (N / 64) < (1ll << 31)) {
    const long long warps = G * (N / 64);
    const int grid = grid_for(warps * 32, 256, 8);
    pnb::group_sum_bf16x2_kernel<<<grid, 256, 0, as_stream(stream)>>>(G, N, group, (const __nv_bfloat16*)x, ldx, out);
  } else if (dtype == PNB_BF16)
    LAUNCH_1D(pnb::group_sum_kernel<__nv_bfloat16>, G * N, G, N, group, (const __nv_bfloat16*)x, ldx, out);
  else
    LAUNCH_1D(pnb::group_sum_kernel<float>, G * N, G, N, group, (const float*)x, ldx, out);
  return finish("group_sum");
}

extern "C" int pnb_pad_head_grad(long long M, int C, const float* src, void* dst_bf16, float* colsum, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1 && C <= 16 && (M == 0 || (src != nullptr && dst_bf16 != nullptr)),
              "pad_head_grad: bad arguments");
  PNB_REQUIRE(((uintptr_t)dst_bf16 % 16) == 0, "pad_head_grad: dst must be 16-byte aligned");
  if (M == 0) return 0;
  pnb::pad_head_grad_kernel<<<grid_for(M * 8, 256, 8), 256, 0, as_stream(stream)>>>(M, C, src, (__nv_bfloat16*)dst_bf16, colsum);
  return finish("pad_head_grad");
}
