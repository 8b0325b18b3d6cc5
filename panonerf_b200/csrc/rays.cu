// K1/K2 + encodings: equirect ray generation, stratified sampling, conical-frustum Gaussians, IPE, pos_enc.
// All kernels are HBM-bound element-wise maps; grids are sized in multiples of the 148 SMs (common.cuh).
// The translation unit is compiled with -fmad=false so that the fp32 operation order below reproduces the
// reference's un-fused PyTorch/NumPy arithmetic; explicit fmaf() is used only where fusing is harmless.
#include "common.cuh"

namespace pnb {

constexpr float kPiF = 3.14159265358979323846f;
constexpr float kHalfPiF = 1.57079632679489661923f;  // (float)(0.5*np.pi): what `y + 0.5*torch.tensor(np.pi)` adds

struct Cam {
  float r[9];
  float t[3];
};

__device__ __forceinline__ void equirect_dir(int row, int col, int H, int W, const Cam& cam, float* d, float* sin_phi) {
  // datasets/pano_datasets.py:163-175
  float theta = (-((float)col + 0.5f)) / (float)W * 2.f * kPiF;
  float phi = ((float)row + 0.5f) / (float)H * kPiF;
  float sp = sinf(phi), cp = cosf(phi);
  float x = sp * sinf(theta), y = cp, z = sp * cosf(theta);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = fmaf(z, cam.r[3 * i + 2], fmaf(y, cam.r[3 * i + 1], x * cam.r[3 * i + 0]));
  *sin_phi = sp;
}

__global__ void raygen_equirect_kernel(int H, int W, int row0, long long n, Cam cam, float near_v, float far_v,
                                       float* __restrict__ origins, float* __restrict__ directions,
                                       float* __restrict__ viewdirs, float* __restrict__ radii,
                                       float* __restrict__ lossmult, float* __restrict__ near_o,
                                       float* __restrict__ far_o, float* __restrict__ noise_var) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int row = row0 + (int)(i / W), col = (int)(i % W);
    float d[3], sp;
    equirect_dir(row, col, H, W, cam, d, &sp);
    float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    // constant-per-column radius from the middle row (pano_datasets.py:201-203); the last column re-uses W-3
    int c0 = col < W - 1 ? col : (W >= 3 ? W - 3 : 0);
    float a[3], b[3], s_;
    equirect_dir(H / 2, c0, H, W, cam, a, &s_);
    equirect_dir(H / 2, c0 + 1 < W ? c0 + 1 : c0, H, W, cam, b, &s_);
    float e0 = a[0] - b[0], e1 = a[1] - b[1], e2 = a[2] - b[2];
    float dx = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      origins[3 * i + k] = cam.t[k];
      directions[3 * i + k] = d[k];
      viewdirs[3 * i + k] = d[k] / nrm;
    }
    radii[i] = dx * 2.f / 3.46410161513775458705f;
    lossmult[i] = 1.f;
    near_o[i] = near_v;
    far_o[i] = far_v;
    noise_var[i] = sp * kPiF / (float)W;  // pano_datasets.py:170
  }
}

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

__device__ __forceinline__ float base_t(float nr, float fr, float s, int disparity) {
  // models/mip.py:135-139
  return disparity ? 1.f / (1.f / nr * (1.f - s) + 1.f / fr * s) : nr + (fr - nr) * s;
}

__device__ __forceinline__ float strat_t(int i, int N, float nr, float fr, const float* __restrict__ s_lin,
                                         const float* __restrict__ rnd, int disparity) {
  float ti = base_t(nr, fr, s_lin[i], disparity);
  if (rnd == nullptr) return ti;
  // models/mip.py:141-146: jitter between the mid-points of neighbouring fence-posts
  float lower = i == 0 ? ti : 0.5f * (ti + base_t(nr, fr, s_lin[i - 1], disparity));
  float upper = i == N ? ti : 0.5f * (base_t(nr, fr, s_lin[i + 1], disparity) + ti);
  return lower + (upper - lower) * rnd[i];
}

__global__ void sample_cast_kernel(long long R, int N, const float* __restrict__ origins, int o_div,
                                   const float* __restrict__ dirs, const float* __restrict__ radii,
                                   const float* __restrict__ near_v, const float* __restrict__ far_v, int d_mod,
                                   const float* __restrict__ s_lin, const float* __restrict__ t_rand, int rand_ld,
                                   int disparity, float* __restrict__ t_out, float* __restrict__ means,
                                   float* __restrict__ covs) {
  const long long total = R * (N + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / (N + 1);
    int i = (int)(idx - r * (N + 1));
    long long rd = d_mod ? r % d_mod : r, ro = r / o_div;
    float nr = near_v[rd], fr = far_v[rd];
    const float* rnd = t_rand ? t_rand + (long long)rand_ld * r : nullptr;
    float t0 = strat_t(i, N, nr, fr, s_lin, rnd, disparity);
    t_out[idx] = t0;
    if (i < N) {
      float t1 = strat_t(i + 1, N, nr, fr, s_lin, rnd, disparity);
      float o[3] = {origins[3 * ro], origins[3 * ro + 1], origins[3 * ro + 2]};
      float d[3] = {dirs[3 * rd], dirs[3 * rd + 1], dirs[3 * rd + 2]};
      float m[3], c[3];
      frustum_gaussian(t0, t1, radii[rd], o, d, m, c);
      long long s = r * N + i;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        means[3 * s + k] = m[k];
        covs[3 * s + k] = c[k];
      }
    }
  }
}

__global__ void cast_rays_kernel(long long R, int N, const float* __restrict__ t, const float* __restrict__ origins,
                                 int o_div, const float* __restrict__ dirs, const float* __restrict__ radii, int d_mod,
                                 float* __restrict__ means, float* __restrict__ covs) {
  const long long total = R * N;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total;
       s += (long long)gridDim.x * blockDim.x) {
    long long r = s / N;
    int i = (int)(s - r * N);
    long long rd = d_mod ? r % d_mod : r, ro = r / o_div;
    float o[3] = {origins[3 * ro], origins[3 * ro + 1], origins[3 * ro + 2]};
    float d[3] = {dirs[3 * rd], dirs[3 * rd + 1], dirs[3 * rd + 2]};
    float m[3], c[3];
    frustum_gaussian(t[r * (N + 1) + i], t[r * (N + 1) + i + 1], radii[rd], o, d, m, c);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      means[3 * s + k] = m[k];
      covs[3 * s + k] = c[k];
    }
  }
}

// ---- IPE ------------------------------------------------------------------------------------------------------
// One thread per (sample, l*3+c): writes the sin feature at column j and the cos feature at column 3L+j, so a
// warp writes two contiguous runs per sample row (coalesced).  exp underflow short-circuits the sinf slow path.
template <typename T>
__global__ void ipe_fwd_kernel(long long M, int min_deg, int L, const float* __restrict__ means,
                               const float* __restrict__ covs, T* __restrict__ enc, int ld) {
  const int F = 3 * L;
  const long long total = M * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long m = idx / F;
    int j = (int)(idx - m * F);
    int l = j / 3, c = j - 3 * l;
    float sc = exp2f((float)(min_deg + l));
    float y = means[3 * m + c] * sc;
    float yv = covs[3 * m + c] * (sc * sc);
    float e = expf(-0.5f * yv);
    float fs = 0.f, fc = 0.f;
    if (e != 0.f) {
      fs = e * sinf(y);
      fc = e * sinf(y + kHalfPiF);
    }
    enc[m * ld + j] = from_f32<T>(fs);
    enc[m * ld + F + j] = from_f32<T>(fc);
  }
}

// d enc / d mean: d/dy [e sin(y)] = e cos(y);  d/dy [e sin(y+pi/2)] = e cos(y+pi/2)   (autograd of mip.py:428)
template <typename T>
__global__ void ipe_vjp_kernel(long long M, int min_deg, int L, const float* __restrict__ means,
                               const float* __restrict__ covs, const T* __restrict__ g, int ld,
                               float* __restrict__ d_means) {
  const int F = 3 * L;
  const long long total = M * 3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long m = idx / 3;
    int c = (int)(idx - m * 3);
    float mean = means[idx], cov = covs[idx];
    float acc = 0.f;
    for (int l = 0; l < L; ++l) {
      float sc = exp2f((float)(min_deg + l));
      float e = expf(-0.5f * (cov * (sc * sc)));
      if (e == 0.f) break;  // larger l only underflow harder
      float y = mean * sc;
      float gs = to_f32<T>(g[m * ld + 3 * l + c]), gc = to_f32<T>(g[m * ld + F + 3 * l + c]);
      acc += sc * (e * (gs * cosf(y) + gc * cosf(y + kHalfPiF)));
    }
    d_means[idx] = acc;
  }
}

template <typename T>
__global__ void ipe_jvp_kernel(long long M, int min_deg, int L, const float* __restrict__ means,
                               const float* __restrict__ covs, const float* __restrict__ v, T* __restrict__ out,
                               int ld) {
  const int F = 3 * L;
  const long long total = M * F;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long m = idx / F;
    int j = (int)(idx - m * F);
    int l = j / 3, c = j - 3 * l;
    float sc = exp2f((float)(min_deg + l));
    float y = means[3 * m + c] * sc;
    float e = expf(-0.5f * (covs[3 * m + c] * (sc * sc)));
    float os = 0.f, oc = 0.f;
    if (e != 0.f) {
      float w = v[3 * m + c] * sc * e;
      os = w * cosf(y);
      oc = w * cosf(y + kHalfPiF);
    }
    out[m * ld + j] = from_f32<T>(os);
    out[m * ld + F + j] = from_f32<T>(oc);
  }
}

__global__ void pos_enc_kernel(long long R, int deg, const float* __restrict__ x, float* __restrict__ out) {
  // models/mip.py:431-441: [x | sin(2^l x) | sin(2^l x + pi/2)], l-major / xyz-minor
  const int F = 3 * deg, W = 3 + 2 * F;
  const long long total = R * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / W;
    int j = (int)(idx - r * W);
    float val;
    if (j < 3) {
      val = x[3 * r + j];
    } else {
      int q = j - 3;
      int cosine = q >= F;
      if (cosine) q -= F;
      int l = q / 3, c = q - 3 * l;
      float xb = x[3 * r + c] * exp2f((float)l);
      val = sinf(cosine ? xb + kHalfPiF : xb);
    }
    out[idx] = val;
  }
}

}  // namespace pnb

using namespace pnb;

extern "C" int pnb_raygen_equirect(int H, int W, int row0, int nrows, const float* c2w_host, float near_v, float far_v,
                                   float* origins, float* directions, float* viewdirs, float* radii, float* lossmult,
                                   float* near_o, float* far_o, float* noise_var, void* stream) {
  PNB_REQUIRE(H > 0 && W > 1 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "raygen: bad image/row range");
  PNB_REQUIRE(c2w_host != nullptr, "raygen: c2w is null");
  if (nrows == 0) return 0;
  Cam cam;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) cam.r[3 * i + j] = c2w_host[4 * i + j];
    cam.t[i] = c2w_host[4 * i + 3];
  }
  long long n = (long long)nrows * W;
  raygen_equirect_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(
      H, W, row0, n, cam, near_v, far_v, origins, directions, viewdirs, radii, lossmult, near_o, far_o, noise_var);
  return finish("raygen_equirect");
}

extern "C" int pnb_sample_cast(int R, int N, const float* origins, int o_div, const float* directions,
                               const float* radii, const float* near_v, const float* far_v, int d_mod,
                               const float* s_lin, const float* t_rand, int rand_ld, int disparity, float* t_out,
                               float* means, float* covs, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && o_div >= 1 && d_mod >= 0, "sample_cast: bad sizes");
  if (R == 0) return 0;
  sample_cast_kernel<<<grid_for((long long)R * (N + 1), 256), 256, 0, as_stream(stream)>>>(
      R, N, origins, o_div, directions, radii, near_v, far_v, d_mod, s_lin, t_rand, rand_ld, disparity, t_out, means,
      covs);
  return finish("sample_cast");
}

extern "C" int pnb_cast_rays(int R, int N, const float* t, const float* origins, int o_div, const float* directions,
                             const float* radii, int d_mod, float* means, float* covs, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && o_div >= 1 && d_mod >= 0, "cast_rays: bad sizes");
  if (R == 0) return 0;
  cast_rays_kernel<<<grid_for((long long)R * N, 256), 256, 0, as_stream(stream)>>>(R, N, t, origins, o_div,
                                                                                   directions, radii, d_mod, means,
                                                                                   covs);
  return finish("cast_rays");
}

extern "C" int pnb_ipe_fwd(int M, const float* means, const float* covs, int min_deg, int max_deg, void* enc, int ld,
                           int dtype, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && ld >= 6 * L, "ipe_fwd: bad sizes");
  if (M == 0) return 0;
  int grid = grid_for((long long)M * 3 * L, 256);
  if (dtype == PNB_BF16)
    ipe_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs,
                                                                       (__nv_bfloat16*)enc, ld);
  else
    ipe_fwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs, (float*)enc, ld);
  return finish("ipe_fwd");
}

extern "C" int pnb_ipe_vjp(int M, const float* means, const float* covs, int min_deg, int max_deg, const void* d_enc,
                           int ld, int dtype, float* d_means, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && ld >= 6 * L, "ipe_vjp: bad sizes");
  if (M == 0) return 0;
  int grid = grid_for((long long)M * 3, 256);
  if (dtype == PNB_BF16)
    ipe_vjp_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs,
                                                                       (const __nv_bfloat16*)d_enc, ld, d_means);
  else
    ipe_vjp_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs, (const float*)d_enc, ld,
                                                               d_means);
  return finish("ipe_vjp");
}

extern "C" int pnb_ipe_jvp(int M, const float* means, const float* covs, int min_deg, int max_deg, const float* v,
                           void* out, int ld, int dtype, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && ld >= 6 * L, "ipe_jvp: bad sizes");
  if (M == 0) return 0;
  int grid = grid_for((long long)M * 3 * L, 256);
  if (dtype == PNB_BF16)
    ipe_jvp_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs, v,
                                                                       (__nv_bfloat16*)out, ld);
  else
    ipe_jvp_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs, v, (float*)out, ld);
  return finish("ipe_jvp");
}

extern "C" int pnb_pos_enc(int R, const float* x, int deg, float* out, void* stream) {
  PNB_REQUIRE(R >= 0 && deg >= 0, "pos_enc: bad sizes");
  if (R == 0) return 0;
  pos_enc_kernel<<<grid_for((long long)R * (3 + 6 * deg), 256), 256, 0, as_stream(stream)>>>(R, deg, x, out);
  return finish("pos_enc");
}
