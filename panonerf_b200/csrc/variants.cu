// Variants of the reference's surface branch that its hot path does not call today:
//   * specular BRDFs on the predicted roughness channel: utils/surface_rendering.py:6-61 (`microfeast_brdf`, GGX /
//     Schlick / Smith for image-based lighting) and :64-101 (`blinn_phong_brdf`), and the `roughness is not None`
//     branch of `surface_rendering` (:147-151,159) that sums them over the D light directions;
//   * `RotToTarget.rot2t` (utils/vector_rotation.py:57-89): the rotation that takes the +y axis onto a target
//     vector (hemisphere of light directions around a surface normal).
// One thread per (ray, direction) or per ray; these are tiny per-ray maps (HBM-bound, a few dozen bytes per unit).
// The backward kernels evaluate the SAME templated forward code on dual numbers (value + tangents w.r.t. the
// differentiable inputs) and contract the Jacobian with the incoming gradient, so forward and backward cannot drift.
#include <math.h>

#include "common.cuh"

namespace pnb {

#define PNB_GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

constexpr float kPiV = 3.14159265358979323846f;

// ---- dual numbers ---------------------------------------------------------------------------------------------
template <int N>
struct Dual {
  float v;
  float d[N];
};
template <int N>
__device__ __forceinline__ Dual<N> dconst(float v) {
  Dual<N> r;
  r.v = v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = 0.f;
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> dvar(float v, int k) {
  Dual<N> r = dconst<N>(v);
  r.d[k] = 1.f;
  return r;
}
// y = f(x) with derivative df
template <int N>
__device__ __forceinline__ Dual<N> lift(const Dual<N>& x, float f, float df) {
  Dual<N> r;
  r.v = f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = df * x.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r;
  r.v = a.v / b.v;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
template <int N> __device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, float b) { return lift(a, a.v + b, 1.f); }
template <int N> __device__ __forceinline__ Dual<N> operator+(float b, const Dual<N>& a) { return lift(a, b + a.v, 1.f); }
template <int N> __device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, float b) { return lift(a, a.v - b, 1.f); }
template <int N> __device__ __forceinline__ Dual<N> operator-(float b, const Dual<N>& a) { return lift(a, b - a.v, -1.f); }
template <int N> __device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, float b) { return lift(a, a.v * b, b); }
template <int N> __device__ __forceinline__ Dual<N> operator*(float b, const Dual<N>& a) { return lift(a, b * a.v, b); }
template <int N> __device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, float b) { return lift(a, a.v / b, 1.f / b); }
template <int N> __device__ __forceinline__ Dual<N> operator/(float b, const Dual<N>& a) {
  return lift(a, b / a.v, -b / (a.v * a.v));
}

// scalar functions, overloaded for float and Dual
__device__ __forceinline__ float value_of(float x) { return x; }
template <int N> __device__ __forceinline__ float value_of(const Dual<N>& x) { return x.v; }
__device__ __forceinline__ float relu_s(float x) { return fmaxf(x, 0.f); }
template <int N> __device__ __forceinline__ Dual<N> relu_s(const Dual<N>& x) {   // torch.relu: gradient 1 for x > 0
  return lift(x, fmaxf(x.v, 0.f), x.v > 0.f ? 1.f : 0.f);
}
__device__ __forceinline__ float sqrt_s(float x) { return sqrtf(x); }
template <int N> __device__ __forceinline__ Dual<N> sqrt_s(const Dual<N>& x) {
  const float s = sqrtf(x.v);
  return lift(x, s, 0.5f / s);
}
__device__ __forceinline__ float sin_s(float x) { return sinf(x); }
template <int N> __device__ __forceinline__ Dual<N> sin_s(const Dual<N>& x) { return lift(x, sinf(x.v), cosf(x.v)); }
__device__ __forceinline__ float cos_s(float x) { return cosf(x); }
template <int N> __device__ __forceinline__ Dual<N> cos_s(const Dual<N>& x) { return lift(x, cosf(x.v), -sinf(x.v)); }
__device__ __forceinline__ float acos_s(float x) { return acosf(x); }
template <int N> __device__ __forceinline__ Dual<N> acos_s(const Dual<N>& x) {
  return lift(x, acosf(x.v), -1.f / sqrtf(1.f - x.v * x.v));
}
// base ** x for a positive constant base
__device__ __forceinline__ float powc_s(float base, float x) { return powf(base, x); }
template <int N> __device__ __forceinline__ Dual<N> powc_s(float base, const Dual<N>& x) {
  const float p = powf(base, x.v);
  return lift(x, p, p * logf(base));
}
// x ** y (x >= 0): d/dx = y x^(y-1), d/dy = x^y log x; both taken as 0 at x == 0 (torch masks the log term there and
// the caller's relu has zero slope at 0)
__device__ __forceinline__ float pow_s(float x, float y) { return powf(x, y); }
template <int N> __device__ __forceinline__ Dual<N> pow_s(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r;
  r.v = powf(x.v, y.v);
  const float dx = x.v > 0.f ? y.v * powf(x.v, y.v - 1.f) : 0.f;
  const float dy = x.v > 0.f ? r.v * logf(x.v) : 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) r.d[i] = dx * x.d[i] + dy * y.d[i];
  return r;
}
template <class T> __device__ __forceinline__ T zero_like(const T&);
template <> __device__ __forceinline__ float zero_like<float>(const float&) { return 0.f; }
template <> __device__ __forceinline__ Dual<4> zero_like<Dual<4>>(const Dual<4>&) { return dconst<4>(0.f); }

template <class T>
__device__ __forceinline__ T dot3(const T n[3], const float b[3]) {
  return (n[0] * b[0] + n[1] * b[1]) + n[2] * b[2];
}

// ---- specular BRDF terms of one (ray, light direction) pair --------------------------------------------------------
// kind 0: utils/surface_rendering.py:27-59 (microfacet), kind 1: :85-99 (Blinn-Phong).  n, r carry the tangents.
// Returns the specular ratio and NoL (clamped for kind 0, raw for kind 1 - as upstream returns them).
// `nan_to_num(nan=0, posinf=0)` (:58, :98): a light at or below the horizon gives 0/0 upstream; here such an entry is
// 0 with ZERO gradient (upstream's gradient through the masked division is NaN for the whole ray - DESIGN.md section 2
// "deviations kept visible").
template <class T>
__device__ __forceinline__ void brdf_terms(int kind, const T n[3], const T& r, const float l[3], const float v[3], T* spec,
                                           T* nol_out) {
  float h[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
  const float hn = fmaxf(sqrtf((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]), 1e-12f);  // F.normalize (eps 1e-12)
#pragma unroll
  for (int k = 0; k < 3; ++k) h[k] = h[k] / hn;
  const T noh = relu_s(dot3(n, h));
  if (kind == 1) {
    *nol_out = dot3(n, l);
    T s = pow_s(noh, r);
    const float sv = value_of(s);
    *spec = (isnan(sv) || (isinf(sv) && sv > 0.f)) ? zero_like(s) : s;
    return;
  }
  const float voh = fmaxf((v[0] * h[0] + v[1] * h[1]) + v[2] * h[2], 0.f);
  const T nol = relu_s(dot3(n, l));
  const T nov = relu_s(dot3(n, v));
  *nol_out = nol;
  const T alpha = r * r;
  const T k = (r * r) / 2.f;
  const T a2 = alpha * alpha;
  const T dden = (noh * noh) * (a2 - 1.f) + 1.f;
  const T dist = a2 / (kPiV * (dden * dden));
  const float fres = 0.04f + (1.f - 0.04f) * powf(2.f, -(5.55473f * voh + 6.98316f) * voh);
  const T g1 = nol / ((1.f - k) * nol + k);
  const T g2 = nov / ((1.f - k) * nov + k);
  const T den = (4.f * nol) * nov;
  if (!(value_of(den) > 0.f)) {  // 0/0 -> nan -> 0 upstream
    *spec = zero_like(den);
    return;
  }
  T s = ((dist * fres) * (g1 * g2)) / den;
  const float sv = value_of(s);
  *spec = (isnan(sv) || (isinf(sv) && sv > 0.f)) ? zero_like(s) : s;
}

__global__ void brdf_terms_fwd_kernel(int kind, long long R, int D, const float* __restrict__ normal,
                                      const float* __restrict__ rough, const float* __restrict__ l, int l_per_ray,
                                      const float* __restrict__ v, float* __restrict__ spec, float* __restrict__ nol) {
  PNB_GRID_STRIDE(i, R * D) {
    const long long r = i / D;
    const int k = (int)(i - r * D);
    const float n[3] = {normal[3 * r], normal[3 * r + 1], normal[3 * r + 2]};
    const float* lp = l + 3 * (l_per_ray ? i : k);
    const float lv[3] = {lp[0], lp[1], lp[2]};
    const float vv[3] = {v[3 * r], v[3 * r + 1], v[3 * r + 2]};
    float s, c;
    brdf_terms<float>(kind, n, rough[r], lv, vv, &s, &c);
    spec[i] = s, nol[i] = c;
  }
}

// d_normal[r] / d_rough[r] = sum over the D directions of  g_spec * d spec + g_nol * d NoL
__global__ void brdf_terms_bwd_kernel(int kind, long long R, int D, const float* __restrict__ normal,
                                      const float* __restrict__ rough, const float* __restrict__ l, int l_per_ray,
                                      const float* __restrict__ v, const float* __restrict__ g_spec,
                                      const float* __restrict__ g_nol, float* __restrict__ d_normal,
                                      float* __restrict__ d_rough) {
  PNB_GRID_STRIDE(r, R) {
    Dual<4> n[3] = {dvar<4>(normal[3 * r], 0), dvar<4>(normal[3 * r + 1], 1), dvar<4>(normal[3 * r + 2], 2)};
    const Dual<4> rg = dvar<4>(rough[r], 3);
    const float vv[3] = {v[3 * r], v[3 * r + 1], v[3 * r + 2]};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < D; ++k) {
      const float* lp = l + 3 * (l_per_ray ? r * D + k : k);
      const float lv[3] = {lp[0], lp[1], lp[2]};
      Dual<4> s, c;
      brdf_terms<Dual<4>>(kind, n, rg, lv, vv, &s, &c);
      const float gs = g_spec ? g_spec[r * D + k] : 0.f, gn = g_nol ? g_nol[r * D + k] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += gs * s.d[j] + gn * c.d[j];
    }
    d_normal[3 * r] = acc[0], d_normal[3 * r + 1] = acc[1], d_normal[3 * r + 2] = acc[2];
    d_rough[r] = acc[3];
  }
}

// utils/surface_rendering.py:149-151,159: diffuse = sum_d (albedo/pi) env NoL omega, specular = sum_d spec env omega
__global__ void shade_sum_fwd_kernel(long long R, int D, const float* __restrict__ env, const float* __restrict__ albedo,
                                     const float* __restrict__ spec, const float* __restrict__ nol,
                                     const float* __restrict__ omega, float* __restrict__ rgb,
                                     float* __restrict__ diffuse, float* __restrict__ specular) {
  PNB_GRID_STRIDE(r, R) {
    float df[3] = {0.f, 0.f, 0.f}, sp[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < D; ++k) {
      const float* e = env + 3 * (r * D + k);
      const float c = nol[r * D + k], s = spec[r * D + k], w = omega[k];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        df[ch] += ((albedo[3 * r + ch] / kPiV) * e[ch]) * c * w;
        sp[ch] += (s * e[ch]) * w;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      diffuse[3 * r + ch] = df[ch], specular[3 * r + ch] = sp[ch];
      rgb[3 * r + ch] = df[ch] + sp[ch];
    }
  }
}

__global__ void shade_sum_bwd_kernel(long long R, int D, const float* __restrict__ env, const float* __restrict__ albedo,
                                     const float* __restrict__ spec, const float* __restrict__ nol,
                                     const float* __restrict__ omega, const float* __restrict__ g_rgb,
                                     const float* __restrict__ g_diffuse, const float* __restrict__ g_specular,
                                     float* __restrict__ d_env, float* __restrict__ d_albedo, float* __restrict__ d_spec,
                                     float* __restrict__ d_nol) {
  PNB_GRID_STRIDE(r, R) {
    float gd[3], gs[3], da[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float g = g_rgb ? g_rgb[3 * r + ch] : 0.f;
      gd[ch] = g + (g_diffuse ? g_diffuse[3 * r + ch] : 0.f);
      gs[ch] = g + (g_specular ? g_specular[3 * r + ch] : 0.f);
    }
    for (int k = 0; k < D; ++k) {
      const float* e = env + 3 * (r * D + k);
      const float c = nol[r * D + k], s = spec[r * D + k], w = omega[k];
      float dc = 0.f, ds = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float a = albedo[3 * r + ch] / kPiV;
        d_env[3 * (r * D + k) + ch] = gd[ch] * a * c * w + gs[ch] * s * w;
        da[ch] += gd[ch] * e[ch] * c * w;
        dc += gd[ch] * a * e[ch] * w;
        ds += gs[ch] * e[ch] * w;
      }
      d_nol[r * D + k] = dc, d_spec[r * D + k] = ds;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) d_albedo[3 * r + ch] = da[ch] / kPiV;
  }
}

// ---- RotToTarget.rot2t: utils/vector_rotation.py:57-89 ---------------------------------------------------------------
template <class T>
__device__ __forceinline__ void rot_to_target(const T t[3], T rm[9], bool* flip) {
  const T theta = acos_s(t[1]);  // (0,1,0) . tvec
  *flip = value_of(theta) == kPiV;
  // cross((0,1,0), t) = (t_z, 0, -t_x), then F.normalize (eps 1e-12)
  const T len2 = t[2] * t[2] + t[0] * t[0];
  T a0, a2;
  if (value_of(len2) > 1e-24f) {
    const T len = sqrt_s(len2);
    a0 = t[2] / len, a2 = (0.f - t[0]) / len;
  } else {
    a0 = t[2] / 1e-12f, a2 = (0.f - t[0]) / 1e-12f;
  }
  const T s = sin_s(theta), c1 = 1.f - cos_s(theta);
  // skew K = [[0,-a2,0],[a2,0,-a0],[0,a0,0]] (a1 == 0);  K K = [[-a2^2, 0, a0 a2],[0, -(a0^2+a2^2), 0],[a0 a2, 0, -a0^2]]
  rm[0] = 1.f + (0.f - a2 * a2) * c1;
  rm[1] = s * (0.f - a2);
  rm[2] = (a0 * a2) * c1;
  rm[3] = s * a2;
  rm[4] = 1.f + (0.f - (a0 * a0 + a2 * a2)) * c1;
  rm[5] = s * (0.f - a0);
  rm[6] = (a0 * a2) * c1;
  rm[7] = s * a0;
  rm[8] = 1.f + (0.f - a0 * a0) * c1;
}

__global__ void rot_to_target_fwd_kernel(long long R, const float* __restrict__ tvec, float* __restrict__ rot) {
  PNB_GRID_STRIDE(r, R) {
    const float t[3] = {tvec[3 * r], tvec[3 * r + 1], tvec[3 * r + 2]};
    float rm[9];
    bool flip;
    rot_to_target<float>(t, rm, &flip);
    if (flip) {  // theta == pi: the fixed matrix diag(1,-1,1) (:87)
      rm[0] = 1.f, rm[1] = 0.f, rm[2] = 0.f, rm[3] = 0.f, rm[4] = -1.f, rm[5] = 0.f, rm[6] = 0.f, rm[7] = 0.f, rm[8] = 1.f;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) rot[9 * r + k] = rm[k];
  }
}

__global__ void rot_to_target_bwd_kernel(long long R, const float* __restrict__ tvec, const float* __restrict__ g_rot,
                                         float* __restrict__ d_tvec) {
  PNB_GRID_STRIDE(r, R) {
    const Dual<3> t[3] = {dvar<3>(tvec[3 * r], 0), dvar<3>(tvec[3 * r + 1], 1), dvar<3>(tvec[3 * r + 2], 2)};
    Dual<3> rm[9];
    bool flip;
    rot_to_target<Dual<3>>(t, rm, &flip);
    float acc[3] = {0.f, 0.f, 0.f};
    if (!flip) {  // the overwritten rows carry no gradient
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float g = g_rot[9 * r + k];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[j] += g * rm[k].d[j];
      }
    }
    d_tvec[3 * r] = acc[0], d_tvec[3 * r + 1] = acc[1], d_tvec[3 * r + 2] = acc[2];
  }
}

}  // namespace pnb

using namespace pnb;

extern "C" int pnb_brdf_terms_fwd(int kind, int R, int D, const float* normal, const float* roughness, const float* l,
                                  int l_per_ray, const float* v, float* spec, float* nol, void* stream) {
  PNB_REQUIRE((kind == 0 || kind == 1) && R >= 0 && D > 0, "brdf_terms_fwd: bad arguments");
  if (R == 0) return 0;
  brdf_terms_fwd_kernel<<<grid_for((long long)R * D, 256), 256, 0, as_stream(stream)>>>(kind, R, D, normal, roughness, l,
                                                                                        l_per_ray, v, spec, nol);
  return finish("brdf_terms_fwd");
}

extern "C" int pnb_brdf_terms_bwd(int kind, int R, int D, const float* normal, const float* roughness, const float* l,
                                  int l_per_ray, const float* v, const float* g_spec, const float* g_nol,
                                  float* d_normal, float* d_roughness, void* stream) {
  PNB_REQUIRE((kind == 0 || kind == 1) && R >= 0 && D > 0, "brdf_terms_bwd: bad arguments");
  if (R == 0) return 0;
  brdf_terms_bwd_kernel<<<grid_for(R, 128), 128, 0, as_stream(stream)>>>(kind, R, D, normal, roughness, l, l_per_ray, v,
                                                                         g_spec, g_nol, d_normal, d_roughness);
  return finish("brdf_terms_bwd");
}

extern "C" int pnb_shade_sum_fwd(int R, int D, const float* env_rgb, const float* albedo, const float* spec,
                                 const float* nol, const float* solid_angle, float* rgb, float* diffuse,
                                 float* specular, void* stream) {
  PNB_REQUIRE(R >= 0 && D > 0, "shade_sum_fwd: bad sizes");
  if (R == 0) return 0;
  shade_sum_fwd_kernel<<<grid_for(R, 128), 128, 0, as_stream(stream)>>>(R, D, env_rgb, albedo, spec, nol, solid_angle,
                                                                        rgb, diffuse, specular);
  return finish("shade_sum_fwd");
}

extern "C" int pnb_shade_sum_bwd(int R, int D, const float* env_rgb, const float* albedo, const float* spec,
                                 const float* nol, const float* solid_angle, const float* g_rgb, const float* g_diffuse,
                                 const float* g_specular, float* d_env, float* d_albedo, float* d_spec, float* d_nol,
                                 void* stream) {
  PNB_REQUIRE(R >= 0 && D > 0, "shade_sum_bwd: bad sizes");
  if (R == 0) return 0;
  shade_sum_bwd_kernel<<<grid_for(R, 128), 128, 0, as_stream(stream)>>>(R, D, env_rgb, albedo, spec, nol, solid_angle,
                                                                        g_rgb, g_diffuse, g_specular, d_env, d_albedo,
                                                                        d_spec, d_nol);
  return finish("shade_sum_bwd");
}

extern "C" int pnb_rot_to_target_fwd(int R, const float* tvec, float* rot, void* stream) {
  PNB_REQUIRE(R >= 0, "rot_to_target_fwd: bad size");
  if (R == 0) return 0;
  rot_to_target_fwd_kernel<<<grid_for(R, 128), 128, 0, as_stream(stream)>>>(R, tvec, rot);
  return finish("rot_to_target_fwd");
}

extern "C" int pnb_rot_to_target_bwd(int R, const float* tvec, const float* g_rot, float* d_tvec, void* stream) {
  PNB_REQUIRE(R >= 0, "rot_to_target_bwd: bad size");
  if (R == 0) return 0;
  rot_to_target_bwd_kernel<<<grid_for(R, 128), 128, 0, as_stream(stream)>>>(R, tvec, g_rot, d_tvec);
  return finish("rot_to_target_bwd");
}
