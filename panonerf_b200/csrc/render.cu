// K6 alpha compositing (fwd/bwd) and K7 hierarchical resampling.  One warp owns one ray: the per-ray scans run
// in registers with warp shuffles, per-ray samples are staged in shared memory, nothing per-sample round-trips
// HBM.  HBM-bound: composite fwd moves 24 B/sample + 36 B/ray, bwd 40 B/sample + 32 B/ray, resample
// 12 B/sample + 8 B/ray.
#include <cstdlib>

#include "common.cuh"
#include "geom.cuh"

namespace pnb {

constexpr int kWarpsPerBlock = 8;

// models/mip.py:510 (`volumetric_lighting_composing`): attenuation = 1 / (1 + t_mid^2)
__device__ __forceinline__ float atten_of(float t_mid) { return 1.f / (1.f + t_mid * t_mid); }

__device__ __forceinline__ float nan_to_num_f(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}

// models/mip.py:458-478
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(long long R, int N, const float* __restrict__ rgb, const float* __restrict__ density,
                     const float* __restrict__ t, const float* __restrict__ dirs, int d_mod, int white_bkgd,
                     float* __restrict__ comp_rgb, float* __restrict__ distance, float* __restrict__ acc_out,
                     float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const bool atten = (white_bkgd & PNB_COMPOSITE_ATTENUATE) != 0;
  white_bkgd &= PNB_COMPOSITE_WHITE_BKGD;
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + (threadIdx.x >> 5);
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    long long rd = d_mod ? r % d_mod : r;
    float d0 = dirs[3 * rd], d1 = dirs[3 * rd + 1], d2 = dirs[3 * rd + 2];
    float dnorm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float* tr = t + r * (N + 1);
    double carry = 0.0;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, a = 0.f, s = 0.f;
    if (N <= 64) {
      // both halves of the ray at once: all loads are issued up front and the two fp64 scans are independent chains
      // (same values, same order of additions as the generic loop below - bit-identical results)
      const int i0 = lane, i1 = lane + 32;
      const bool ok0 = i0 < N, ok1 = i1 < N;
      const float t00 = ok0 ? tr[i0] : 0.f, t01 = ok0 ? tr[i0 + 1] : 0.f;
      const float t10 = ok1 ? tr[i1] : 0.f, t11 = ok1 ? tr[i1 + 1] : 0.f;
      const float den0 = ok0 ? density[r * N + i0] : 0.f, den1 = ok1 ? density[r * N + i1] : 0.f;
      float ca[3] = {0.f, 0.f, 0.f}, cb[3] = {0.f, 0.f, 0.f};
      if (ok0) {
        const float* c = rgb + 3 * (r * N + i0);
        ca[0] = c[0], ca[1] = c[1], ca[2] = c[2];
      }
      if (ok1) {
        const float* c = rgb + 3 * (r * N + i1);
        cb[0] = c[0], cb[1] = c[1], cb[2] = c[2];
      }
      const float sd0 = ok0 ? den0 * ((t01 - t00) * dnorm) : 0.f;
      const float sd1 = ok1 ? den1 * ((t11 - t10) * dnorm) : 0.f;
      const double incl0 = warp_scan_incl((double)sd0, lane);
      const double incl1 = warp_scan_incl((double)sd1, lane);
      const double prev0 = __shfl_up_sync(0xffffffffu, incl0, 1), prev1 = __shfl_up_sync(0xffffffffu, incl1, 1);
      const double tot0 = __shfl_sync(0xffffffffu, incl0, 31);
      const double excl0 = 0.0 + (lane == 0 ? 0.0 : prev0);
      const double excl1 = (0.0 + tot0) + (lane == 0 ? 0.0 : prev1);
      const float w0 = (1.f - expf(-sd0)) * expf(-(float)excl0);
      const float w1 = (1.f - expf(-sd1)) * expf(-(float)excl1);
      if (ok0) {
        weights[r * N + i0] = w0;
        const float wa = atten ? w0 * atten_of(0.5f * (t00 + t01)) : w0;
        c0 += wa * ca[0], c1 += wa * ca[1], c2 += wa * ca[2];
        a += w0;
        s += w0 * (0.5f * (t00 + t01));
      }
      if (ok1) {
        weights[r * N + i1] = w1;
        const float wa = atten ? w1 * atten_of(0.5f * (t10 + t11)) : w1;
        c0 += wa * cb[0], c1 += wa * cb[1], c2 += wa * cb[2];
        a += w1;
        s += w1 * (0.5f * (t10 + t11));
      }
    } else
    for (int base = 0; base < N; base += 32) {
      int i = base + lane;
      bool ok = i < N;
      float t0 = ok ? tr[i] : 0.f, t1 = ok ? tr[i + 1] : 0.f;
      float sd = ok ? density[r * N + i] * ((t1 - t0) * dnorm) : 0.f;
      double incl = warp_scan_incl((double)sd, lane);
      double prev = __shfl_up_sync(0xffffffffu, incl, 1);
      double excl = carry + (lane == 0 ? 0.0 : prev);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      float w = (1.f - expf(-sd)) * expf(-(float)excl);
      if (ok) {
        weights[r * N + i] = w;
        const float* c = rgb + 3 * (r * N + i);
        const float wa = atten ? w * atten_of(0.5f * (t0 + t1)) : w;
        c0 += wa * c[0];
        c1 += wa * c[1];
        c2 += wa * c[2];
        a += w;
        s += w * (0.5f * (t0 + t1));
      }
    }
    c0 = warp_sum(c0), c1 = warp_sum(c1), c2 = warp_sum(c2), a = warp_sum(a), s = warp_sum(s);
    if (lane == 0) {
      float dist = fminf(fmaxf(nan_to_num_f(s / a), tr[0]), tr[N]);
      if (white_bkgd) {
        float bg = 1.f - a;
        c0 += bg, c1 += bg, c2 += bg;
      }
      comp_rgb[3 * r] = c0, comp_rgb[3 * r + 1] = c1, comp_rgb[3 * r + 2] = c2;
      distance[r] = dist;
      acc_out[r] = a;
    }
  }
}

// Blocked variant of the kernel above (the default for N <= 256): a ray is owned by L lanes that hold K CONSECUTIVE
// samples each, so 32 / L rays share a warp (N = 64: two rays, the env rays' N = 10: eight), density / colour / weight
// rows move as 128-bit accesses when rows are 16-byte aligned (VEC), the fp64 exclusive scan is K - 1 in-lane adds plus
// log2(L) shuffle steps, and the five per-ray sums reduce over L lanes only.  Same formulas and fp32 / fp64 types as
// above; the association of the sums differs (results agree to rounding, tests/test_kernels_gpu.py).
//
// ACT = true fuses compute_graph's activations (models/pano_mip_nerf.py:264-278, the arithmetic of act_fwd_kernel in
// shade.cu, same expressions in the same order) into the loads: `rgb` / `density` are then the RAW head outputs
// ([M,3] and [M,C], channel 0 = density, 1..3 = albedo), the activated colours and densities live in registers only,
// and the per-sample albedos (needed by the normals stage) leave as one more [M,3] row when `act.albedo` is set.
struct ActArgs {
  int C;
  float bias, pad;
  float* albedo;  // nullable
};
template <int K, int L, bool VEC, bool ACT>
__global__ void __launch_bounds__(256)
composite_fwd_blocked_kernel(int R, int N, const float* __restrict__ rgb, const float* __restrict__ density,
                             const float* __restrict__ t, const float* __restrict__ dirs, int d_mod, int white_bkgd,
                             float* __restrict__ comp_rgb, float* __restrict__ distance, float* __restrict__ acc_out,
                             float* __restrict__ weights, const ActArgs act) {
  constexpr int kRaysPerWarp = 32 / L;
  const bool atten = (white_bkgd & PNB_COMPOSITE_ATTENUATE) != 0;
  white_bkgd &= PNB_COMPOSITE_WHITE_BKGD;
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, l = lane % L;  // ray slot inside the warp, lane inside the ray
  const int j0 = l * K;                    // first sample of this lane
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r0 = (((int)blockIdx.x * (int)blockDim.x + (int)threadIdx.x) >> 5) * kRaysPerWarp; r0 < R;
       r0 += warps * kRaysPerWarp) {
    const int r = r0 + sub;
    const bool ray_ok = r < R;
    const int rr = ray_ok ? r : R - 1;  // (idle slots of the last warp recompute the last ray, stores are masked)
    const int rd = d_mod ? rr % d_mod : rr;
    const float d0 = dirs[3 * rd], d1 = dirs[3 * rd + 1], d2 = dirs[3 * rd + 2];
    const float dnorm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float* tr = t + (size_t)rr * (N + 1);
    const size_t s0 = (size_t)rr * N + j0;
    float tv[K + 1], den[K], col[3 * K];
#pragma unroll
    for (int k = 0; k <= K; ++k) tv[k] = (j0 + k <= N) ? tr[j0 + k] : 0.f;
    if (VEC) {  // N % 4 == 0 and K % 4 == 0: whole 16-byte groups are either inside or outside the ray
      if (ACT && act.C == 5) {  // [M,5] rows: the 4 samples of a group are 20 contiguous floats = five 16-byte loads
#pragma unroll
        for (int k = 0; k < K; k += 4) {
          float rw[20];
          const bool in = j0 + k < N;
#pragma unroll
          for (int v = 0; v < 5; ++v) {
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in) q = *reinterpret_cast<const float4*>(density + (s0 + k) * 5 + 4 * v);
            rw[4 * v] = q.x, rw[4 * v + 1] = q.y, rw[4 * v + 2] = q.z, rw[4 * v + 3] = q.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) den[k + i] = rw[5 * i];
          if (act.albedo != nullptr && in && ray_ok) {
            float ab[12];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int q = 0; q < 3; ++q) ab[3 * i + q] = (1.f / (1.f + expf(-rw[5 * i + 1 + q]))) * 0.77f + 0.03f;
            float4* o = reinterpret_cast<float4*>(act.albedo + 3 * (s0 + k));
#pragma unroll
            for (int v = 0; v < 3; ++v) o[v] = make_float4(ab[4 * v], ab[4 * v + 1], ab[4 * v + 2], ab[4 * v + 3]);
          }
        }
      } else if (ACT && act.C != 1) {  // raw density is channel 0 of a [M,C] row
#pragma unroll
        for (int k = 0; k < K; ++k) den[k] = (j0 + k < N) ? density[(s0 + k) * act.C] : 0.f;
      } else {
#pragma unroll
        for (int k = 0; k < K; k += 4) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j0 + k < N) v = *reinterpret_cast<const float4*>(density + s0 + k);
          den[k] = v.x, den[k + 1] = v.y, den[k + 2] = v.z, den[k + 3] = v.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 3 * K; k += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + k / 3 < N) v = *reinterpret_cast<const float4*>(rgb + 3 * s0 + k);
        col[k] = v.x, col[k + 1] = v.y, col[k + 2] = v.z, col[k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const bool ok = j0 + k < N;
        den[k] = ok ? density[(s0 + k) * (ACT ? act.C : 1)] : 0.f;
        col[3 * k] = ok ? rgb[3 * (s0 + k)] : 0.f;
        col[3 * k + 1] = ok ? rgb[3 * (s0 + k) + 1] : 0.f;
        col[3 * k + 2] = ok ? rgb[3 * (s0 + k) + 2] : 0.f;
      }
    }
    if (ACT) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (j0 + k < N) {
          den[k] = softplus_f(den[k] + act.bias);
#pragma unroll
          for (int q = 0; q < 3; ++q) col[3 * k + q] = softplus_f(col[3 * k + q]) * (1.f + 2.f * act.pad) - act.pad;
          if (act.albedo != nullptr && ray_ok && !(VEC && act.C == 5)) {  // (VEC, C = 5: written with the loads above)
            const float* ra = density + (s0 + k) * act.C + 1;
#pragma unroll
            for (int q = 0; q < 3; ++q) act.albedo[3 * (s0 + k) + q] = (1.f / (1.f + expf(-ra[q]))) * 0.77f + 0.03f;
          }
        }
      }
    }
    float sd[K];
    double ex[K];
    double run = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      sd[k] = (j0 + k < N) ? den[k] * ((tv[k + 1] - tv[k]) * dnorm) : 0.f;
      ex[k] = run;
      run += (double)sd[k];
    }
    double incl = run;  // inclusive scan of the lane totals over the L lanes of this ray
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, incl, o, L);
      if (l >= o) incl += n;
    }
    const double base = incl - run;
    float w[K];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, a = 0.f, sm = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      w[k] = (1.f - expf(-sd[k])) * expf(-(float)(base + ex[k]));
      if (j0 + k < N) {
        const float wa = atten ? w[k] * atten_of(0.5f * (tv[k] + tv[k + 1])) : w[k];
        c0 += wa * col[3 * k], c1 += wa * col[3 * k + 1], c2 += wa * col[3 * k + 2];
        a += w[k];
        sm += w[k] * (0.5f * (tv[k] + tv[k + 1]));
      }
    }
    if (ray_ok) {
      if (VEC) {
#pragma unroll
        for (int k = 0; k < K; k += 4)
          if (j0 + k < N) *reinterpret_cast<float4*>(weights + s0 + k) = make_float4(w[k], w[k + 1], w[k + 2], w[k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
          if (j0 + k < N) weights[s0 + k] = w[k];
      }
    }
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o, L), c1 += __shfl_xor_sync(0xffffffffu, c1, o, L);
      c2 += __shfl_xor_sync(0xffffffffu, c2, o, L), a += __shfl_xor_sync(0xffffffffu, a, o, L);
      sm += __shfl_xor_sync(0xffffffffu, sm, o, L);
    }
    if (l == 0 && ray_ok) {
      const float dist = fminf(fmaxf(nan_to_num_f(sm / a), tv[0]), tr[N]);
      if (white_bkgd) {
        const float bg = 1.f - a;
        c0 += bg, c1 += bg, c2 += bg;
      }
      comp_rgb[3 * r] = c0, comp_rgb[3 * r + 1] = c1, comp_rgb[3 * r + 2] = c2;
      distance[r] = dist;
      acc_out[r] = a;
    }
  }
}

// Hand-derived backward of the block above.  With sd_i = sigma_i*delta_i, T_i = exp(-sum_{j<i} sd_j),
// w_i = (1-exp(-sd_i)) T_i and G_i = dL/dw_i:   dL/dsd_i = G_i (T_i - w_i) - sum_{j>i} G_j w_j.
// ACT = true: `rgb` / `density` are the raw head outputs (see composite_fwd_blocked_kernel), the activations are
// recomputed, and the kernel applies act_bwd_kernel's chain-rule factors (shade.cu, same expressions) on the way out:
// d_rgb receives d raw_rgb [M,3], d_density receives d raw_density [M,C] (channel 0 = density, 1..3 = the albedo
// gradient `g_alb` through the sigmoid, the rest zero).
template <bool ACT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(long long R, int N, const float* __restrict__ rgb, const float* __restrict__ density,
                     const float* __restrict__ t, const float* __restrict__ dirs, int d_mod, int white_bkgd,
                     const float* __restrict__ g_comp, const float* __restrict__ g_dist,
                     const float* __restrict__ g_acc, const float* __restrict__ g_w, float* __restrict__ d_rgb,
                     float* __restrict__ d_density, const ActArgs act, const float* __restrict__ g_alb,
                     float* __restrict__ d_t) {
  extern __shared__ float smem[];
  const bool atten = (white_bkgd & PNB_COMPOSITE_ATTENUATE) != 0;
  white_bkgd &= PNB_COMPOSITE_WHITE_BKGD;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* sT = smem + (size_t)wib * 3 * N;  // transmittance
  float* sW = sT + N;                      // weights
  float* sQ = sW + N;                      // G_i * w_i
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + wib;
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    const long long rd = d_mod ? (long long)((unsigned)r % (unsigned)d_mod) : r;  // (R is an int: 32-bit modulo)
    float d0 = dirs[3 * rd], d1 = dirs[3 * rd + 1], d2 = dirs[3 * rd + 2];
    float dnorm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    const float* tr = t + r * (N + 1);
    double carry = 0.0;
    float a = 0.f, s = 0.f;
    for (int base = 0; base < N; base += 32) {
      int i = base + lane;
      bool ok = i < N;
      float t0 = ok ? tr[i] : 0.f, t1 = ok ? tr[i + 1] : 0.f;
      float den_i = 0.f;
      if (ok) den_i = ACT ? softplus_f(density[(r * N + i) * act.C] + act.bias) : density[r * N + i];
      float sd = ok ? den_i * ((t1 - t0) * dnorm) : 0.f;
      double incl = warp_scan_incl((double)sd, lane);
      double prev = __shfl_up_sync(0xffffffffu, incl, 1);
      double excl = carry + (lane == 0 ? 0.0 : prev);
      carry += __shfl_sync(0xffffffffu, incl, 31);
      float T = expf(-(float)excl);
      float w = (1.f - expf(-sd)) * T;
      if (ok) {
        sT[i] = T, sW[i] = w;
        a += w;
        s += w * (0.5f * (t0 + t1));
      }
    }
    a = warp_sum(a), s = warp_sum(s);
    float gc0 = g_comp ? g_comp[3 * r] : 0.f, gc1 = g_comp ? g_comp[3 * r + 1] : 0.f,
          gc2 = g_comp ? g_comp[3 * r + 2] : 0.f;
    float ga = (g_acc ? g_acc[r] : 0.f) - (white_bkgd ? (gc0 + gc1 + gc2) : 0.f);
    float draw = s / a;
    float dnum = nan_to_num_f(draw);
    // torch.clamp passes the gradient on the closed interval, nan_to_num only for finite inputs
    bool pass = isfinite(draw) && dnum >= tr[0] && dnum <= tr[N];
    float gd = (g_dist && pass) ? g_dist[r] / a : 0.f;
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
      const float* cr = rgb + 3 * (r * N + i);
      float c[3] = {cr[0], cr[1], cr[2]};
      if (ACT) {
#pragma unroll
        for (int q = 0; q < 3; ++q) c[q] = softplus_f(c[q]) * (1.f + 2.f * act.pad) - act.pad;
      }
      float tm = 0.5f * (tr[i] + tr[i + 1]);
      const float att = atten ? atten_of(tm) : 1.f;
      float G = (gc0 * c[0] + gc1 * c[1] + gc2 * c[2]) * att + ga + (g_w ? g_w[r * N + i] : 0.f);
      if (gd != 0.f) G += gd * (tm - draw);  // `draw` is NaN on an empty ray; the reference would propagate it
      float w = sW[i];
      sQ[i] = G * w;
      sT[i] = G * (sT[i] - w);  // re-use: G_i (T_i - w_i)
      float* o = d_rgb + 3 * (r * N + i);
      const float wa = w * att;
      if (ACT) {  // act_bwd_kernel: d_rgb * (1 + 2 pad) * softplus'(raw)   (0 without any colour gradient)
        const float gk[3] = {wa * gc0, wa * gc1, wa * gc2};
#pragma unroll
        for (int q = 0; q < 3; ++q) o[q] = g_comp ? gk[q] * (1.f + 2.f * act.pad) * softplus_d1(cr[q]) : 0.f;
      } else {
        o[0] = wa * gc0, o[1] = wa * gc1, o[2] = wa * gc2;
      }
    }
    __syncwarp();
    // suffix-exclusive sum of Q, walking the chunks from the far end of the ray
    float tail = 0.f;
    for (int base = ((N - 1) / 32) * 32; base >= 0; base -= 32) {
      int i = base + (31 - lane);  // lane 0 holds the farthest sample of the chunk
      bool ok = i < N;
      float q = ok ? sQ[i] : 0.f;
      float incl = warp_scan_incl(q, lane);
      float excl = tail + incl - q;
      tail += __shfl_sync(0xffffffffu, incl, 31);
      if (ok && d_t != nullptr) {
        // dL/d t (stop_resample_grad = False): sd_i = sigma_i (t_{i+1} - t_i) |d| and the distance's t_mid_i
        const float sig = ACT ? softplus_f(density[(r * N + i) * act.C] + act.bias) : density[r * N + i];
        const float ds = (sT[i] - excl) * (sig * dnorm), dm = 0.5f * (gd * sW[i]);
        sQ[i] = dm - ds;  // share of t_i      (this lane has consumed sQ[i] above; sW[i] is not read again)
        sW[i] = dm + ds;  // share of t_{i+1}
      }
      if (ok) {
        const float dd = (sT[i] - excl) * ((tr[i + 1] - tr[i]) * dnorm);
        if (ACT) {
          const float* rw = density + (r * N + i) * act.C;
          float* od = d_density + (r * N + i) * act.C;
          od[0] = dd * softplus_d1(rw[0] + act.bias);
          for (int k = 1; k < act.C; ++k) {
            float g = 0.f;
            if (g_alb != nullptr && k <= 3) {
              const float sg = 1.f / (1.f + expf(-rw[k]));
              g = g_alb[3 * (r * N + i) + k - 1] * 0.77f * sg * (1.f - sg);
            }
            od[k] = g;
          }
        } else {
          d_density[r * N + i] = dd;
        }
      }
    }
    __syncwarp();
    if (d_t != nullptr) {
      float* dt = d_t + r * (N + 1);
      for (int j = lane; j <= N; j += 32) {
        float g = (j < N ? sQ[j] : 0.f) + (j > 0 ? sW[j - 1] : 0.f);
        // torch.clamp(dist, t_0, t_N) hands the gradient to the bound that clips
        if (g_dist != nullptr && !pass && (j == 0 || j == N)) {
          if (j == 0 && dnum < tr[0]) g += g_dist[r];
          if (j == N && dnum > tr[N]) g += g_dist[r];
        }
        dt[j] = g;
      }
      __syncwarp();
    }
  }
}

// models/mip.py:324-329 (blur-pool) + 253-300 (pad, pdf, cdf, searchsorted(right=True), gather, lerp), optionally
// followed by cast_rays on the new fence-posts (mip.py:351) in the same kernel.  One warp per ray, KMAX = ceil(N/32)
// samples per lane.  Bit-exactness notes (the indices must equal torch's on the CPU reference path):
//   * torch.sum over the contiguous sample axis is NOT a plain left-to-right sum: ATen's CPU kernel (SumKernel.cpp,
//     vectorized_inner_sum / row_sum) keeps 4 interleaved accumulators of 8-lane vectors, i.e. element e of a row
//     goes to accumulator (e / 8) % 4, lane e % 8; then acc0 + acc1 + acc2 + acc3, then the scalar tail (N % 8) and
//     the 8 lanes are added left to right.  Lane L of the warp plays accumulator L / 8, vector lane L % 8, so the
//     strided layout i = L + 32 k reproduces that order exactly (checked against torch.sum on the host for every N
//     the tests use; N < 8 takes ATen's scalar path: 4 interleaved scalar accumulators);
//   * torch.cumsum on the CPU accumulates in double and rounds every prefix to fp32: the scan runs in fp64.
//
// BWD = true is the backward pass of the same function for `stop_resample_grad = False` (the else branch of
// models/mip.py:336-350): the forward quantities (blur, sum, pdf, cdf, indices) are recomputed exactly as above, then
// `g_t` = dL/d new_t [R,N+1] (arriving through `new_t`) is pulled back through the lerp onto the two gathered cdf
// entries (shared-memory atomics), through min(1, cumsum) (suffix sums; torch's tie rule of `minimum`: half at
// equality), the normalisation (with the 1e-5 padding branch) and the blur-pool's maxima (torch.maximum: the larger
// operand takes the gradient, ties split) into dL/d weights [R,N] (written to `d_weights`).  Bins and u carry no gradient.
// FULL: N == 32 * KMAX exactly (64, 128, 256: the configurations of the YAMLs), blur-pool on, no index output: the
// ragged-row / short-row paths and the bounds tests of the lane loops fold away at compile time (same arithmetic).
template <int KMAX, bool BWD, bool FULL = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
resample_kernel(long long R, int N, const float* __restrict__ t, const float* __restrict__ weights, float padding,
                int blur_pool, const float* __restrict__ u, int u_ld, float* __restrict__ new_t,
                long long* __restrict__ inds_out, const float* __restrict__ origins, const float* __restrict__ dirs,
                const float* __restrict__ radii, float* __restrict__ means, float* __restrict__ covs, int stage_cast,
                float* __restrict__ d_weights) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int stride = 3 * N + 4 + (stage_cast ? 6 * N : 0) + (BWD ? 2 * N + 4 : 0);
  float* sa = smem + (size_t)wib * stride;  // raw weights, later the cdf (N+1 entries)
  float* sp = sa + N + 1;                   // blurred weights, then the pdf, then the new fence-posts (N+1)
  float* sb = sp + N + 1;                   // bins (t), N+1 entries
  float* sg = sb + N + 2;                   // staged Gaussians [3N | 3N] (16-byte aligned: the stride is a multiple of 4)
  float* sdc = sb + N + 2;                  // BWD: dL/d cdf (N+1), later dL/d (blurred weight) (N)
  float* smk = sdc + N + 2;                 // BWD: pass factor of min(1, cumsum) per cdf entry (N+1)
  if (FULL) N = 32 * KMAX, blur_pool = 1, inds_out = nullptr;
  const int K = FULL ? KMAX : (N + 31) >> 5;  // samples per lane
  const int vec = N >> 3, ilp = vec >> 2;     // ATen: 8-float vectors, 4 interleaved accumulators
  const bool ragged = FULL ? false : (N & 31) != 0;
  const long long warp0 = blockIdx.x * (long long)kWarpsPerBlock + wib;
  for (long long r = warp0; r < R; r += (long long)gridDim.x * kWarpsPerBlock) {
    const float* wr = weights + r * N;
    const float* tr = t + r * (N + 1);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int i = lane + 32 * k;
      if (i < N) sa[i] = wr[i];
      if (i <= N) sb[i] = tr[i];
    }
    if (lane == 0 && 32 * KMAX <= N) sb[N] = tr[N];  // (N == 32 * KMAX: the last fence-post)
    __syncwarp();
    // blur-pool: wm[k] = max(w[max(k-1,0)], w[min(k,N-1)]), blur[i] = .5 (wm[i] + wm[i+1]) + padding
    float blur[KMAX];
    float local = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int i = lane + 32 * k;
      blur[k] = 0.f;
      if (i < N) {
        const float wc = sa[i];
        if (blur_pool) {
          const float wl = sa[i > 0 ? i - 1 : 0], wrt = sa[i + 1 < N ? i + 1 : N - 1];
          blur[k] = 0.5f * (fmaxf(wl, wc) + fmaxf(wc, wrt)) + padding;
        } else {
          blur[k] = wc;
        }
        if (ragged || N < 8) sp[i] = blur[k];
      }
      if (k < ilp) local += blur[k];  // accumulator lane/8, vector lane%8: elements i4*32 + lane, i4 ascending
    }
    float wsum;
    if (!FULL && N < 8) {  // ATen scalar path: 4 interleaved scalar accumulators, leftovers into accumulator 0
      __syncwarp();
      float p4[4] = {0.f, 0.f, 0.f, 0.f};
      const int q4 = N >> 2;
      for (int i = 0; i < q4; ++i)
        for (int k = 0; k < 4; ++k) p4[k] += sp[i * 4 + k];
      for (int i = q4 * 4; i < N; ++i) p4[0] += sp[i];
      wsum = ((p4[0] + p4[1]) + p4[2]) + p4[3];
    } else {
      if (ragged) {  // vectors beyond the last full group of 4 go to accumulator 0, then the scalar tail
        __syncwarp();
        if (lane < 8)
          for (int j = ilp * 4; j < vec; ++j) local += sp[j * 8 + lane];
      }
      float p = local + __shfl_down_sync(0xffffffffu, local, 8);
      p += __shfl_down_sync(0xffffffffu, local, 16);
      p += __shfl_down_sync(0xffffffffu, local, 24);
      float acc = 0.f;
      if (ragged)
        for (int e = vec * 8; e < N; ++e) acc += sp[e];
#pragma unroll
      for (int l = 0; l < 8; ++l) acc += __shfl_sync(0xffffffffu, p, l);
      wsum = acc;
    }
    __syncwarp();  // every lane is done with the raw / blurred values in shared memory
    const float pad = fmaxf(0.f, 1e-5f - wsum);  // mip.py:253-257
    const float add = pad / (float)N;
    wsum += pad;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int i = lane + 32 * k;
      if (i < N) sp[i] = (blur[k] + add) / wsum;
    }
    __syncwarp();
    // cdf[i] = min(1, float(sum_{j<i} double(pdf_j))), lane-blocked: lane owns samples [lane*K, lane*K + K)
    {
      double e[KMAX];
      double run = 0.0;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int i = lane * K + j;
        e[j] = run;
        if (j < K && i < N) run += (double)sp[i];
      }
      const double incl = warp_scan_incl(run, lane);
      const double base = incl - run;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int i = lane * K + j;
        if (j < K && i < N) {
          const float pre = (float)(base + e[j]);
          sa[i] = i == 0 ? 0.f : fminf(1.f, pre);
          if (BWD) smk[i] = i == 0 ? 0.f : (pre < 1.f ? 1.f : (pre == 1.f ? 0.5f : 0.f)), sdc[i] = 0.f;
        }
      }
      if (lane == 0) {
        sa[N] = 1.f;
        if (BWD) smk[N] = 0.f, sdc[N] = 0.f;
      }
    }
    __syncwarp();
    // searchsorted(cdf, u, right=True): lane owns Q consecutive outputs; u is sorted along a ray (linspace, or the
    // stratified draw of mip.py:271-276), so after one binary search the index only moves forwards - a merge walk of
    // a few steps instead of a binary search per output (an unsorted u falls back to the binary search)
    const float* ur = u + (long long)u_ld * r;
    constexpr int Q = KMAX + 1;
    {
      int lo = 0;
      float prev_u = 0.f;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = lane * Q + q;
        if (j > N) break;
        const float uj = ur[j];
        if (q == 0 || uj < prev_u) {
          int a = 0, b = N + 1;  // first index with cdf > u
          while (a < b) {
            const int mid = (a + b) >> 1;
            if (sa[mid] <= uj) a = mid + 1; else b = mid;
          }
          lo = a;
        } else {
          while (lo <= N && sa[lo] <= uj) ++lo;
        }
        prev_u = uj;
        const int below = lo - 1 > 0 ? lo - 1 : 0, above = lo < N ? lo : N;
        const float c0 = sa[below], c1 = sa[above];
        float den = c1 - c0;
        if (den < 1e-5f) den = 1.f;
        const float frac = (uj - c0) / den;
        const float b0 = sb[below], b1 = sb[above];
        if (BWD) {
          const float g = new_t[r * (N + 1) + j] * (b1 - b0);  // (`new_t` carries dL/d new_t in this mode)
          if (c1 - c0 < 1e-5f) {
            atomicAdd(&sdc[below], -g);
          } else {
            const float inv2 = 1.f / (den * den);
            atomicAdd(&sdc[below], g * (uj - c1) * inv2);
            atomicAdd(&sdc[above], -(g * (uj - c0)) * inv2);
          }
          continue;
        }
        const float nt = b0 + frac * (b1 - b0);
        new_t[r * (N + 1) + j] = nt;
        sp[j] = nt;
        if (inds_out) inds_out[r * (N + 1) + j] = lo;
      }
    }
    __syncwarp();
    if (BWD) {
      // dL/d pdf_k = sum_{i = k+1}^{N-1} gcdf_i (cdf_i = sum_{k < i} pdf_k, i = 1..N-1), gcdf = dL/d cdf * pass factor
      double sfx[KMAX];
      double run = 0.0;
#pragma unroll
      for (int j = KMAX - 1; j >= 0; --j) {
        const int i = lane * K + j;  // entry i contributes to pdf_k, k < i
        sfx[j] = run;                // sum of the lane's entries above i
        if (j < K && i < N) run += (double)(sdc[i] * smk[i]);
      }
      // exclusive suffix over the lanes: lanes above this one
      double incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += n;
      }
      const double above_lanes = incl - run;
      float dpdf[KMAX];
      float dot = 0.f, tot = 0.f;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int k = lane * K + j;
        dpdf[j] = 0.f;
        if (j < K && k < N) {
          dpdf[j] = (float)(above_lanes + sfx[j]);  // entries i > k (entry k itself belongs to pdf_{k-1} and below)
          dot += dpdf[j] * sp[k];
          tot += dpdf[j];
        }
      }
      dot = warp_sum(dot), tot = warp_sum(tot);
      __syncwarp();  // every lane has read its dL/d cdf entries: the array is re-used for dL/d (blurred weight)
      const float sub = pad > 0.f ? tot / (float)N : dot;  // (pad > 0: weights' = w + (eps - S) / N, sum' = eps)
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const int k = lane * K + j;
        if (j < K && k < N) sdc[k] = (dpdf[j] - sub) / wsum;
      }
      __syncwarp();
      float* dw = d_weights + r * N;
      for (int i = lane; i < N; i += 32) {
        if (!blur_pool) {
          dw[i] = sdc[i];
          continue;
        }
        // blur[i] = .5 (wm[i] + wm[i+1]), wm[k] = max(w[max(k-1,0)], w[min(k,N-1)]), k = 0..N
        auto dwm = [&](int k) { return 0.5f * ((k >= 1 ? sdc[k - 1] : 0.f) + (k <= N - 1 ? sdc[k] : 0.f)); };
        const float wi = wr[i];
        float g = 0.f;
        if (i == 0) g += dwm(0);
        else {
          const float wl = wr[i - 1];
          g += (wi > wl ? 1.f : (wi == wl ? 0.5f : 0.f)) * dwm(i);
        }
        if (i == N - 1) g += dwm(N);
        else {
          const float wn = wr[i + 1];
          g += (wi > wn ? 1.f : (wi == wn ? 0.5f : 0.f)) * dwm(i + 1);
        }
        dw[i] = g;
      }
      __syncwarp();
      continue;
    }
    if (means != nullptr) {  // cast_rays on the new fence-posts (models/mip.py:351, 67-89)
      const float o[3] = {origins[3 * r], origins[3 * r + 1], origins[3 * r + 2]};
      const float d[3] = {dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]};
      const RayGeom geom = ray_geom(o, d, radii[r]);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int i = lane + 32 * k;
        if (i < N) {
          float m[3], c[3];
          frustum_gaussian(sp[i], sp[i + 1], geom, m, c);
          if (stage_cast) {  // 16-byte rows through shared memory (see sample_cast_kernel)
#pragma unroll
            for (int q = 0; q < 3; ++q) sg[3 * i + q] = m[q], sg[3 * N + 3 * i + q] = c[q];
          } else {
            const long long sidx = r * N + i;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
              means[3 * sidx + q] = m[q];
              covs[3 * sidx + q] = c[q];
            }
          }
        }
      }
      if (stage_cast) {
        __syncwarp();
        float4* gm = reinterpret_cast<float4*>(means + 3 * r * N);
        float4* gv = reinterpret_cast<float4*>(covs + 3 * r * N);
        const int nv = (3 * N) >> 2;
        for (int v = lane; v < nv; v += 32) {
          gm[v] = reinterpret_cast<const float4*>(sg)[v];
          gv[v] = reinterpret_cast<const float4*>(sg + 3 * N)[v];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace pnb

using namespace pnb;
// CTAs queued per SM by the warp-per-ray kernels (resampling, compositing backward): 4 -> 16 measured -7 % on the resampler
static int pnb_rays_ctas() { return 16; }

template <bool ACT>
static int launch_composite_fwd(int R, int N, const float* rgb, const float* density, const float* t,
                                const float* dirs, int d_mod, int white_bkgd, float* comp_rgb, float* distance,
                                float* acc, float* weights, const ActArgs act, cudaStream_t st) {
  const bool vec = N % 4 == 0 && ((uintptr_t)rgb % 16 == 0) && ((uintptr_t)density % 16 == 0) &&
                   ((uintptr_t)weights % 16 == 0) && ((uintptr_t)act.albedo % 16 == 0);
  static const bool legacy = getenv("PNB_COMPOSITE_LEGACY") != nullptr;  // A/B experiments
#define PNB_LAUNCH_COMPOSITE(K, L)                                                                                   \
  do {                                                                                                               \
    const long long warps = ((long long)R + 32 / (L) - 1) / (32 / (L));                                              \
    const int grid = grid_for(warps * 32, 256, 8);                                                                   \
    if (vec)                                                                                                         \
      composite_fwd_blocked_kernel<K, L, true, ACT><<<grid, 256, 0, st>>>(                                          \
          R, N, rgb, density, t, dirs, d_mod, white_bkgd, comp_rgb, distance, acc, weights, act);                    \
    else                                                                                                             \
      composite_fwd_blocked_kernel<K, L, false, ACT><<<grid, 256, 0, st>>>(                                         \
          R, N, rgb, density, t, dirs, d_mod, white_bkgd, comp_rgb, distance, acc, weights, act);                    \
    return finish("composite_fwd");                                                                                  \
  } while (0)
  if (!legacy || ACT) {
    if (N <= 16) PNB_LAUNCH_COMPOSITE(4, 4);
    if (N <= 32) PNB_LAUNCH_COMPOSITE(4, 8);
    if (N <= 64) PNB_LAUNCH_COMPOSITE(4, 16);
    if (N <= 128) PNB_LAUNCH_COMPOSITE(4, 32);
    if (N <= 256) PNB_LAUNCH_COMPOSITE(8, 32);
  }
#undef PNB_LAUNCH_COMPOSITE
  if (ACT) {
    set_error_msg("act_composite_fwd: N > 256 is not supported by the fused kernel");
    return PNB_ERR_ARG;
  }
  int grid = grid_for((long long)R * 32, kWarpsPerBlock * 32, 8);
  composite_fwd_kernel<<<grid, kWarpsPerBlock * 32, 0, st>>>(R, N, rgb, density, t, dirs, d_mod, white_bkgd, comp_rgb,
                                                             distance, acc, weights);
  return finish("composite_fwd");
}

extern "C" int pnb_composite_fwd(int R, int N, const float* rgb, const float* density, const float* t,
                                 const float* dirs, int d_mod, int white_bkgd, float* comp_rgb, float* distance,
                                 float* acc, float* weights, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && d_mod >= 0, "composite_fwd: bad sizes");
  if (R == 0) return 0;
  return launch_composite_fwd<false>(R, N, rgb, density, t, dirs, d_mod, white_bkgd, comp_rgb, distance, acc, weights,
                                     ActArgs{1, 0.f, 0.f, nullptr}, as_stream(stream));
}

extern "C" int pnb_act_composite_fwd(int R, int N, int C, const float* raw_rgb, const float* raw_den,
                                     float density_bias, float rgb_padding, const float* t, const float* dirs,
                                     int d_mod, int white_bkgd, float* comp_rgb, float* distance, float* acc,
                                     float* weights, float* albedo, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && N <= 256 && d_mod >= 0 && C >= 1 && (albedo == nullptr || C >= 4),
              "act_composite_fwd: bad sizes (N <= 256, albedo needs C >= 4)");
  if (R == 0) return 0;
  return launch_composite_fwd<true>(R, N, raw_rgb, raw_den, t, dirs, d_mod, white_bkgd, comp_rgb, distance, acc,
                                    weights, ActArgs{C, density_bias, rgb_padding, albedo}, as_stream(stream));
}

extern "C" int pnb_composite_bwd(int R, int N, const float* rgb, const float* density, const float* t,
                                 const float* dirs, int d_mod, int white_bkgd, const float* g_comp,
                                 const float* g_dist, const float* g_acc, const float* g_weights, float* d_rgb,
                                 float* d_density, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && d_mod >= 0, "composite_bwd: bad sizes");
  if (R == 0) return 0;
  size_t smem = (size_t)kWarpsPerBlock * 3 * N * sizeof(float);
  PNB_REQUIRE(smem <= 200 * 1024, "composite_bwd: N too large for the per-warp shared-memory staging");
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(composite_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = grid_for((long long)R * 32, kWarpsPerBlock * 32, pnb_rays_ctas());
  composite_bwd_kernel<false><<<grid, kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
      R, N, rgb, density, t, dirs, d_mod, white_bkgd, g_comp, g_dist, g_acc, g_weights, d_rgb, d_density,
      ActArgs{1, 0.f, 0.f, nullptr}, nullptr, nullptr);
  return finish("composite_bwd");
}

extern "C" int pnb_act_composite_bwd(int R, int N, int C, const float* raw_rgb, const float* raw_den,
                                     float density_bias, float rgb_padding, const float* t, const float* dirs,
                                     int d_mod, int white_bkgd, const float* g_comp, const float* g_dist,
                                     const float* g_acc, const float* g_weights, const float* g_albedo,
                                     float* d_raw_rgb, float* d_raw_den, float* d_t, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && d_mod >= 0 && C >= 1 && (g_albedo == nullptr || C >= 4),
              "act_composite_bwd: bad sizes (albedo needs C >= 4)");
  PNB_REQUIRE(d_t == nullptr || (white_bkgd & PNB_COMPOSITE_ATTENUATE) == 0,
              "act_composite_bwd: no fence-post gradient with the attenuation flag");
  if (R == 0) return 0;
  size_t smem = (size_t)kWarpsPerBlock * 3 * N * sizeof(float);
  PNB_REQUIRE(smem <= 200 * 1024, "act_composite_bwd: N too large for the per-warp shared-memory staging");
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(composite_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = grid_for((long long)R * 32, kWarpsPerBlock * 32, pnb_rays_ctas());
  composite_bwd_kernel<true><<<grid, kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
      R, N, raw_rgb, raw_den, t, dirs, d_mod, white_bkgd, g_comp, g_dist, g_acc, g_weights, d_raw_rgb, d_raw_den,
      ActArgs{C, density_bias, rgb_padding, nullptr}, g_albedo, d_t);
  return finish("act_composite_bwd");
}

template <int KMAX>
static int launch_resample(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                           const float* u, int u_ld, float* new_t, long long* inds, const float* origins,
                           const float* dirs, const float* radii, float* means, float* covs, cudaStream_t st) {
  const int stage = means != nullptr && N % 4 == 0 && ((uintptr_t)means % 16 == 0) && ((uintptr_t)covs % 16 == 0);
  const size_t smem = (size_t)kWarpsPerBlock * (3 * N + 4 + (stage ? 6 * N : 0)) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(resample_kernel<KMAX, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = grid_for((long long)R * 32, kWarpsPerBlock * 32, pnb_rays_ctas());
  if (N == 32 * KMAX && blur_pool && inds == nullptr) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(resample_kernel<KMAX, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resample_kernel<KMAX, false, true><<<grid, kWarpsPerBlock * 32, smem, st>>>(
        R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, origins, dirs, radii, means, covs, stage, nullptr);
    return finish("resample");
  }
  resample_kernel<KMAX, false><<<grid, kWarpsPerBlock * 32, smem, st>>>(
      R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, origins, dirs, radii, means, covs, stage, nullptr);
  return finish("resample");
}

template <int KMAX>
static int launch_resample_bwd(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                               const float* u, int u_ld, const float* g_new_t, float* d_weights, cudaStream_t st) {
  const size_t smem = (size_t)kWarpsPerBlock * (3 * N + 4 + 2 * N + 4) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(resample_kernel<KMAX, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = grid_for((long long)R * 32, kWarpsPerBlock * 32, pnb_rays_ctas());
  resample_kernel<KMAX, true><<<grid, kWarpsPerBlock * 32, smem, st>>>(
      R, N, t, weights, padding, blur_pool, u, u_ld, const_cast<float*>(g_new_t), nullptr, nullptr, nullptr, nullptr,
      nullptr, nullptr, 0, d_weights);
  return finish("resample_bwd");
}

extern "C" int pnb_resample_bwd(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                                const float* u, int u_ld, const float* g_new_t, float* d_weights, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && N <= 256 && (u_ld == 0 || u_ld >= N + 1), "resample_bwd: need 0 < N <= 256");
  PNB_REQUIRE(R == 0 || (t && weights && u && g_new_t && d_weights), "resample_bwd: null argument");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int k = (N + 31) / 32;
  if (k <= 1) return launch_resample_bwd<1>(R, N, t, weights, padding, blur_pool, u, u_ld, g_new_t, d_weights, st);
  if (k <= 2) return launch_resample_bwd<2>(R, N, t, weights, padding, blur_pool, u, u_ld, g_new_t, d_weights, st);
  if (k <= 4) return launch_resample_bwd<4>(R, N, t, weights, padding, blur_pool, u, u_ld, g_new_t, d_weights, st);
  return launch_resample_bwd<8>(R, N, t, weights, padding, blur_pool, u, u_ld, g_new_t, d_weights, st);
}

extern "C" int pnb_resample_cast(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                                 const float* u, int u_ld, float* new_t, long long* inds, const float* origins,
                                 const float* directions, const float* radii, float* means, float* covs,
                                 void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && N <= 256 && (u_ld == 0 || u_ld >= N + 1), "resample: need 0 < N <= 256");
  PNB_REQUIRE(R == 0 || (t && weights && u && new_t), "resample: null argument");
  PNB_REQUIRE((means == nullptr) == (covs == nullptr), "resample: means and covs go together");
  PNB_REQUIRE(R == 0 || means == nullptr || (origins && directions && radii), "resample: the fused cast needs the rays");
  if (R == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (N <= 64)
    return launch_resample<2>(R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, origins, directions, radii,
                              means, covs, st);
  if (N <= 128)
    return launch_resample<4>(R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, origins, directions, radii,
                              means, covs, st);
  return launch_resample<8>(R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, origins, directions, radii,
                            means, covs, st);
}

extern "C" int pnb_resample(int R, int N, const float* t, const float* weights, float padding, int blur_pool,
                            const float* u, int u_ld, float* new_t, long long* inds, void* stream) {
  return pnb_resample_cast(R, N, t, weights, padding, blur_pool, u, u_ld, new_t, inds, nullptr, nullptr, nullptr,
                           nullptr, nullptr, stream);
}
