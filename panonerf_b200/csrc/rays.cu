// K1/K2 + encodings: equirect ray generation, stratified sampling, conical-frustum Gaussians, IPE, pos_enc.
// All kernels are HBM-bound element-wise maps; grids are sized in multiples of the 148 SMs (common.cuh).
// The translation unit is compiled with -fmad=false so that the fp32 operation order below reproduces the
// reference's un-fused PyTorch/NumPy arithmetic; explicit fmaf() is used only where fusing is harmless.
#include <cstdlib>

#include "common.cuh"
#include "geom.cuh"

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

// One warp per ray: the per-ray constants are loaded once, the lanes walk the fence-posts, and there is no 64-bit
// index division per sample (the flat-index version spent most of its instructions there).
// STAGE: the Gaussians of a ray are staged in shared memory and leave as 16-byte rows (3 N floats per array, N % 4 == 0)
// instead of 4-byte stores at a 12-byte lane stride, which hit every 32-byte sector from three instructions.
// MODE: 0 = every run-time option; 1 / 2 = the main rays' calls (one origin / direction / radius row per ray, linear
// spacing), deterministic / randomized: the 64-bit index divisions and the option branches fold away at compile time.
template <bool STAGE, int MODE>
__global__ void __launch_bounds__(256)
sample_cast_kernel(long long R, int N, const float* __restrict__ origins, int o_div,
                   const float* __restrict__ dirs, const float* __restrict__ radii,
                   const float* __restrict__ near_v, const float* __restrict__ far_v, int d_mod, int dir_mod,
                   const float* __restrict__ s_lin, const float* __restrict__ t_rand, int rand_ld,
                   int disparity, float* __restrict__ t_out, float* __restrict__ means,
                   float* __restrict__ covs) {
  extern __shared__ float sc_smem[];
  const int lane = threadIdx.x & 31;
  float* sm = sc_smem + (size_t)(threadIdx.x >> 5) * 6 * N;  // [3N means | 3N covs] of this warp's ray
  float* sv = sm + 3 * N;
  const long long warp0 = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
  // The nine per-ray scalars of the NEXT ray are requested before the current ray is worked on: a warp otherwise starts
  // every ray with a dependent ~700 ns load (ncu: long-scoreboard stalls dominated, 50 % occupancy could not hide them).
  float nxt[9];
  if (MODE != 0) disparity = 0;
  if (MODE == 1) t_rand = nullptr;
  auto fetch = [&](long long r) {
    // (R is an int at the C boundary: 32-bit divisions instead of the 64-bit subroutines)
    const unsigned ru = (unsigned)r;
    const long long rd = (MODE == 0 && d_mod) ? ru % (unsigned)d_mod : r, ro = (MODE == 0 && o_div != 1) ? ru / (unsigned)o_div : r;
    const long long rdir = (MODE == 0 && dir_mod) ? ru % (unsigned)dir_mod : r;  // (sample_each_points_hemisp: one direction per ray)
    nxt[0] = near_v[rd], nxt[1] = far_v[rd], nxt[2] = radii[rd];
#pragma unroll
    for (int k = 0; k < 3; ++k) nxt[3 + k] = origins[3 * ro + k], nxt[6 + k] = dirs[3 * rdir + k];
  };
  if (warp0 < R) fetch(warp0);
  for (long long r = warp0; r < R; r += wstride) {
    const float nr = nxt[0], fr = nxt[1], rad = nxt[2];
    const float o[3] = {nxt[3], nxt[4], nxt[5]};
    const float d[3] = {nxt[6], nxt[7], nxt[8]};
    if (r + wstride < R) fetch(r + wstride);
    const float* rnd = t_rand ? t_rand + (long long)rand_ld * r : nullptr;
    float* trow = t_out + r * (N + 1);
    const RayGeom geom = ray_geom(o, d, rad);
    for (int i = lane; i < N; i += 32) {  // (the last fence-post rides with the last interval: no 1-lane extra round)
      const float t0 = strat_t(i, N, nr, fr, s_lin, rnd, disparity);
      const float t1 = strat_t(i + 1, N, nr, fr, s_lin, rnd, disparity);
      trow[i] = t0;
      if (i == N - 1) trow[N] = t1;
      float m[3], c[3];
      frustum_gaussian(t0, t1, geom, m, c);
      if (STAGE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) sm[3 * i + k] = m[k], sv[3 * i + k] = c[k];
      } else {
        const long long sidx = r * N + i;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          means[3 * sidx + k] = m[k];
          covs[3 * sidx + k] = c[k];
        }
      }
    }
    if (STAGE) {
      __syncwarp();
      float4* gm = reinterpret_cast<float4*>(means + 3 * r * N);
      float4* gv = reinterpret_cast<float4*>(covs + 3 * r * N);
      const int nv = (3 * N) >> 2;
      for (int v = lane; v < nv; v += 32) {
        gm[v] = reinterpret_cast<const float4*>(sm)[v];
        gv[v] = reinterpret_cast<const float4*>(sv)[v];
      }
      __syncwarp();
    }
  }
}

// Short rays (the env rays' N = 10: a warp-per-ray walk would idle 22 of 32 lanes): one thread per sample, the per-ray
// scalars come through L1 (a ray's samples sit in neighbouring lanes), same device functions - same bits.
__global__ void __launch_bounds__(256)
sample_cast_short_kernel(unsigned total, int N, const float* __restrict__ origins, int o_div,
                         const float* __restrict__ dirs, const float* __restrict__ radii,
                         const float* __restrict__ near_v, const float* __restrict__ far_v, int d_mod, int dir_mod,
                         const float* __restrict__ s_lin, const float* __restrict__ t_rand, int rand_ld,
                         int disparity, float* __restrict__ t_out, float* __restrict__ means,
                         float* __restrict__ covs) {
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += gridDim.x * blockDim.x) {
    const unsigned r = s / (unsigned)N;
    const int i = (int)(s - r * (unsigned)N);
    const unsigned rd = d_mod ? r % (unsigned)d_mod : r, ro = o_div != 1 ? r / (unsigned)o_div : r;
    const unsigned rdir = dir_mod ? r % (unsigned)dir_mod : r;
    const float nr = near_v[rd], fr = far_v[rd];
    const float o[3] = {origins[3 * ro], origins[3 * ro + 1], origins[3 * ro + 2]};
    const float d[3] = {dirs[3 * rdir], dirs[3 * rdir + 1], dirs[3 * rdir + 2]};
    const float* rnd = t_rand ? t_rand + (size_t)rand_ld * r : nullptr;
    const RayGeom geom = ray_geom(o, d, radii[rd]);
    const float t0 = strat_t(i, N, nr, fr, s_lin, rnd, disparity);
    const float t1 = strat_t(i + 1, N, nr, fr, s_lin, rnd, disparity);
    float* trow = t_out + (size_t)r * (N + 1);
    trow[i] = t0;
    if (i == N - 1) trow[N] = t1;
    float m[3], c[3];
    frustum_gaussian(t0, t1, geom, m, c);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      means[3 * (size_t)s + k] = m[k];
      covs[3 * (size_t)s + k] = c[k];
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
// sin / cos of fp32 arguments up to ~2^15 * |x| (scale 2^15 of the highest IPE degree): libdevice's sinf()/cosf()
// fall into the Payne-Hanek slow path (local memory, ~100s of instructions) above |y| ~ 1e5, which made these
// kernels compute-bound at 0.12 of the HBM roofline.  Here the argument is reduced in double precision
// (y - n*pi/2 with a two-term pi/2: exact to ~1e-11 for |y| < 1e8) and the Cephes minimax polynomials are evaluated
// in fp32 on [-pi/4, pi/4] (<= 2 ulp): the result stays within ~2.5e-7 of the correctly rounded sin/cos of the SAME
// fp32 argument the reference hands to torch.sin (mip.py:428), far inside the 2e-6 parity bound of the tests.
__device__ __forceinline__ void sincos_accurate(float y, float* sn, float* cs) {
  const double yd = (double)y;
  const double n = rint(yd * 0.63661977236758134308);
  double r = fma(-n, 1.57079632679489655800, yd);
  r = fma(-n, 6.12323399573676603587e-17, r);
  const float rf = (float)r;
  const int q = (int)n;
  const float z = rf * rf;
  const float ps = fmaf(rf * z, fmaf(z, fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), rf);
  const float pc = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f), -0.5f), 1.0f);
  const float s0 = (q & 1) ? pc : ps;
  const float c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
__device__ __forceinline__ float sin_accurate(float y) {
  float s, c;
  sincos_accurate(y, &s, &c);
  return s;
}
__device__ __forceinline__ float cos_accurate(float y) {
  float s, c;
  sincos_accurate(y, &s, &c);
  return c;
}

// Exact-phase variants for the gradient kernels: the phase of y = mean * 2^sh modulo 2 pi is a left shift of the 64-bit
// fixed-point fraction of mean / (2 pi) (one double multiply per mean instead of a double-precision reduction per
// feature); sin / cos of the reduced argument come from the same Cephes polynomials as sincos_accurate, so the
// accuracy is unchanged (~1 ulp), unlike the SFU path of the forward kernel.
__device__ __forceinline__ void phase_fixed(float mean, uint32_t* hi, uint32_t* lo) {
  double u = (double)mean * 0.15915494309189534561;  // turns
  u -= floor(u);
  const unsigned long long U = (unsigned long long)(u * 18446744073709551616.0);
  *hi = (uint32_t)(U >> 32), *lo = (uint32_t)U;
}
__device__ __forceinline__ void sincos_phase(uint32_t ph, float* sn, float* cs) {
  const uint32_t q = (ph + 0x20000000u) >> 30;       // nearest quarter turn
  const int f = (int)(ph - (q << 30));               // remainder in [-2^29, 2^29) units of 2^-32 turn
  const float rf = (float)f * 1.46291807926715968e-9f;  // * 2 pi / 2^32  -> [-pi/4, pi/4]
  const float z = rf * rf;
  const float ps = fmaf(rf * z, fmaf(z, fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), rf);
  const float pc = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f), -0.5f), 1.0f);
  const float s0 = (q & 1) ? pc : ps;
  const float c0 = (q & 1) ? ps : pc;
  *sn = (q & 2) ? -s0 : s0;
  *cs = ((q + 1) & 2) ? -c0 : c0;
}
// cos(y) and cos(fl32(y + fl32(pi/2))) for y = mean * 2^sh with phase `ph` (the second argument is what the reference
// differentiates, mip.py:428): fl32(y + pi/2) = y + pi/2 + eps with eps recovered exactly by a TwoSum.
__device__ __forceinline__ void cos_pair_phase(float y, uint32_t ph, float* cy, float* cz) {
  float sn, cs;
  sincos_phase(ph, &sn, &cs);
  const float z = y + kHalfPiF;
  const float bb = z - y;
  const float err = (y - (z - bb)) + (kHalfPiF - bb);
  const float eps = 4.37113900018624283e-8f - err;
  const float ce = fmaf(-0.5f * eps, eps, 1.0f);
  const float se = eps * fmaf(-0.16666667f * eps, eps, 1.0f);
  *cy = cs;
  *cz = -(sn * ce + cs * se);
}

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
      fs = e * sin_accurate(y);
      fc = e * sin_accurate(y + kHalfPiF);
    }
    enc[m * ld + j] = from_f32<T>(fs);
    enc[m * ld + F + j] = from_f32<T>(fc);
  }
}

// Fast path of the forward encoding (the default): HBM-bound instead of instruction-bound.
//   * one thread per (sample, xyz component) walks the 16 degrees: the phase of 2^l * mean modulo 2 pi is an exact
//     left shift of the 64-bit fixed-point fraction of mean / (2 pi) (one double multiply per component instead of
//     a range reduction per feature), sin / cos come from the SFU (MUFU, |err| < 5e-7 on [-pi, pi]);
//   * the reference's second half is sin(fl32(y + fl32(pi/2))), NOT cos(y): the fp32 addition rounds by up to
//     1.6e-2 rad at the highest degrees.  The rounding error is recovered exactly with a TwoSum and applied as a
//     small-angle rotation, cos(y + eps) = cos y cos eps - sin y sin eps, so the kernel matches mip.py:428 and not
//     the textbook formula;
//   * features are staged in shared memory and leave as full 16-byte rows (coalesced), instead of 2-byte scatters.
constexpr int kIpeTile = 64;  // samples per block iteration
template <typename T, int L, int MIN_DEG>  // MIN_DEG >= 0: compile-time lowest degree (every shift, scale and shared-memory
                                            // offset of the fully unrolled loop is an immediate); -1: run-time `min_deg`
__global__ void __launch_bounds__(3 * kIpeTile) ipe_fwd_tile_kernel(long long M, int min_deg,
                                                                     const float* __restrict__ means,
                                                                     const float* __restrict__ covs, T* __restrict__ enc,
                                                                     int ld) {
  constexpr int F = 6 * L;
  // Row pitch padded by 16 bytes: with the natural pitch (192 B / 384 B) the 32 threads of a warp - ~11 samples x 3
  // components writing the same feature column - fall into 2 (bf16) or 1 (fp32) banks, a 6- to 11-way conflict on
  // every one of the 96 stores per sample (ncu: 19.5 shared-memory wavefronts per sample for 3 warp-stores); the
  // padded pitch spreads 8 consecutive samples over 8 banks.
  constexpr int FP = F + 16 / (int)sizeof(T);
  __shared__ __align__(16) T tile[kIpeTile * FP];
  const int tid = threadIdx.x;
  const int s = tid / 3, c = tid - 3 * s;
  for (long long base = (long long)blockIdx.x * kIpeTile; base < M; base += (long long)gridDim.x * kIpeTile) {
    const long long idx = base * 3 + tid;
    if (idx < M * 3) {
      const float mean = means[idx], cov = covs[idx];
      double u = (double)mean * 0.15915494309189534561;  // turns
      u -= floor(u);
      const unsigned long long U = (unsigned long long)(u * 18446744073709551616.0);
      const uint32_t hi = (uint32_t)(U >> 32), lo = (uint32_t)U;
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int sh = (MIN_DEG >= 0 ? MIN_DEG : min_deg) + l;
        const uint32_t ph = __funnelshift_l(lo, hi, sh);  // top 32 bits of U << sh: phase in 2^-32 turns
        const float r = (float)(int)ph * 1.46291807926715968e-9f;  // 2 pi / 2^32 -> [-pi, pi)
        const float sn = __sinf(r), cs = __cosf(r);
        const float sc = __uint_as_float((uint32_t)(127 + sh) << 23);  // 2^sh
        const float y = mean * sc;
        // exp(-0.5 * (cov * 4^sh)): both scalings are by powers of two, i.e. exact, so the reference's fp32 argument is
        // x = -2^(2 sh - 1) cov exactly and exp(x) = 2^(fl32(x * log2 e)); folding the power of two into the constant
        // gives the same single rounding with one multiply (MUFU.EX2; results below 2^-126 flush to zero)
        const float kl = -1.44269504088896341f * __uint_as_float((uint32_t)(127 + 2 * sh - 1) << 23);
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(cov * kl));
        // fl32(y + pi/2) = y + pi/2 + eps exactly (TwoSum)
        const float z = y + kHalfPiF;
        const float bb = z - y;
        const float err = (y - (z - bb)) + (kHalfPiF - bb);
        const float eps = 4.37113900018624283e-8f - err;
        const float e2 = eps * eps;
        const float ce = fmaf(-0.5f, e2, 1.0f);
        const float se = eps * fmaf(-0.16666667f, e2, 1.0f);
        const float c2 = fmaf(cs, ce, -(sn * se));
        tile[s * FP + l * 3 + c] = from_f32<T>(e * sn);
        tile[s * FP + 3 * L + l * 3 + c] = from_f32<T>(e * c2);
      }
    }
    __syncthreads();
    constexpr int kVecPerRow = F * (int)sizeof(T) / 16;
    const int rows = (M - base) < kIpeTile ? (int)(M - base) : kIpeTile;
    for (int v = tid; v < rows * kVecPerRow; v += 3 * kIpeTile) {
      const int row = v / kVecPerRow, j = v - row * kVecPerRow;
      reinterpret_cast<uint4*>(enc + (base + row) * ld)[j] = reinterpret_cast<const uint4*>(tile + row * FP)[j];
    }
    __syncthreads();
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
    const bool fast = min_deg >= 0 && min_deg + L <= 31;
    uint32_t hi = 0, lo = 0;
    if (fast) phase_fixed(mean, &hi, &lo);
    for (int l = 0; l < L; ++l) {
      float sc = exp2f((float)(min_deg + l));
      float e = expf(-0.5f * (cov * (sc * sc)));
      if (e == 0.f) break;  // larger l only underflow harder
      float y = mean * sc;
      float gs = to_f32<T>(g[m * ld + 3 * l + c]), gc = to_f32<T>(g[m * ld + F + 3 * l + c]);
      float cy, cz;
      if (fast) cos_pair_phase(y, __funnelshift_l(lo, hi, min_deg + l), &cy, &cz);
      else cy = cos_accurate(y), cz = cos_accurate(y + kHalfPiF);
      acc += sc * (e * (gs * cy + gc * cz));
    }
    d_means[idx] = acc;
  }
}

// ---- stop_resample_grad = False (models/mip.py:336-350): gradients w.r.t. the Gaussians' variances and fence-posts ----
// sin(y), cos(y), sin(z), cos(z) for y = mean * 2^sh with phase `ph` and z = fl32(y + fl32(pi/2)) = y + pi/2 + eps
__device__ __forceinline__ void sincos_pair_phase(float y, uint32_t ph, float* sy, float* cy, float* sz, float* cz) {
  float sn, cs;
  sincos_phase(ph, &sn, &cs);
  const float z = y + kHalfPiF;
  const float bb = z - y;
  const float err = (y - (z - bb)) + (kHalfPiF - bb);
  const float eps = 4.37113900018624283e-8f - err;
  const float ce = fmaf(-0.5f * eps, eps, 1.0f);
  const float se = eps * fmaf(-0.16666667f * eps, eps, 1.0f);
  *sy = sn, *cy = cs;
  *sz = cs * ce - sn * se;     // sin(y + pi/2 + eps) = cos(y + eps)
  *cz = -(sn * ce + cs * se);  // cos(y + pi/2 + eps) = -sin(y + eps)
}

// One thread per (sample, component).  With e = exp(-0.5 cov 4^l), enc_s = e sin(y), enc_c = e sin(z):
//   d enc / d cov = -0.5 4^l enc                                    -> d_covs  = sum_l -0.5 4^l e (g_s sin y + g_c sin z)
// and, when `h` (= d raw_sigma / d enc from the Jacobian sweep) and `dv` (= dL/d v, v = J_ipe^T h the un-normalised
// density gradient) are given, the explicit dependence of v on the Gaussian (the ReLU network is piece-wise linear in
// enc, so h is locally constant and the only second derivatives are the encoding's own):
//   d v_c / d mean_c = sum_l -4^l e (h_s sin y + h_c sin z),   d v_c / d cov_c = sum_l -0.5 4^l 2^l e (h_s cos y + h_c cos z)
// d_means / d_covs: `accumulate` != 0 adds into the destination.
template <typename T>
__global__ void ipe_cov_hess_kernel(long long M, int min_deg, int L, const float* __restrict__ means,
                                    const float* __restrict__ covs, const T* __restrict__ g, int ldg,
                                    const float* __restrict__ h, int ldh, const float* __restrict__ dv,
                                    float* __restrict__ d_means, float* __restrict__ d_covs, int accumulate) {
  const int F = 3 * L;
  const long long total = M * 3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / 3;
    const int c = (int)(idx - m * 3);
    const float mean = means[idx], cov = covs[idx];
    uint32_t hi, lo;
    phase_fixed(mean, &hi, &lo);
    float acc_c = 0.f, acc_hm = 0.f, acc_hc = 0.f;
    for (int l = 0; l < L; ++l) {
      const float sc = exp2f((float)(min_deg + l)), sc2 = sc * sc;
      const float e = expf(-0.5f * (cov * sc2));
      if (e == 0.f) break;
      float sy, cy, sz, cz;
      sincos_pair_phase(mean * sc, __funnelshift_l(lo, hi, min_deg + l), &sy, &cy, &sz, &cz);
      if (g != nullptr) {
        const float gs = to_f32<T>(g[m * ldg + 3 * l + c]), gc = to_f32<T>(g[m * ldg + F + 3 * l + c]);
        acc_c += (-0.5f * sc2) * (e * (gs * sy + gc * sz));
      }
      if (h != nullptr) {
        const float hs = h[m * ldh + 3 * l + c], hc = h[m * ldh + F + 3 * l + c];
        acc_hm += -sc2 * (e * (hs * sy + hc * sz));
        acc_hc += (-0.5f * sc2 * sc) * (e * (hs * cy + hc * cz));
      }
    }
    float dm = 0.f, dc = acc_c;
    if (h != nullptr) {
      const float w = dv[idx];
      dm = w * acc_hm, dc += w * acc_hc;
    }
    if (d_means != nullptr) d_means[idx] = accumulate ? d_means[idx] + dm : dm;
    if (d_covs != nullptr) d_covs[idx] = accumulate ? d_covs[idx] + dc : dc;
  }
}

// Backward of cast_rays (models/mip.py:67-89, 36-58, 8-22) w.r.t. the fence-posts: one warp per ray.
//   mean_k = o_k + d_k t_mean,  cov_k = t_var d_k^2 + r_var (1 - d_k^2 / |d|^2)   with mu = (t0+t1)/2, hw = (t1-t0)/2
__global__ void __launch_bounds__(256)
cast_rays_bwd_kernel(long long R, int N, const float* __restrict__ t, const float* __restrict__ dirs,
                     const float* __restrict__ radii, const float* __restrict__ g_means,
                     const float* __restrict__ g_covs, float* __restrict__ d_t, int accumulate) {
  extern __shared__ float cb_smem[];
  const int lane = threadIdx.x & 31;
  float* s0 = cb_smem + (size_t)(threadIdx.x >> 5) * 2 * N;  // dL/dt0 of sample i
  float* s1 = s0 + N;                                        // dL/dt1 of sample i
  const long long warp0 = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  for (long long r = warp0; r < R; r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const float d[3] = {dirs[3 * r], dirs[3 * r + 1], dirs[3 * r + 2]};
    const float rad = radii[r];
    const RayGeom geom = ray_geom(d, d, rad);  // (origins do not enter the derivative)
    const float* tr = t + r * (N + 1);
    for (int i = lane; i < N; i += 32) {
      const float t0 = tr[i], t1 = tr[i + 1];
      const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
      const float mu2 = mu * mu, hw2 = hw * hw, hw3 = hw2 * hw, hw4 = hw2 * hw2;
      const float A = 3.f * mu2 + hw2, iA = 1.f / A, iA2 = iA * iA, iA3 = iA2 * iA;
      const float B = hw4 * (12.f * mu2 - hw2);
      const float tm_mu = 1.f + 2.f * hw2 * iA - 12.f * mu2 * hw2 * iA2;
      const float tm_hw = 4.f * mu * hw * iA - 4.f * mu * hw3 * iA2;
      const float k415 = (float)(4.0 / 15.0);
      const float tv_mu = -k415 * (24.f * mu * hw4 * iA2 - 12.f * B * mu * iA3);
      const float tv_hw = 2.f * hw * (float)(1.0 / 3.0) - k415 * ((48.f * mu2 * hw3 - 6.f * hw4 * hw) * iA2 - 4.f * B * hw * iA3);
      const float rv_mu = geom.rad2 * (mu / 2.f + k415 * (6.f * mu * hw4 * iA2));
      const float rv_hw = geom.rad2 * ((float)(5.0 / 6.0) * hw - k415 * (4.f * hw3 * iA - 2.f * hw4 * hw * iA2));
      const float* gm = g_means + 3 * (r * N + i);
      const float* gc = g_covs + 3 * (r * N + i);
      float G_tm = 0.f, G_tv = 0.f, G_rv = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (g_means != nullptr) G_tm += gm[k] * geom.d[k];
        if (g_covs != nullptr) G_tv += gc[k] * geom.dd[k], G_rv += gc[k] * geom.perp[k];
      }
      const float g_mu = G_tm * tm_mu + G_tv * tv_mu + G_rv * rv_mu;
      const float g_hw = G_tm * tm_hw + G_tv * tv_hw + G_rv * rv_hw;
      s0[i] = 0.5f * (g_mu - g_hw);
      s1[i] = 0.5f * (g_mu + g_hw);
    }
    __syncwarp();
    float* dt = d_t + r * (N + 1);
    for (int j = lane; j <= N; j += 32) {
      const float g = (j < N ? s0[j] : 0.f) + (j > 0 ? s1[j - 1] : 0.f);
      dt[j] = accumulate ? dt[j] + g : g;
    }
    __syncwarp();
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
      float cy, cz;
      if (min_deg >= 0 && min_deg + L <= 31) {
        uint32_t hi, lo;
        phase_fixed(means[3 * m + c], &hi, &lo);
        cos_pair_phase(y, __funnelshift_l(lo, hi, min_deg + l), &cy, &cz);
      } else {
        cy = cos_accurate(y), cz = cos_accurate(y + kHalfPiF);
      }
      os = w * cy;
      oc = w * cz;
    }
    out[m * ld + j] = from_f32<T>(os);
    out[m * ld + F + j] = from_f32<T>(oc);
  }
}

// Tiled variant of ipe_jvp_kernel (the default when L == 16 and rows are 16-byte aligned), built like
// ipe_fwd_tile_kernel: one thread per (sample, xyz component) walks the 16 degrees with the exact fixed-point phase,
// rows leave shared memory as 16-byte vectors (the per-feature kernel re-derives the phase 16 times per component and
// stores 2- / 4-byte pieces: 5.5 vs 1.2 ms per 16.8 M bf16 rows).  FAST = false keeps the arithmetic of the kernel above
// bit for bit (Cephes polynomials on the exact phase, expf); FAST = true - the bf16 product path, whose operands carry
// 3 decimal digits - takes sin / cos / exp2 from the SFU like the forward kernel.
template <int MIN_DEG_UNUSED, bool FAST>
__device__ __forceinline__ void ipe_degree_terms(float mean, float cov, uint32_t hi, uint32_t lo, int sh, float* sc_out,
                                                 float* e_out, float* cy, float* cz) {
  const uint32_t ph = __funnelshift_l(lo, hi, sh);
  const float sc = __uint_as_float((uint32_t)(127 + sh) << 23);  // 2^sh
  const float y = mean * sc;
  *sc_out = sc;
  if (FAST) {
    const float kl = -1.44269504088896341f * __uint_as_float((uint32_t)(127 + 2 * sh - 1) << 23);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(cov * kl));
    *e_out = e;
    const float r = (float)(int)ph * 1.46291807926715968e-9f;
    const float sn = __sinf(r), cs = __cosf(r);
    const float z = y + kHalfPiF;
    const float bb = z - y;
    const float err = (y - (z - bb)) + (kHalfPiF - bb);
    const float eps = 4.37113900018624283e-8f - err;
    const float e2 = eps * eps;
    const float ce = fmaf(-0.5f, e2, 1.0f);
    const float se = eps * fmaf(-0.16666667f, e2, 1.0f);
    *cy = cs;
    *cz = -(sn * ce + cs * se);
  } else {
    *e_out = expf(-0.5f * (cov * (sc * sc)));
    cos_pair_phase(y, ph, cy, cz);
  }
}

template <typename T, int L, bool FAST>
__global__ void __launch_bounds__(3 * kIpeTile) ipe_jvp_tile_kernel(long long M, int min_deg,
                                                                     const float* __restrict__ means,
                                                                     const float* __restrict__ covs,
                                                                     const float* __restrict__ v, T* __restrict__ out,
                                                                     int ld) {
  constexpr int F = 6 * L;
  constexpr int FP = F + 16 / (int)sizeof(T);
  constexpr int kVecPerRow = F * (int)sizeof(T) / 16;
  __shared__ __align__(16) T tile[kIpeTile * FP];
  const int tid = threadIdx.x;
  const int s = tid / 3, c = tid - 3 * s;
  for (long long base = (long long)blockIdx.x * kIpeTile; base < M; base += (long long)gridDim.x * kIpeTile) {
    const long long idx = base * 3 + tid;
    if (idx < M * 3) {
      const float mean = means[idx], cov = covs[idx], vv = v[idx];
      uint32_t hi, lo;
      phase_fixed(mean, &hi, &lo);
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float sc, e, cy, cz;
        ipe_degree_terms<0, FAST>(mean, cov, hi, lo, min_deg + l, &sc, &e, &cy, &cz);
        float os = 0.f, oc = 0.f;
        if (e != 0.f) {
          const float w = vv * sc * e;
          os = w * cy, oc = w * cz;
        }
        tile[s * FP + 3 * l + c] = from_f32<T>(os);
        tile[s * FP + 3 * L + 3 * l + c] = from_f32<T>(oc);
      }
    }
    __syncthreads();
    const int rows = (M - base) < kIpeTile ? (int)(M - base) : kIpeTile;
    for (int q = tid; q < rows * kVecPerRow; q += 3 * kIpeTile) {
      const int row = q / kVecPerRow, j = q - row * kVecPerRow;
      reinterpret_cast<uint4*>(out + (base + row) * ld)[j] = reinterpret_cast<const uint4*>(tile + row * FP)[j];
    }
    __syncthreads();
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

static int launch_sample_cast(int R, int N, const float* origins, int o_div, const float* directions,
                              const float* radii, const float* near_v, const float* far_v, int d_mod, int dir_mod,
                              const float* s_lin, const float* t_rand, int rand_ld, int disparity, float* t_out,
                              float* means, float* covs, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && o_div >= 1 && d_mod >= 0, "sample_cast: bad sizes");
  if (R == 0) return 0;
  const size_t smem = (size_t)8 * 6 * N * sizeof(float);
  const bool stage = N % 4 == 0 && smem <= 48 * 1024 && ((uintptr_t)means % 16 == 0) && ((uintptr_t)covs % 16 == 0);
  const int grid = grid_for((long long)R * 32, 256, 8);
  const bool plain = o_div == 1 && d_mod == 0 && dir_mod == 0 && !disparity;
#define PNB_LAUNCH_SC(STAGE_, MODE_)                                                                                 \
  sample_cast_kernel<STAGE_, MODE_><<<grid, 256, (STAGE_) ? smem : 0, as_stream(stream)>>>(                           \
      R, N, origins, o_div, directions, radii, near_v, far_v, d_mod, dir_mod, s_lin, t_rand, rand_ld, disparity, t_out, \
      means, covs)
  if (N <= 16 && (long long)R * N < (1ll << 31)) {
    const unsigned total = (unsigned)R * (unsigned)N;
    sample_cast_short_kernel<<<grid_for((long long)total, 256, 8), 256, 0, as_stream(stream)>>>(
        total, N, origins, o_div, directions, radii, near_v, far_v, d_mod, dir_mod, s_lin, t_rand, rand_ld, disparity,
        t_out, means, covs);
    return finish("sample_cast");
  }
  if (stage) {
    if (plain && t_rand == nullptr) PNB_LAUNCH_SC(true, 1);
    else if (plain) PNB_LAUNCH_SC(true, 2);
    else PNB_LAUNCH_SC(true, 0);
  } else {
    PNB_LAUNCH_SC(false, 0);
  }
#undef PNB_LAUNCH_SC
  return finish("sample_cast");
}

extern "C" int pnb_sample_cast(int R, int N, const float* origins, int o_div, const float* directions,
                               const float* radii, const float* near_v, const float* far_v, int d_mod,
                               const float* s_lin, const float* t_rand, int rand_ld, int disparity, float* t_out,
                               float* means, float* covs, void* stream) {
  return launch_sample_cast(R, N, origins, o_div, directions, radii, near_v, far_v, d_mod, d_mod, s_lin, t_rand,
                            rand_ld, disparity, t_out, means, covs, stream);
}

extern "C" int pnb_sample_cast_hemisp(int R, int N, const float* origins, int o_div, const float* directions,
                                      const float* radii, const float* near_v, const float* far_v, int d_mod,
                                      const float* s_lin, const float* t_rand, int rand_ld, float* t_out, float* means,
                                      float* covs, void* stream) {
  return launch_sample_cast(R, N, origins, o_div, directions, radii, near_v, far_v, d_mod, 0, s_lin, t_rand, rand_ld, 0,
                            t_out, means, covs, stream);
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
  const int esz = dtype == PNB_BF16 ? 2 : 4;
  if (L == 16 && min_deg >= 0 && min_deg + L <= 31 && ((size_t)ld * esz) % 16 == 0 && ((uintptr_t)enc % 16) == 0 &&
      getenv("PNB_IPE_SLOW") == nullptr) {
    long long tiles = ((long long)M + pnb::kIpeTile - 1) / pnb::kIpeTile;
    long long cap = (long long)pnb::kNumSMs * 16;  // (measured: 8 -> 16 CTAs per SM queued = +4 % fp32, +9 % bf16 rows)
    int g = (int)(tiles < cap ? tiles : cap);
    cudaStream_t st = as_stream(stream);
    if (dtype == PNB_BF16) {
      if (min_deg == 0)
        pnb::ipe_fwd_tile_kernel<__nv_bfloat16, 16, 0><<<g, 3 * pnb::kIpeTile, 0, st>>>(M, 0, means, covs,
                                                                                       (__nv_bfloat16*)enc, ld);
      else
        pnb::ipe_fwd_tile_kernel<__nv_bfloat16, 16, -1><<<g, 3 * pnb::kIpeTile, 0, st>>>(M, min_deg, means, covs,
                                                                                        (__nv_bfloat16*)enc, ld);
    } else {
      if (min_deg == 0)
        pnb::ipe_fwd_tile_kernel<float, 16, 0><<<g, 3 * pnb::kIpeTile, 0, st>>>(M, 0, means, covs, (float*)enc, ld);
      else
        pnb::ipe_fwd_tile_kernel<float, 16, -1><<<g, 3 * pnb::kIpeTile, 0, st>>>(M, min_deg, means, covs,
                                                                                (float*)enc, ld);
    }
    return finish("ipe_fwd");
  }
  int grid = grid_for((long long)M * 3 * L, 256);
  if (dtype == PNB_BF16)
    ipe_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs,
                                                                       (__nv_bfloat16*)enc, ld);
  else
    ipe_fwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(M, min_deg, L, means, covs, (float*)enc, ld);
  return finish("ipe_fwd");
}

static bool ipe_tile_ok(int L, int min_deg, const void* rows, int ld, int dtype) {
  const int esz = dtype == PNB_BF16 ? 2 : 4;
  return L == 16 && min_deg >= 0 && min_deg + L <= 31 && ((size_t)ld * esz) % 16 == 0 && ((uintptr_t)rows % 16) == 0 &&
         getenv("PNB_IPE_SLOW") == nullptr;
}
static int ipe_tile_grid(long long M) {
  long long tiles = (M + pnb::kIpeTile - 1) / pnb::kIpeTile;
  long long cap = (long long)pnb::kNumSMs * 16;
  return (int)(tiles < cap ? tiles : cap);
}

extern "C" int pnb_ipe_vjp(int M, const float* means, const float* covs, int min_deg, int max_deg, const void* d_enc,
                           int ld, int dtype, float* d_means, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && ld >= 6 * L, "ipe_vjp: bad sizes");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  // (a tiled variant like ipe_jvp_tile_kernel was measured slower here: 1.43 vs 1.25 ms per 16.8 M bf16 rows)
  int grid = grid_for((long long)M * 3, 256);
  if (dtype == PNB_BF16)
    ipe_vjp_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, (const __nv_bfloat16*)d_enc, ld,
                                                        d_means);
  else
    ipe_vjp_kernel<float><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, (const float*)d_enc, ld, d_means);
  return finish("ipe_vjp");
}

extern "C" int pnb_ipe_cov_hess(int M, const float* means, const float* covs, int min_deg, int max_deg,
                                const void* d_enc, int ld, int dtype, const float* h_enc, int ldh, const float* d_v,
                                float* d_means, float* d_covs, int accumulate, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && min_deg >= 0 && max_deg <= 31, "ipe_cov_hess: degrees must lie in [0, 31]");
  PNB_REQUIRE((d_enc == nullptr || ld >= 6 * L) && (h_enc == nullptr || (ldh >= 6 * L && d_v != nullptr)),
              "ipe_cov_hess: bad row strides / missing d_v");
  PNB_REQUIRE(d_enc != nullptr || h_enc != nullptr, "ipe_cov_hess: nothing to do");
  if (M == 0) return 0;
  int grid = grid_for((long long)M * 3, 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == PNB_BF16)
    ipe_cov_hess_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, (const __nv_bfloat16*)d_enc, ld,
                                                             h_enc, ldh, d_v, d_means, d_covs, accumulate);
  else
    ipe_cov_hess_kernel<float><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, (const float*)d_enc, ld, h_enc, ldh,
                                                     d_v, d_means, d_covs, accumulate);
  return finish("ipe_cov_hess");
}

extern "C" int pnb_cast_rays_bwd(int R, int N, const float* t, const float* directions, const float* radii,
                                 const float* g_means, const float* g_covs, float* d_t, int accumulate, void* stream) {
  PNB_REQUIRE(R >= 0 && N > 0 && (R == 0 || (t && directions && radii && d_t && (g_means || g_covs))),
              "cast_rays_bwd: bad arguments");
  if (R == 0) return 0;
  const size_t smem = (size_t)8 * 2 * N * sizeof(float);
  PNB_REQUIRE(smem <= 200 * 1024, "cast_rays_bwd: N too large");
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(cast_rays_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = grid_for((long long)R * 32, 256, 4);
  cast_rays_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(R, N, t, directions, radii, g_means, g_covs, d_t,
                                                              accumulate);
  return finish("cast_rays_bwd");
}

extern "C" int pnb_ipe_jvp(int M, const float* means, const float* covs, int min_deg, int max_deg, const float* v,
                           void* out, int ld, int dtype, void* stream) {
  int L = max_deg - min_deg;
  PNB_REQUIRE(M >= 0 && L > 0 && ld >= 6 * L, "ipe_jvp: bad sizes");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (ipe_tile_ok(L, min_deg, out, ld, dtype)) {
    // bf16 rows (the tensor-core path's operand) take the SFU variant, fp32 rows (parity mode) the exact one
    const int g = ipe_tile_grid(M), th = 3 * pnb::kIpeTile;
    if (dtype == PNB_BF16)
      pnb::ipe_jvp_tile_kernel<__nv_bfloat16, 16, true><<<g, th, 0, st>>>(M, min_deg, means, covs, v, (__nv_bfloat16*)out, ld);
    else
      pnb::ipe_jvp_tile_kernel<float, 16, false><<<g, th, 0, st>>>(M, min_deg, means, covs, v, (float*)out, ld);
    return finish("ipe_jvp");
  }
  int grid = grid_for((long long)M * 3 * L, 256);
  if (dtype == PNB_BF16)
    ipe_jvp_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, v, (__nv_bfloat16*)out, ld);
  else
    ipe_jvp_kernel<float><<<grid, 256, 0, st>>>(M, min_deg, L, means, covs, v, (float*)out, ld);
  return finish("ipe_jvp");
}

extern "C" int pnb_pos_enc(int R, const float* x, int deg, float* out, void* stream) {
  PNB_REQUIRE(R >= 0 && deg >= 0, "pos_enc: bad sizes");
  if (R == 0) return 0;
  pos_enc_kernel<<<grid_for((long long)R * (3 + 6 * deg), 256), 256, 0, as_stream(stream)>>>(R, deg, x, out);
  return finish("pos_enc");
}
