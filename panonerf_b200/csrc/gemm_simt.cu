// fp32 "parity mode" GEMMs of the MLP (forward, dgrad, wgrad) on the CUDA cores (FFMA).  This path exists so
// that the whole pipeline can be checked against the fp32 reference at <=1e-5 relative; the throughput path is
// gemm_tc.cu (tcgen05).  64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
#include "common.cuh"

namespace pnb {

constexpr int BM = 64, BN = 64, BK = 16;

// element (m,k) of op(A) and (k,n) of op(B) for the three modes of pnb_gemm_f32
template <int MODE>
__device__ __forceinline__ long long a_off(long long m, long long k, int lda) {
  return MODE == 2 ? k * lda + m : m * lda + k;
}
template <int MODE>
__device__ __forceinline__ long long b_off(long long k, long long n, int ldb) {
  return MODE == 0 ? n * ldb + k : k * ldb + n;
}

template <int MODE>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(long long M, int N, long long K, long long k_chunk, const float* __restrict__ A, int lda,
                const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc, const float* __restrict__ bias,
                const float* __restrict__ row_bias, int row_group, const float* __restrict__ mask_src, int ld_mask,
                int flags) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long k_begin = (long long)blockIdx.z * k_chunk;
  const long long k_end = k_begin + k_chunk < K ? k_begin + k_chunk : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
    // A tile: modes 0/1 are contiguous along k, mode 2 along m
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int kk, mm;
      if (MODE == 2) {
        mm = tid & 63, kk = (tid >> 6) + 4 * p;
      } else {
        kk = tid & 15, mm = (tid >> 4) + 16 * p;
      }
      long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_end) ? A[a_off<MODE>(m, k, lda)] : 0.f;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int kk, nn;
      if (MODE == 0) {
        kk = tid & 15, nn = (tid >> 4) + 16 * p;
      } else {
        nn = tid & 63, kk = (tid >> 6) + 4 * p;
      }
      long long n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < k_end) ? B[b_off<MODE>(k, n, ldb)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (MODE == 2) {
        atomicAdd(C + m * ldc + n, v);  // split over the sample axis; C is an accumulating gradient buffer
      } else {
        if (flags & PNB_EPI_BIAS) v += bias[n];
        if (row_bias) v += row_bias[(m / row_group) * N + n];
        if (flags & PNB_EPI_ACCUM) v += C[m * ldc + n];
        if (flags & PNB_EPI_RELU) v = fmaxf(v, 0.f);
        if (flags & PNB_EPI_MASK) v = mask_src[m * ld_mask + n] > 0.f ? v : 0.f;
        C[m * ldc + n] = v;
      }
    }
  }
}

}  // namespace pnb

using namespace pnb;

extern "C" int pnb_gemm_f32(int mode, long long M, int N, long long K, const float* A, int lda, const float* B, int ldb,
                            float* C, int ldc, const float* bias, const float* row_bias, int row_group,
                            const float* mask_src, int ld_mask, int flags, void* stream) {
  PNB_REQUIRE(mode >= 0 && mode <= 2 && M >= 0 && N > 0 && K >= 0, "gemm_f32: bad arguments");
  PNB_REQUIRE(!(flags & PNB_EPI_BIAS) || bias != nullptr, "gemm_f32: bias flag without bias");
  PNB_REQUIRE(!(flags & PNB_EPI_MASK) || mask_src != nullptr, "gemm_f32: mask flag without mask source");
  PNB_REQUIRE(row_bias == nullptr || row_group > 0, "gemm_f32: row_bias needs row_group");
  if (M == 0 || K == 0) return 0;
  long long gm = (M + BM - 1) / BM;
  PNB_REQUIRE(gm < (1ll << 31), "gemm_f32: M too large");
  dim3 grid((unsigned)gm, (unsigned)((N + BN - 1) / BN), 1);
  long long k_chunk = K;
  if (mode == 2) {
    // spread the sample axis over ~4 waves of CTAs
    long long tiles = gm * grid.y;
    long long splits = (4ll * kNumSMs + tiles - 1) / tiles;
    long long max_splits = (K + 4 * BK - 1) / (4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    grid.z = (unsigned)((K + k_chunk - 1) / k_chunk);
  }
  cudaStream_t st = as_stream(stream);
  if (mode == 0)
    gemm_f32_kernel<0><<<grid, 256, 0, st>>>(M, N, K, k_chunk, A, lda, B, ldb, C, ldc, bias, row_bias, row_group, mask_src,
                                             ld_mask, flags);
  else if (mode == 1)
    gemm_f32_kernel<1><<<grid, 256, 0, st>>>(M, N, K, k_chunk, A, lda, B, ldb, C, ldc, bias, row_bias, row_group, mask_src,
                                             ld_mask, flags);
  else
    gemm_f32_kernel<2><<<grid, 256, 0, st>>>(M, N, K, k_chunk, A, lda, B, ldb, C, ldc, bias, row_bias, row_group, mask_src,
                                             ld_mask, flags);
  return finish("gemm_f32");
}
