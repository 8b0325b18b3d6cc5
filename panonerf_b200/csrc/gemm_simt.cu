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

template <int MODE, typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(long long M, int N, long long K, long long k_chunk, const TIn* __restrict__ A, int lda,
                const TIn* __restrict__ B, int ldb, TOut* __restrict__ C, int ldc, const float* __restrict__ bias,
                const float* __restrict__ row_bias, int row_group, const TIn* __restrict__ mask_src, int ld_mask,
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
      As[kk][mm] = (m < M && k < k_end) ? to_f32<TIn>(A[a_off<MODE>(m, k, lda)]) : 0.f;
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
      Bs[kk][nn] = (n < N && k < k_end) ? to_f32<TIn>(B[b_off<MODE>(k, n, ldb)]) : 0.f;
    }
    __syncthreads();
    // two-level summation: a 16-term partial per k-block, then one add into the running total (error grows with
    // K/16 + 16 instead of K terms - keeps the ill-conditioned Jacobian chain as accurate as a blocked CPU GEMM)
    float part[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
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
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
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
        // split over the sample axis; C is an accumulating fp32 gradient buffer
        atomicAdd(reinterpret_cast<float*>(C) + m * ldc + n, v);
      } else {
        if (flags & PNB_EPI_BIAS) v += bias[n];
        if (row_bias) v += row_bias[(m / row_group) * N + n];
        if (flags & PNB_EPI_ACCUM) v += to_f32<TOut>(C[m * ldc + n]);
        if (flags & PNB_EPI_RELU) v = fmaxf(v, 0.f);
        if (flags & PNB_EPI_MASK) v = to_f32<TIn>(mask_src[m * ld_mask + n]) > 0.f ? v : 0.f;
        C[m * ldc + n] = from_f32<TOut>(v);
      }
    }
  }
}

}  // namespace pnb

using namespace pnb;

template <typename TIn, typename TOut>
static int launch_gemm(int mode, long long M, int N, long long K, const void* A, int lda, const void* B, int ldb, void* C,
                       int ldc, const float* bias, const float* row_bias, int row_group, const void* mask_src,
                       int ld_mask, int flags, void* stream) {
  long long gm = (M + BM - 1) / BM;
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
  const TIn* a = (const TIn*)A;
  const TIn* b = (const TIn*)B;
  const TIn* ms = (const TIn*)mask_src;
  TOut* c = (TOut*)C;
  if (mode == 0)
    gemm_f32_kernel<0, TIn, TOut><<<grid, 256, 0, st>>>(M, N, K, k_chunk, a, lda, b, ldb, c, ldc, bias, row_bias, row_group,
                                                       ms, ld_mask, flags);
  else if (mode == 1)
    gemm_f32_kernel<1, TIn, TOut><<<grid, 256, 0, st>>>(M, N, K, k_chunk, a, lda, b, ldb, c, ldc, bias, row_bias, row_group,
                                                       ms, ld_mask, flags);
  else
    gemm_f32_kernel<2, TIn, TOut><<<grid, 256, 0, st>>>(M, N, K, k_chunk, a, lda, b, ldb, c, ldc, bias, row_bias, row_group,
                                                       ms, ld_mask, flags);
  return finish("gemm_simt");
}

static int check_gemm_args(int mode, long long M, int N, long long K, const float* bias, const float* row_bias,
                           int row_group, const void* mask_src, int flags) {
  PNB_REQUIRE(mode >= 0 && mode <= 2 && M >= 0 && N > 0 && K >= 0, "gemm: bad arguments");
  PNB_REQUIRE(M == 0 || !(flags & PNB_EPI_BIAS) || bias != nullptr, "gemm: bias flag without bias");
  PNB_REQUIRE(M == 0 || !(flags & PNB_EPI_MASK) || mask_src != nullptr, "gemm: mask flag without mask source");
  PNB_REQUIRE(row_bias == nullptr || row_group > 0, "gemm: row_bias needs row_group");
  PNB_REQUIRE((M + BM - 1) / BM < (1ll << 31), "gemm: M too large");
  return 0;
}

extern "C" int pnb_gemm_f32(int mode, long long M, int N, long long K, const float* A, int lda, const float* B, int ldb,
                            float* C, int ldc, const float* bias, const float* row_bias, int row_group,
                            const float* mask_src, int ld_mask, int flags, void* stream) {
  int rc = check_gemm_args(mode, M, N, K, bias, row_bias, row_group, mask_src, flags);
  if (rc) return rc;
  if (M == 0 || K == 0) return 0;
  return launch_gemm<float, float>(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, row_bias, row_group, mask_src, ld_mask,
                                   flags, stream);
}

// Same GEMMs on bf16-stored operands (fp32 FFMA arithmetic): the CUDA-core twin of the tcgen05 path, used by the
// tests to check the tensor-core kernels on identical bf16 data (only the accumulation order differs).
extern "C" int pnb_gemm_bf16_simt(int mode, long long M, int N, long long K, const void* A, int lda, const void* B,
                                  int ldb, void* C, int ldc, int c_dtype, const float* bias, const float* row_bias,
                                  int row_group, const void* mask_src, int ld_mask, int flags, void* stream) {
  int rc = check_gemm_args(mode, M, N, K, bias, row_bias, row_group, mask_src, flags);
  if (rc) return rc;
  PNB_REQUIRE(mode != 2 || c_dtype == PNB_F32, "gemm_bf16_simt: wgrad accumulates into fp32");
  if (M == 0 || K == 0) return 0;
  if (c_dtype == PNB_BF16)
    return launch_gemm<__nv_bfloat16, __nv_bfloat16>(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, row_bias, row_group,
                                                     mask_src, ld_mask, flags, stream);
  return launch_gemm<__nv_bfloat16, float>(mode, M, N, K, A, lda, B, ldb, C, ldc, bias, row_bias, row_group, mask_src,
                                           ld_mask, flags, stream);
}
