// K3/K4/K5: the MLP GEMMs on the 5th-generation tensor cores (sm_100a only).
//
//   pnb_linear_tc : C[M,Nout] = epi(A[M,K] * W[Nout,K]^T)   bf16 operands, fp32 accumulation in TMEM
//   pnb_wgrad_tc  : dW[Nw,Kw] += dZ[M,Nw]^T * X[M,Kw]        reduction over the M samples
//
// Design (one persistent CTA per SM, warp-specialised, no CUTLASS):
//   warp 0      TMA producer  : cp.async.bulk.tensor.2d tiles, 128B-swizzled, completing on mbarriers
//   warp 1      MMA issuer    : one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16)
//                               straight from shared memory; accumulators live in TMEM (double-buffered)
//   warps 2..5  epilogue      : tcgen05.ld 32x32b -> bias / ReLU / mask -> bf16|fp32 -> global
// linear: the weight matrix is loaded ONCE per CTA and stays resident in shared memory (<=160 KB), only the
// activation tiles (128 rows x 64 bf16) stream through a 3-4 stage ring, so HBM traffic is the algorithmic
// minimum: read A once, write C once.  wgrad: both operands are "MN-major" (the reduction axis is the slow axis
// in memory), which tcgen05 consumes directly through MN-major shared-memory descriptors - no transposes.
#include "tc_common.cuh"

namespace pnb {
namespace tc {

// ---------------------------------------------------------------------------------------------------------------
// linear: C = epi(A * W^T)
// ---------------------------------------------------------------------------------------------------------------
struct LinearParams {
  long long M;
  int Nout, K, num_kb, stages, tail_k16;
  void* C;
  int ldc, c_dtype;
  const float* bias;
  const float* row_bias;
  int row_group;
  const __nv_bfloat16* mask_src;
  int ld_mask, flags;
  long long num_m_tiles;
  float* colsum;  // nullable: += column sums of the bf16 output (fast path only) = bias gradient of the producer
};

template <int BLOCK_N, typename OutT>
__device__ __forceinline__ void store_row_chunk(const LinearParams& p, long long m, int n_base, float* v, int count) {
  // v[0..count) are columns n_base.. of row m.  Epilogue order (same as gemm_simt.cu):
  //   + bias, + row_bias, + previous C (ACCUM), ReLU, mask, store.
  OutT* crow = reinterpret_cast<OutT*>(p.C) + m * p.ldc;
  const bool full = n_base + count <= p.Nout;
  const bool vec_c = full && (p.ldc % 8 == 0) && (n_base % 8 == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i >= count) break;
    int n = n_base + i;
    if (n < p.Nout) {
      if (p.flags & PNB_EPI_BIAS) v[i] += p.bias[n];
      if (p.row_bias) v[i] += p.row_bias[(m / p.row_group) * p.Nout + n];
    }
  }
  if (p.flags & PNB_EPI_ACCUM) {
    if (sizeof(OutT) == 2 && vec_c) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        if (i >= count) break;
        uint4 prev = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(crow) + n_base + i);
        const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&prev);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i + j] += __bfloat162float(ph[j]);
      }
    } else if (sizeof(OutT) == 4 && vec_c) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (i >= count) break;
        float4 prev = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(crow) + n_base + i);
        v[i] += prev.x, v[i + 1] += prev.y, v[i + 2] += prev.z, v[i + 3] += prev.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= count) break;
        if (n_base + i < p.Nout) v[i] += to_f32<OutT>(crow[n_base + i]);
      }
    }
  }
  if (p.flags & PNB_EPI_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i >= count) break;
      v[i] = fmaxf(v[i], 0.f);
    }
  }
  if ((p.flags & PNB_EPI_MASK) != 0) {
    const __nv_bfloat16* mrow = p.mask_src + m * p.ld_mask;
    if (full && (p.ld_mask % 8 == 0) && (n_base % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.mask_src) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        if (i >= count) break;
        uint4 raw = *reinterpret_cast<const uint4*>(mrow + n_base + i);
        const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(__bfloat162float(h[j]) > 0.f)) v[i + j] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= count) break;
        if (n_base + i < p.Nout && !(__bfloat162float(mrow[n_base + i]) > 0.f)) v[i] = 0.f;
      }
    }
  }
  if (sizeof(OutT) == 2) {
    __nv_bfloat16* brow = reinterpret_cast<__nv_bfloat16*>(crow);
    if (vec_c) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        if (i >= count) break;
        __nv_bfloat162 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
        *reinterpret_cast<uint4*>(brow + n_base + i) = *reinterpret_cast<uint4*>(h);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= count) break;
        if (n_base + i < p.Nout) brow[n_base + i] = __float2bfloat16_rn(v[i]);
      }
    }
  } else {
    float* frow = reinterpret_cast<float*>(crow);
    if (vec_c) {  // ldc % 8 == 0 implies 16-byte aligned float4 rows
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (i >= count) break;
        *reinterpret_cast<float4*>(frow + n_base + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i >= count) break;
        if (n_base + i < p.Nout) frow[n_base + i] = v[i];
      }
    }
  }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  constexpr int kWTileBytes = BLOCK_N * kBlockK * 2;  // one k-block of the resident weight
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + (size_t)p.num_kb * kWTileBytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem_a + (size_t)p.stages * kATileBytes);
  constexpr uint32_t kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;  // power of two for every BLOCK_N we use

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars->w_full, (uint32_t)(p.num_kb * kWTileBytes));
      for (int kb = 0; kb < p.num_kb; ++kb)
        tma_load_2d(smem_w + (size_t)kb * kWTileBytes, &tmW, &bars->w_full, kb * kBlockK, n0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_expect_tx(&bars->full[stage], kATileBytes);
          tma_load_2d(smem_a + (size_t)stage * kATileBytes, &tmA, &bars->full[stage], kb * kBlockK,
                      (int)(tile * kBlockM));
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(kBlockM, BLOCK_N, 0, 0);
      mbar_wait(&bars->w_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = (uint32_t)(it & 1);
        mbar_wait(&bars->tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const int nk16 = (kb == p.num_kb - 1) ? p.tail_k16 : kBlockK / 16;
          const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kATileBytes);
          const uint32_t b_addr = smem_u32(smem_w + (size_t)kb * kWTileBytes);
          for (int k = 0; k < nk16; ++k) {
            // K-major, 128B swizzle: 8-row groups are 1024 B apart; a K=16 slice is 32 B further along the row
            uint64_t ad = smem_desc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t bd = smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
        umma_commit(&bars->tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    long long it = 0;
    for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = (uint32_t)(it & 1);
      mbar_wait(&bars->tmem_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const long long m = tile * kBlockM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BLOCK_N;
      constexpr int kChunk = BLOCK_N >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += kChunk) {
        float v[32];
        if (kChunk == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
        if (m < p.M && n0 + c0 < p.Nout) {
          if (p.c_dtype == PNB_BF16)
            store_row_chunk<BLOCK_N, __nv_bfloat16>(p, m, n0 + c0, v, kChunk);
          else
            store_row_chunk<BLOCK_N, float>(p, m, n0 + c0, v, kChunk);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// linear, fast path: bf16 output, epilogue flags known at compile time.
//   * 8 epilogue warps (two per TMEM lane quadrant, alternating 64-column groups) hide instruction latency;
//   * the epilogue is straight-line code (no per-element flag branches), bias comes from shared memory;
//   * each warp stages its 32x64 bf16 sub-tile in 128B-swizzled shared memory and one lane issues a TMA store
//     (cp.async.bulk.tensor, bulk-group tracked, double-buffered) -> fully coalesced 128-byte global writes.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kFastThreads = 64 + 8 * 32;
constexpr int kStoreBoxBytes = 32 * 64 * 2;  // 32 rows x 64 bf16

#define PNB_EPI_ROWBIAS 16  // compile-time only: row_bias pointer is present


template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kFastThreads, 1)
linear_tc_fast_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmC, LinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  constexpr int kWTileBytes = BLOCK_N * kBlockK * 2;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  constexpr int kGroups = BLOCK_N / 64;  // 64-column groups per tile
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem_w + (size_t)p.num_kb * kWTileBytes;
  uint8_t* smem_c = smem_a + (size_t)p.stages * kATileBytes;          // 8 warps x 4 KB staging for the TMA stores
  float* smem_bias = reinterpret_cast<float*>(smem_c + 8 * kStoreBoxBytes);
  Barriers* bars = reinterpret_cast<Barriers*>(smem_bias + BLOCK_N);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BLOCK_N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars->tmem_full[b], 1);
      mbar_init(&bars->tmem_empty[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, kTmemCols);
  if (EPI & PNB_EPI_BIAS) {
    for (int i = threadIdx.x; i < BLOCK_N; i += kFastThreads) smem_bias[i] = (n0 + i < p.Nout) ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&bars->w_full, (uint32_t)(p.num_kb * kWTileBytes));
      for (int kb = 0; kb < p.num_kb; ++kb)
        tma_load_2d(smem_w + (size_t)kb * kWTileBytes, &tmW, &bars->w_full, kb * kBlockK, n0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          mbar_expect_tx(&bars->full[stage], kATileBytes);
          tma_load_2d(smem_a + (size_t)stage * kATileBytes, &tmA, &bars->full[stage], kb * kBlockK,
                      (int)(tile * kBlockM));
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(kBlockM, BLOCK_N, 0, 0);
      mbar_wait(&bars->w_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = (uint32_t)(it & 1);
        mbar_wait(&bars->tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const int nk16 = (kb == p.num_kb - 1) ? p.tail_k16 : kBlockK / 16;
          const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kATileBytes);
          const uint32_t b_addr = smem_u32(smem_w + (size_t)kb * kWTileBytes);
          for (int k = 0; k < nk16; ++k) {
            uint64_t ad = smem_desc_sw128(a_addr + k * 32, 16, 1024);
            uint64_t bd = smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) stage = 0, phase ^= 1;
        }
        umma_commit(&bars->tmem_full[buf]);
      }
    }
    __syncwarp();
  } else {
    const int ew = warp - 2;       // 0..7
    const int q = warp & 3;        // TMEM lane quadrant (hardware rule: warp id % 4)
    const int half = ew >> 2;      // which of the two warps sharing this quadrant
    uint8_t* stage_c = smem_c + (size_t)ew * kStoreBoxBytes;
    const __nv_bfloat16* mask_base = p.mask_src;
    __nv_bfloat16* c_base = reinterpret_cast<__nv_bfloat16*>(p.C);
    float cs00 = 0.f, cs01 = 0.f, cs10 = 0.f, cs11 = 0.f;  // column-sum partials: [group slot][column pair]
    long long it = 0;
    for (long long tile = blockIdx.x; tile < p.num_m_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = (uint32_t)(it & 1);
      mbar_wait(&bars->tmem_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const long long m = tile * kBlockM + q * 32 + lane;
      const bool row_ok = m < p.M;
      const float* rb_row = nullptr;
      if (EPI & PNB_EPI_ROWBIAS) rb_row = p.row_bias + (row_ok ? (m / p.row_group) : 0) * (long long)p.Nout;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BLOCK_N;
#pragma unroll 1
      for (int g = half; g < kGroups; g += 2) {
        const int nb = n0 + g * 64;  // first global column of this group
        uint32_t r[64];
        tmem_ld64(taddr + g * 64, r);
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
        if (EPI & PNB_EPI_BIAS) {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            float4 b4 = *reinterpret_cast<const float4*>(smem_bias + g * 64 + i);
            v[i] += b4.x, v[i + 1] += b4.y, v[i + 2] += b4.z, v[i + 3] += b4.w;
          }
        }
        if (EPI & PNB_EPI_ROWBIAS) {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            float4 b4 = *reinterpret_cast<const float4*>(rb_row + nb + i);
            v[i] += b4.x, v[i + 1] += b4.y, v[i + 2] += b4.z, v[i + 3] += b4.w;
          }
        }
        if (EPI & PNB_EPI_ACCUM) {
          if (row_ok) {
            const uint4* prow = reinterpret_cast<const uint4*>(c_base + m * p.ldc + nb);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 raw = prow[i];
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[8 * i + j] += __bfloat162float(h[j]);
            }
          }
        }
        if (EPI & PNB_EPI_RELU) {
#pragma unroll
          for (int i = 0; i < 64; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (EPI & PNB_EPI_MASK) {
          if (row_ok) {
            const uint4* mrow = reinterpret_cast<const uint4*>(mask_base + m * p.ld_mask + nb);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint4 raw = mrow[i];
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (!(__bfloat162float(h[j]) > 0.f)) v[8 * i + j] = 0.f;
            }
          }
        }
        // the TMA store that last read this staging buffer must have finished reading it
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        uint8_t* dst = stage_c + lane * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[8 * i + 2 * j], v[8 * i + 2 * j + 1]);
          *reinterpret_cast<uint4*>(dst + ((i ^ (lane & 7)) << 4)) = *reinterpret_cast<uint4*>(h);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, stage_c, nb, (int)(tile * kBlockM + q * 32));
          bulk_commit();
        }
        if (p.colsum != nullptr) {
          // bias gradient of the layer that produced this tile: column sums of the bf16-rounded values just staged.
          // Lane l owns columns 2l, 2l+1 of the 64-column group; rows beyond M were computed from zero-filled
          // operands and are exactly zero unless a bias/accumulate epilogue is active (never combined with colsum).
          const int rows_valid = (int)((p.M - (tile * kBlockM + q * 32)) < 32 ? (p.M - (tile * kBlockM + q * 32)) : 32);
          float s0 = 0.f, s1 = 0.f;
          const int chunk = lane >> 2, within = (lane & 3) * 4;  // 16-byte chunk and byte offset inside it
          for (int rr = 0; rr < rows_valid; ++rr) {
            const uint8_t* src = stage_c + rr * 128 + ((chunk ^ (rr & 7)) << 4) + within;
            __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(src);
            s0 += __bfloat162float(h.x), s1 += __bfloat162float(h.y);
          }
          if (g < 2) {
            cs00 += s0, cs01 += s1;
          } else {
            cs10 += s0, cs11 += s1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
    }
    if (lane == 0) bulk_wait_all();
    if (p.colsum != nullptr) {
      if (half < kGroups) {
        atomicAdd(p.colsum + n0 + half * 64 + 2 * lane, cs00);
        atomicAdd(p.colsum + n0 + half * 64 + 2 * lane + 1, cs01);
      }
      if (half + 2 < kGroups) {
        atomicAdd(p.colsum + n0 + (half + 2) * 64 + 2 * lane, cs10);
        atomicAdd(p.colsum + n0 + (half + 2) * 64 + 2 * lane + 1, cs11);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: dW[Nw,Kw] += dZ^T X, both operands MN-major.  Each CTA reduces a contiguous slab of samples into TMEM
// (Nw/128 accumulators of 128 x Kw fp32), then writes its partial to a workspace; a second kernel sums the
// partials in a fixed order (deterministic) into the accumulating gradient buffer.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  long long M;
  int Nw, Kw, n_halves, kw_chunks, stages;
  long long rows_per_cta;
  float* partial;  // [gridDim.x, Nw, Kw]
};

constexpr int kWgRows = 64;                          // samples per pipeline stage
constexpr int kWgBoxBytes = kWgRows * kBlockK * 2;   // one 64x64 bf16 box = 8 KB

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmX, WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const int z_chunks = p.n_halves * 2;  // 64-wide column chunks of dZ
  const int stage_bytes = (z_chunks + p.kw_chunks) * kWgBoxBytes;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + (size_t)p.stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_begin = (long long)blockIdx.x * p.rows_per_cta;
  long long row_end = row_begin + p.rows_per_cta;
  if (row_end > p.M) row_end = p.M;
  const int num_it = row_end > row_begin ? (int)((row_end - row_begin + kWgRows - 1) / kWgRows) : 0;
  const int kw_pad = p.kw_chunks * 64;  // TMEM columns per accumulator
  const uint32_t tmem_cols = 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmZ);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->tmem_full[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        mbar_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        const int row = (int)(row_begin + (long long)it * kWgRows);
        // slabs are multiples of 64 rows, so only the global tail is partial and TMA zero-fills it
        for (int c = 0; c < z_chunks; ++c)
          tma_load_2d(base + (size_t)c * kWgBoxBytes, &tmZ, &bars->full[stage], c * 64, row);
        for (int c = 0; c < p.kw_chunks; ++c)
          tma_load_2d(base + (size_t)(z_chunks + c) * kWgBoxBytes, &tmX, &bars->full[stage], c * 64, row);
        if (++stage == p.stages) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc_bf16(128, p.Kw, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t x_addr = base + z_chunks * kWgBoxBytes;
#pragma unroll 1
        for (int k = 0; k < kWgRows / 16; ++k) {
          // MN-major, 128B swizzle: 64-element column chunks are one box (8 KB) apart (LBO), 8-sample groups are
          // 1024 B apart (SBO); a K=16 slice (16 samples) starts 2048 B further.
          uint64_t bd = smem_desc_sw128(x_addr + k * 2048, kWgBoxBytes, 1024);
          for (int h = 0; h < p.n_halves; ++h) {
            uint64_t ad = smem_desc_sw128(base + h * 2 * kWgBoxBytes + k * 2048, kWgBoxBytes, 1024);
            umma_f16(tmem_base + h * kw_pad, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
        }
        umma_commit(&bars->empty[stage]);
        if (++stage == p.stages) stage = 0, phase ^= 1;
      }
      umma_commit(&bars->tmem_full[0]);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(&bars->tmem_full[0], 0);
    tc_fence_after();
    float* out = p.partial + (size_t)blockIdx.x * p.Nw * p.Kw;
    for (int h = 0; h < p.n_halves; ++h) {
      const int n = h * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * kw_pad;
#pragma unroll 1
      for (int c0 = 0; c0 < p.Kw; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (num_it == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (n < p.Nw) {
          float* orow = out + (size_t)n * p.Kw + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c0 + i < p.Kw) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void wgrad_reduce_kernel(int parts, int Nw, int Kw, const float* __restrict__ partial,
                                    float* __restrict__ dW, int ldw) {
  const int total = Nw * Kw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < parts; ++c) s += partial[(size_t)c * total + i];
    int n = i / Kw, k = i - n * Kw;
    dW[(size_t)n * ldw + k] += s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------

template <int BLOCK_N>
static int launch_linear(const CUtensorMap& tmA, const CUtensorMap& tmW, const LinearParams& p, int grid_y,
                         size_t smem_bytes, cudaStream_t st) {
  auto kern = linear_tc_kernel<BLOCK_N>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) {
    set_error("linear_tc(smem attr)", e);
    return (int)e;
  }
  long long gx = p.num_m_tiles < kNumSMs / grid_y ? p.num_m_tiles : kNumSMs / grid_y;
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)grid_y, 1);
  kern<<<grid, kThreads, smem_bytes, st>>>(tmA, tmW, p);
  return finish("linear_tc");
}

}  // namespace tc
}  // namespace pnb

using namespace pnb;
using namespace pnb::tc;

extern "C" int pnb_tc_available(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

extern "C" int pnb_linear_tc(long long M, int Nout, int K, const void* A, int lda, const void* W, int ldw, void* C,
                             int ldc, int c_dtype, const float* bias, const float* row_bias, int row_group,
                             const void* mask_src, int ld_mask, int flags, float* colsum_out, void* stream) {
  PNB_REQUIRE(M >= 0 && Nout >= 1 && Nout <= 512 && K >= 16 && K <= 384 && K % 16 == 0,
              "linear_tc: need 1<=Nout<=512, 16<=K<=384, K%16==0");
  PNB_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0),
              "linear_tc: A/W rows must be 16-byte aligned (ld % 8 == 0)");
  PNB_REQUIRE(!(flags & PNB_EPI_BIAS) || bias, "linear_tc: bias flag without bias");
  PNB_REQUIRE(!(flags & PNB_EPI_MASK) || mask_src, "linear_tc: mask flag without mask source");
  PNB_REQUIRE(row_bias == nullptr || row_group > 0, "linear_tc: row_bias needs row_group");
  if (M == 0) return 0;
  LinearParams p;
  p.M = M, p.Nout = Nout, p.K = K;
  p.num_kb = (K + kBlockK - 1) / kBlockK;
  p.tail_k16 = (K - (p.num_kb - 1) * kBlockK) / 16;
  p.C = C, p.ldc = ldc, p.c_dtype = c_dtype;
  p.bias = bias, p.row_bias = row_bias, p.row_group = row_group;
  p.mask_src = (const __nv_bfloat16*)mask_src, p.ld_mask = ld_mask, p.flags = flags;
  p.num_m_tiles = (M + kBlockM - 1) / kBlockM;
  p.colsum = nullptr;
  // smallest tile width covering Nout whose resident weight leaves room for >= 3 activation stages
  int bn = Nout <= 16 ? 16 : Nout <= 32 ? 32 : Nout <= 64 ? 64 : Nout <= 128 ? 128 : 256;
  auto smem_need = [&](int bn_, int stages) {
    return (size_t)p.num_kb * bn_ * kBlockK * 2 + (size_t)stages * kATileBytes + sizeof(Barriers) + 1024;
  };
  while (bn > 16 && smem_need(bn, 3) > (size_t)kSmemLimit) bn /= 2;
  p.stages = smem_need(bn, 4) <= (size_t)kSmemLimit ? 4 : 3;
  PNB_REQUIRE(smem_need(bn, p.stages) <= (size_t)kSmemLimit, "linear_tc: weight does not fit shared memory");
  int grid_y = (Nout + bn - 1) / bn;
  CUtensorMap tmA, tmW;
  if (!make_map(&tmA, A, (unsigned long long)M, (unsigned long long)K, (unsigned long long)lda, kBlockK, kBlockM))
    return PNB_ERR_ARG;
  cudaStream_t st = as_stream(stream);

  // ---- fast path: bf16 output, 64-column-aligned tiles, epilogue specialised at compile time -----------------
  const int epi = (flags & (PNB_EPI_BIAS | PNB_EPI_RELU | PNB_EPI_MASK | PNB_EPI_ACCUM)) | (row_bias ? PNB_EPI_ROWBIAS : 0);
  const bool aligned_c = (ldc % 8 == 0) && ((uintptr_t)C % 16 == 0) &&
                         (!(flags & PNB_EPI_MASK) || ((ld_mask % 8 == 0) && ((uintptr_t)mask_src % 16 == 0)));
  if (c_dtype == PNB_BF16 && aligned_c && Nout % 64 == 0 && (bn == 128 || bn == 256) && !getenv("PNB_NO_FAST_EPILOGUE")) {
    auto fast_need = [&](int bn_, int stages) {
      return (size_t)p.num_kb * bn_ * kBlockK * 2 + (size_t)stages * kATileBytes + 8 * (size_t)kStoreBoxBytes +
             (size_t)bn_ * 4 + sizeof(Barriers) + 1024;
    };
    int fbn = bn;
    while (fbn > 128 && fast_need(fbn, 2) > (size_t)kSmemLimit) fbn /= 2;
    int fst = 4;
    while (fst > 2 && fast_need(fbn, fst) > (size_t)kSmemLimit) --fst;
    if (fast_need(fbn, fst) <= (size_t)kSmemLimit) {
      LinearParams fp = p;
      fp.stages = fst;
      fp.colsum = colsum_out;
      int fgy = (Nout + fbn - 1) / fbn;
      CUtensorMap tmC;
      if (!make_map(&tmW, W, (unsigned long long)Nout, (unsigned long long)K, (unsigned long long)ldw, kBlockK, fbn))
        return PNB_ERR_ARG;
      if (!make_map(&tmC, C, (unsigned long long)M, (unsigned long long)Nout, (unsigned long long)ldc, 64, 32))
        return PNB_ERR_ARG;
      size_t fsm = fast_need(fbn, fst);
      long long gx = fp.num_m_tiles < kNumSMs / fgy ? fp.num_m_tiles : kNumSMs / fgy;
      if (gx < 1) gx = 1;
      dim3 grid((unsigned)gx, (unsigned)fgy, 1);
#define PNB_FAST_CASE(BN, E)                                                                                    \
  if (fbn == BN && epi == (E)) {                                                                               \
    auto kern = linear_tc_fast_kernel<BN, (E)>;                                                                \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);         \
    if (e != cudaSuccess) {                                                                                    \
      set_error("linear_tc_fast(smem attr)", e);                                                               \
      return (int)e;                                                                                           \
    }                                                                                                          \
    kern<<<grid, kFastThreads, fsm, st>>>(tmA, tmW, tmC, fp);                                                  \
    return finish("linear_tc_fast");                                                                           \
  }
#define PNB_FAST_BOTH(E) PNB_FAST_CASE(256, E) PNB_FAST_CASE(128, E)
      PNB_FAST_BOTH(PNB_EPI_BIAS | PNB_EPI_RELU)
      PNB_FAST_BOTH(PNB_EPI_BIAS)
      PNB_FAST_BOTH(PNB_EPI_RELU | PNB_EPI_ROWBIAS)
      PNB_FAST_BOTH(PNB_EPI_MASK)
      PNB_FAST_BOTH(PNB_EPI_MASK | PNB_EPI_ACCUM)
      PNB_FAST_BOTH(0)
#undef PNB_FAST_BOTH
#undef PNB_FAST_CASE
    }
  }

  // ---- general path (fp32 outputs, narrow heads, unusual flag combinations) ------------------------------------
  PNB_REQUIRE(colsum_out == nullptr, "linear_tc: colsum_out needs the bf16 fast path (Nout % 64 == 0, aligned)");
  size_t smem_bytes = smem_need(bn, p.stages);
  if (!make_map(&tmW, W, (unsigned long long)Nout, (unsigned long long)K, (unsigned long long)ldw, kBlockK, bn))
    return PNB_ERR_ARG;
  switch (bn) {
    case 16: return launch_linear<16>(tmA, tmW, p, grid_y, smem_bytes, st);
    case 32: return launch_linear<32>(tmA, tmW, p, grid_y, smem_bytes, st);
    case 64: return launch_linear<64>(tmA, tmW, p, grid_y, smem_bytes, st);
    case 128: return launch_linear<128>(tmA, tmW, p, grid_y, smem_bytes, st);
    default: return launch_linear<256>(tmA, tmW, p, grid_y, smem_bytes, st);
  }
}

extern "C" long long pnb_wgrad_tc_workspace(int Nw, int Kw) { return (long long)kNumSMs * Nw * Kw * 4; }

extern "C" int pnb_wgrad_tc(long long M, int Nw, int Kw, const void* dZ, int ldz, const void* X, int ldx, float* dW,
                            int ldw, float* workspace, void* stream) {
  PNB_REQUIRE(M >= 0 && (Nw == 128 || Nw == 256) && Kw >= 16 && Kw <= 256 && Kw % 16 == 0,
              "wgrad_tc: need Nw in {128,256}, 16<=Kw<=256, Kw%16==0");
  PNB_REQUIRE(ldz % 8 == 0 && ldx % 8 == 0 && ((uintptr_t)dZ % 16 == 0) && ((uintptr_t)X % 16 == 0),
              "wgrad_tc: operand rows must be 16-byte aligned");
  PNB_REQUIRE(workspace != nullptr, "wgrad_tc: workspace required (pnb_wgrad_tc_workspace bytes)");
  if (M == 0) return 0;
  WgradParams p;
  p.M = M, p.Nw = Nw, p.Kw = Kw;
  p.n_halves = Nw / 128;
  p.kw_chunks = (Kw + 63) / 64;
  int stage_bytes = (p.n_halves * 2 + p.kw_chunks) * kWgBoxBytes;
  p.stages = 4;
  while (p.stages > 2 && (size_t)p.stages * stage_bytes + sizeof(Barriers) + 1024 > (size_t)kSmemLimit) --p.stages;
  size_t smem_bytes = (size_t)p.stages * stage_bytes + sizeof(Barriers) + 1024;
  PNB_REQUIRE(smem_bytes <= (size_t)kSmemLimit, "wgrad_tc: tile does not fit shared memory");
  long long blocks64 = (M + kWgRows - 1) / kWgRows;
  int grid = (int)(blocks64 < kNumSMs ? blocks64 : kNumSMs);
  p.rows_per_cta = (blocks64 + grid - 1) / grid * kWgRows;
  p.partial = workspace;
  CUtensorMap tmZ, tmX;
  if (!make_map(&tmZ, dZ, (unsigned long long)M, (unsigned long long)Nw, (unsigned long long)ldz, 64, kWgRows))
    return PNB_ERR_ARG;
  if (!make_map(&tmX, X, (unsigned long long)M, (unsigned long long)Kw, (unsigned long long)ldx, 64, kWgRows))
    return PNB_ERR_ARG;
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) {
    set_error("wgrad_tc(smem attr)", e);
    return (int)e;
  }
  cudaStream_t st = as_stream(stream);
  wgrad_tc_kernel<<<grid, kThreads, smem_bytes, st>>>(tmZ, tmX, p);
  int rc = finish("wgrad_tc");
  if (rc) return rc;
  int total = Nw * Kw;
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(grid, Nw, Kw, workspace, dW, ldw);
  return finish("wgrad_reduce");
}
