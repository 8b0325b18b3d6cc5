// K5-batched: every weight gradient of one backward pass in ONE launch.
//
//   job j:  dW_j[Nw,Kw] += Z_j[M,Nw]^T X_j[M,Kw]   and (optionally)   db_j[Nw] += column sums of Z_j
//
// (autograd of the nn.Linear layers of models/pano_mip_nerf.py:54-76; Z = dL/d(pre-activation), X = layer input).
// All jobs reduce over the same M samples.  The 148 CTAs are split between the jobs in proportion to their operand
// bytes; a CTA owns one job and one contiguous slab of samples for the whole launch, keeps the fp32 partial of dW in
// TMEM (Nw/128 accumulators of 128 x Kw), and writes it out once.  Compared with one launch per layer this divides
// the partial-sum traffic by ~15, removes 40+ launches and their tails per training step, and - because the Z tiles
// pass through shared memory anyway - the four otherwise idle warps add up the bias gradient on the side, so no
// separate column-sum pass over the dz planes is needed.
//
// Both operands are MN-major (the reduction axis is the slow axis in memory): TMA brings 64-sample x 64-column
// boxes (128B swizzle) straight from the row-major planes, tcgen05 consumes them through MN-major descriptors.
// A second tiny kernel adds the per-CTA partials in a fixed order (deterministic) into the gradient buffers.
#include "tc_common.cuh"

namespace pnb {
namespace tc {

constexpr int kWbMaxMaps = 8, kWbMaxJobs = 40;
constexpr int kWbRows = 64;                         // samples per pipeline stage
constexpr int kWbBoxBytes = kWbRows * kBlockK * 2;  // one 64x64 bf16 box = 8 KB
constexpr int kWbPartial = 256 * 256 + 256;         // floats per CTA in the workspace: dW partial + colsum partial

struct WbJob {
  int zmap, zplane, xmap, xplane;
  int Nw, Kw, cta0, ncta;
};
struct WbParams {
  CUtensorMap maps[kWbMaxMaps];
  WbJob jobs[kWbMaxJobs];
  unsigned char cta_job[kNumSMs];
  unsigned char job_colsum[kWbMaxJobs];
  long long M;
  float* partial;  // [gridDim.x][kWbPartial]
};
constexpr int kWbMaxRanges = 4;  // jobs that accumulate into the same dW are reduced together (no races)
struct WbReduceJob {
  float* dW;
  int Nw, Kw, ldw, n_ranges;
  int cta0[kWbMaxRanges], ncta[kWbMaxRanges];
};
struct WbBiasJob {
  float* db;
  int Nw, cta0, ncta, pad;
};
struct WbReduceParams {
  WbReduceJob jobs[kWbMaxJobs];
  WbBiasJob bias[kWbMaxJobs];
  int n_unique, n_bias;
  const float* partial;
};

struct WbBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_batch_kernel(const __grid_constant__ WbParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jid = p.cta_job[blockIdx.x];
  const WbJob job = p.jobs[jid];
  const bool colsum = p.job_colsum[jid] != 0;
  const int n_halves = job.Nw / 128;
  const int z_chunks = job.Nw / 64;
  const int kw_chunks = (job.Kw + 63) / 64;
  const int stage_bytes = (z_chunks + kw_chunks) * kWbBoxBytes;
  int stages = (kSmemLimit - 1024 - (int)sizeof(WbBarriers)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  WbBarriers* bars = reinterpret_cast<WbBarriers*>(smem + (size_t)stages * stage_bytes);
  // slab of this CTA: multiples of 64 rows
  const long long blocks64 = (p.M + kWbRows - 1) / kWbRows;
  const long long per = (blocks64 + job.ncta - 1) / job.ncta;
  const long long b0 = (long long)(blockIdx.x - job.cta0) * per;
  long long b1 = b0 + per;
  if (b1 > blocks64) b1 = blocks64;
  const int num_it = b1 > b0 ? (int)(b1 - b0) : 0;
  const int kw_pad = kw_chunks * 64;  // TMEM columns per accumulator

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.maps[job.zmap]);
    tma_prefetch_desc(&p.maps[job.xmap]);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 5);  // MMA commit + the four column-sum warps
    }
    mbar_init(&bars->tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
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
        const int row = (int)((b0 + it) * kWbRows);  // the global tail is partial: TMA zero-fills it
        for (int c = 0; c < z_chunks; ++c)
          tma_load_3d(base + (size_t)c * kWbBoxBytes, &p.maps[job.zmap], &bars->full[stage], c * 64, row, job.zplane);
        for (int c = 0; c < kw_chunks; ++c)
          tma_load_3d(base + (size_t)(z_chunks + c) * kWbBoxBytes, &p.maps[job.xmap], &bars->full[stage], c * 64, row,
                      job.xplane);
        if (++stage == stages) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc_bf16(128, job.Kw, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&bars->full[stage], phase);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t x_addr = base + z_chunks * kWbBoxBytes;
#pragma unroll 1
        for (int k = 0; k < kWbRows / 16; ++k) {
          // MN-major, 128B swizzle: 64-element column chunks are one box (8 KB) apart (LBO), 8-sample groups are
          // 1024 B apart (SBO); a K=16 slice (16 samples) starts 2048 B further.
          const uint64_t bd = smem_desc_sw128(x_addr + k * 2048, kWbBoxBytes, 1024);
          for (int h = 0; h < n_halves; ++h) {
            const uint64_t ad = smem_desc_sw128(base + h * 2 * kWbBoxBytes + k * 2048, kWbBoxBytes, 1024);
            umma_f16(tmem_base + h * kw_pad, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
        }
        umma_commit(&bars->empty[stage]);
        if (++stage == stages) stage = 0, phase ^= 1;
      }
      umma_commit(&bars->tmem_full);
    }
    __syncwarp();
  } else {
    // ---- warps 2..5: bias gradient on the side, then the epilogue -------------------------------------------
    const int tid = threadIdx.x - 64;  // 0..127: columns 2*tid, 2*tid+1 of Z
    float s0 = 0.f, s1 = 0.f;
    {
      const int chunk = tid >> 5, w = tid & 31;
      const bool active = colsum && chunk < z_chunks;
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < num_it; ++it) {
        mbar_wait(&bars->full[stage], phase);
        if (active) {
          const uint8_t* base = smem + (size_t)stage * stage_bytes + (size_t)chunk * kWbBoxBytes + (w & 3) * 4;
          const int j = w >> 2;
#pragma unroll 8
          for (int r = 0; r < kWbRows; ++r) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(base + r * 128 + ((j ^ (r & 7)) << 4));
            s0 += __uint_as_float(v << 16);
            s1 += __uint_as_float(v & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[stage]);
        if (++stage == stages) stage = 0, phase ^= 1;
      }
    }
    float* out = p.partial + (size_t)blockIdx.x * kWbPartial;
    if (colsum && 2 * tid < job.Nw) {
      out[256 * 256 + 2 * tid] = s0;
      out[256 * 256 + 2 * tid + 1] = s1;
    }
    const int q = warp & 3;
    mbar_wait(&bars->tmem_full, 0);
    tc_fence_after();
    for (int h = 0; h < n_halves; ++h) {
      const int n = h * 128 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * kw_pad;
#pragma unroll 1
      for (int c0 = 0; c0 < job.Kw; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (num_it == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        float* orow = out + (size_t)n * job.Kw + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (c0 + i < job.Kw) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void wgrad_batch_reduce_kernel(const __grid_constant__ WbReduceParams p) {
  if ((int)blockIdx.y == p.n_unique) {  // bias gradients: one block, entry after entry (duplicates stay ordered)
    if (blockIdx.x != 0) return;
    for (int e = 0; e < p.n_bias; ++e) {
      const WbBiasJob b = p.bias[e];
      for (int n = threadIdx.x; n < b.Nw; n += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < b.ncta; ++c) s += p.partial[(size_t)(b.cta0 + c) * kWbPartial + 256 * 256 + n];
        b.db[n] += s;
      }
    }
    return;
  }
  const WbReduceJob& j = p.jobs[blockIdx.y];
  const int total = j.Nw * j.Kw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < j.n_ranges; ++r)
      for (int c = 0; c < j.ncta[r]; ++c) s += p.partial[(size_t)(j.cta0[r] + c) * kWbPartial + i];
    const int n = i / j.Kw, k = i - n * j.Kw;
    j.dW[(size_t)n * j.ldw + k] += s;
  }
}

static bool make_map3(CUtensorMap* out, const void* base, unsigned long long planes, unsigned long long rows,
                      unsigned long long cols, unsigned long long ld) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error_msg("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  cuuint64_t dims[3] = {cols, rows, planes};
  cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)kWbRows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error_msg("cuTensorMapEncodeTiled failed for a wgrad operand");
    return false;
  }
  return true;
}

}  // namespace tc
}  // namespace pnb

using namespace pnb;
using namespace pnb::tc;

extern "C" long long pnb_wgrad_batch_workspace(void) { return (long long)kNumSMs * kWbPartial * 4; }

// maps:  map_base[i] = device pointer of a bf16 tensor [planes][M][ld]; map_desc[3i..] = {planes, ld, cols}
// jobs:  jobs[8j..] = {zmap, zplane, xmap, xplane, Nw, Kw, ldw, want_colsum}; dW[j] fp32 [Nw, ldw] (accumulated),
//        db[j] fp32 [Nw] (accumulated, may be null)
extern "C" int pnb_wgrad_batch(long long M, int n_maps, const void* const* map_base, const long long* map_desc,
                               int n_jobs, const long long* jobs, const void* const* dW, const void* const* db,
                               float* workspace, void* stream) {
  PNB_REQUIRE(M >= 0 && n_maps >= 1 && n_maps <= kWbMaxMaps && n_jobs >= 1 && n_jobs <= kWbMaxJobs,
              "wgrad_batch: too many maps / jobs");
  PNB_REQUIRE(map_base && map_desc && jobs && dW && db && workspace, "wgrad_batch: null argument");
  PNB_REQUIRE(M < (1ll << 31) - 64, "wgrad_batch: M too large for 32-bit TMA coordinates");
  if (M == 0) return 0;
  static thread_local WbParams p;  // large: kept off the stack; one per host thread (autograd's backward thread)
  static thread_local WbReduceParams rp;
  for (int i = 0; i < n_maps; ++i) {
    const long long planes = map_desc[3 * i], ld = map_desc[3 * i + 1], cols = map_desc[3 * i + 2];
    PNB_REQUIRE(planes >= 1 && ld % 8 == 0 && cols >= 1 && cols <= ld && ((uintptr_t)map_base[i] % 16 == 0),
                "wgrad_batch: operand rows must be 16-byte aligned");
    if (!make_map3(&p.maps[i], map_base[i], (unsigned long long)planes, (unsigned long long)M,
                   (unsigned long long)cols, (unsigned long long)ld))
      return PNB_ERR_ARG;
  }
  // split the CTAs between the jobs in proportion to their operand bytes (every job gets at least one)
  long long cost[kWbMaxJobs], total = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long* q = jobs + 8 * j;
    const int Nw = (int)q[4], Kw = (int)q[5];
    PNB_REQUIRE((Nw == 128 || Nw == 256) && Kw >= 16 && Kw <= 256 && Kw % 16 == 0,
                "wgrad_batch: need Nw in {128,256}, 16<=Kw<=256, Kw%16==0");
    PNB_REQUIRE(q[0] >= 0 && q[0] < n_maps && q[2] >= 0 && q[2] < n_maps && dW[j] != nullptr,
                "wgrad_batch: bad job");
    cost[j] = Nw / 64 + (Kw + 63) / 64;
    total += cost[j];
  }
  PNB_REQUIRE(n_jobs <= kNumSMs, "wgrad_batch: more jobs than CTAs");
  int ncta[kWbMaxJobs], used = 0;
  for (int j = 0; j < n_jobs; ++j) {
    ncta[j] = (int)((long long)kNumSMs * cost[j] / total);
    if (ncta[j] < 1) ncta[j] = 1;
    used += ncta[j];
  }
  while (used > kNumSMs) {  // take from the most generously served job
    int best = 0;
    for (int j = 1; j < n_jobs; ++j)
      if ((double)ncta[j] / cost[j] > (double)ncta[best] / cost[best] && ncta[j] > 1) best = j;
    --ncta[best], --used;
  }
  while (used < kNumSMs) {  // give to the most starved job
    int best = 0;
    for (int j = 1; j < n_jobs; ++j)
      if ((double)ncta[j] / cost[j] < (double)ncta[best] / cost[best]) best = j;
    ++ncta[best], ++used;
  }
  const long long blocks64 = (M + kWbRows - 1) / kWbRows;
  int cta = 0;
  rp.n_unique = 0, rp.n_bias = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long* q = jobs + 8 * j;
    WbJob& jb = p.jobs[j];
    jb.zmap = (int)q[0], jb.zplane = (int)q[1], jb.xmap = (int)q[2], jb.xplane = (int)q[3];
    jb.Nw = (int)q[4], jb.Kw = (int)q[5];
    if (ncta[j] > blocks64) ncta[j] = (int)blocks64;  // tiny batches: never more CTAs than 64-row blocks
    jb.cta0 = cta, jb.ncta = ncta[j];
    p.job_colsum[j] = (unsigned char)(q[7] != 0 && db[j] != nullptr);
    for (int c = 0; c < ncta[j]; ++c) p.cta_job[cta + c] = (unsigned char)j;
    {
      float* dst = reinterpret_cast<float*>(const_cast<void*>(dW[j]));
      int r = 0;
      for (; r < rp.n_unique; ++r)
        if (rp.jobs[r].dW == dst && rp.jobs[r].Nw == jb.Nw && rp.jobs[r].Kw == jb.Kw && rp.jobs[r].ldw == (int)q[6] &&
            rp.jobs[r].n_ranges < kWbMaxRanges)
          break;
      if (r == rp.n_unique) {
        ++rp.n_unique;
        rp.jobs[r].dW = dst, rp.jobs[r].Nw = jb.Nw, rp.jobs[r].Kw = jb.Kw, rp.jobs[r].ldw = (int)q[6];
        rp.jobs[r].n_ranges = 0;
      }
      rp.jobs[r].cta0[rp.jobs[r].n_ranges] = cta, rp.jobs[r].ncta[rp.jobs[r].n_ranges] = ncta[j];
      ++rp.jobs[r].n_ranges;
      if (p.job_colsum[j]) {
        WbBiasJob& bj = rp.bias[rp.n_bias++];
        bj.db = reinterpret_cast<float*>(const_cast<void*>(db[j]));
        bj.Nw = jb.Nw, bj.cta0 = cta, bj.ncta = ncta[j];
      }
    }
    cta += ncta[j];
  }
  p.M = M, p.partial = workspace;
  rp.partial = workspace;
  const size_t smem_bytes = kSmemLimit;
  cudaError_t e = cudaFuncSetAttribute(wgrad_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) {
    set_error("wgrad_batch(smem attr)", e);
    return (int)e;
  }
  cudaStream_t st = as_stream(stream);
  wgrad_batch_kernel<<<cta, kThreads, smem_bytes, st>>>(p);
  int rc = finish("wgrad_batch");
  if (rc) return rc;
  wgrad_batch_reduce_kernel<<<dim3(32, rp.n_unique + 1), 256, 0, st>>>(rp);
  return finish("wgrad_batch_reduce");
}
