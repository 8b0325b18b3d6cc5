// K5-batched: every weight gradient of one backward pass in ONE launch.
//
//   job j:  dW_j[Nw,Kw] += Z_j[M,Nw]^T X_j[M,Kw]   and (optionally)   db_j[Nw] += column sums of Z_j
//
// (autograd of the nn.Linear layers of models/pano_mip_nerf.py:54-76; Z = dL/d(pre-activation), X = layer input).
// All jobs reduce over the same M samples.  The work of a launch is the list of (job, 64-sample block) units in
// job-major order, weighted by the operand bytes of the unit; the list is cut into gridDim.x pieces of EQUAL BYTES,
// so a CTA owns a contiguous run of blocks that may straddle a job boundary ("segments": usually one or two per
// CTA).  Per segment the CTA keeps the fp32 partial of dW in TMEM (Nw/128 accumulators of 128 x Kw) and writes it out
// once into its own workspace slot.  (Round 1-2 gave every job a whole number of CTAs: with 12-33 jobs on 148 CTAs
// the rounding left SMs 3-13 % tensor-active while others still had a quarter of their slab to go.)  Compared with
// one launch per layer this divides the partial-sum traffic by ~15, removes 40+ launches and their tails per
// training step, and - because the Z tiles pass through shared memory anyway - the four otherwise idle warps add up
// the bias gradient on the side, so no separate column-sum pass over the dz planes is needed.
//
// Both operands are MN-major (the reduction axis is the slow axis in memory): TMA brings 64-sample x 64-column
// boxes (128B swizzle) straight from the row-major planes, tcgen05 consumes them through MN-major descriptors.
// A second tiny kernel adds the per-segment partials in a fixed order (deterministic) into the gradient buffers.
#include "tc_common.cuh"

namespace pnb {
namespace tc {

constexpr int kWbMaxMaps = 8, kWbMaxJobs = 40;
constexpr int kWbRows = 64;                         // samples per pipeline stage
constexpr int kWbBoxBytes = kWbRows * kBlockK * 2;  // one 64x64 bf16 box = 8 KB
constexpr int kWbPartial = 256 * 256 + 256;         // floats per workspace slot: dW partial + colsum partial
constexpr int kWbRingBoxes = 27;                    // shared-memory ring: 27 boxes = 216 KB
constexpr int kWbMaxStages = 9;                     // a stage = the boxes of one 64-sample block of the current job
constexpr int kWbMaxSlots = kNumSMs + kWbMaxJobs;   // every (CTA, job) segment owns a slot

struct WbJob {
  int zmap, zplane, xmap, xplane;
  int Nw, Kw, cost, pad;  // cost = operand bytes per sample row
};
struct WbParams {
  CUtensorMap maps[kWbMaxMaps];
  WbJob jobs[kWbMaxJobs];
  long long start[kWbMaxJobs + 1];  // position of block 0 of job j in the linear order (bytes / 64)
  long long cut[kNumSMs + 1];       // CTA c owns the blocks whose position lies in [cut[c], cut[c+1])
  unsigned short cta_slot0[kNumSMs];
  unsigned char cta_job0[kNumSMs];
  unsigned char job_colsum[kWbMaxJobs];
  int n_jobs;
  long long blocks64;
  float* partial;  // [slots][kWbPartial]
  int max_stages;   // <= kWbMaxStages (PNB_WB_MAXSTAGES: experiments)
  long long* prof;  // PNB_WB_PROF=1: per-CTA {start ns, end ns} (debugging aid), else null
};
constexpr int kWbMaxRanges = 4;  // jobs that accumulate into the same dW are reduced together (no races)
struct WbReduceJob {
  float* dW;
  int Nw, Kw, ldw, n_ranges;
  int slot0[kWbMaxRanges], nslot[kWbMaxRanges];
};
struct WbBiasJob {
  float* db;
  int Nw, slot0, nslot, pad;
};
struct WbReduceParams {
  WbReduceJob jobs[kWbMaxJobs];
  WbBiasJob bias[kWbMaxJobs];
  int n_unique, n_bias;
  const float* partial;
};

struct WbBarriers {
  uint64_t full[kWbMaxStages];
  uint64_t empty[kWbMaxStages];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint32_t tmem_base;
};

// The cut of the linear work list: CTA c owns the blocks whose start position lies in [cut[c], cut[c+1]).  Host (slot
// numbering) and device (the work itself) evaluate the same integer expressions on the same cuts.
__host__ __device__ inline void wb_segment(long long lo, long long hi, long long start, int cost, long long blocks64,
                                           long long* b0, long long* b1) {
  const long long a = lo - start, b = hi - start;
  long long x0 = a <= 0 ? 0 : (a + cost - 1) / cost;
  long long x1 = b <= 0 ? 0 : (b + cost - 1) / cost;
  *b0 = x0 > blocks64 ? blocks64 : x0;
  *b1 = x1 > blocks64 ? blocks64 : x1;
}

__device__ __forceinline__ int wb_stages(int boxes_per_block, int max_stages) {
  const int s = kWbRingBoxes / boxes_per_block;
  return s > max_stages ? max_stages : s;
}

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
  WbBarriers* bars = reinterpret_cast<WbBarriers*>(smem + (size_t)kWbRingBoxes * kWbBoxBytes);
  const long long lo = p.cut[blockIdx.x], hi = p.cut[blockIdx.x + 1];
  const int job0 = p.cta_job0[blockIdx.x];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kWbMaxStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 5);  // MMA commit + the four column-sum warps
    }
    mbar_init(&bars->tmem_full, 1);
    mbar_init(&bars->tmem_empty, 4);  // the four epilogue warps
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if (p.prof != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[2 * blockIdx.x] = t;
  }

  // Every role walks the same list of segments.  The ring is re-cut per segment into as many stages as fit (a stage =
  // the boxes of one block of that job: 3 stages for two 256-wide operands, 9 for the narrow head jobs), so that every
  // CTA keeps ~200 KB in flight whatever its job: under a saturated HBM a CTA's share of the bandwidth is proportional
  // to its bytes in flight, and equal shares are what makes the equal-bytes cut finish together.  Every role restarts
  // at stage 0 of the new geometry; the producer first waits until the old stages have been consumed.  Bit s of
  // `par` is the parity of the role's next wait on stage s (the barriers live on across segments).
  if (warp == 0) {
    if (lane == 0) {
      uint32_t par = 0;
      for (int j = job0; j < p.n_jobs && p.start[j] < hi; ++j) {
        const WbJob job = p.jobs[j];
        long long b0, b1;
        wb_segment(lo, hi, p.start[j], job.cost, p.blocks64, &b0, &b1);
        if (b1 <= b0) continue;
        tma_prefetch_desc(&p.maps[job.zmap]);
        tma_prefetch_desc(&p.maps[job.xmap]);
        const int z_chunks = job.Nw / 64, kw_chunks = (job.Kw + 63) / 64;
        const uint32_t bytes = (uint32_t)((z_chunks + kw_chunks) * kWbBoxBytes);
        const int stages = wb_stages(z_chunks + kw_chunks, p.max_stages);
        // drain: the previous segment's stages overlap the new ones at other offsets (waits do not advance `par`)
        for (int s = 0; s < kWbMaxStages; ++s) mbar_wait(&bars->empty[s], ((par >> s) & 1) ^ 1);
        int stage = 0;
        for (long long b = b0; b < b1; ++b) {
          mbar_wait(&bars->empty[stage], ((par >> stage) & 1) ^ 1);
          par ^= 1u << stage;
          mbar_expect_tx(&bars->full[stage], bytes);
          uint8_t* base = smem + (size_t)stage * bytes;
          const int row = (int)(b * kWbRows);  // the global tail is partial: TMA zero-fills it
          for (int c = 0; c < z_chunks; ++c)
            tma_load_3d(base + (size_t)c * kWbBoxBytes, &p.maps[job.zmap], &bars->full[stage], c * 64, row, job.zplane);
          for (int c = 0; c < kw_chunks; ++c)
            tma_load_3d(base + (size_t)(z_chunks + c) * kWbBoxBytes, &p.maps[job.xmap], &bars->full[stage], c * 64, row,
                        job.xplane);
          if (++stage == stages) stage = 0;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int seg = 0;
      uint32_t par = 0;
      for (int j = job0; j < p.n_jobs && p.start[j] < hi; ++j) {
        const WbJob job = p.jobs[j];
        long long b0, b1;
        wb_segment(lo, hi, p.start[j], job.cost, p.blocks64, &b0, &b1);
        if (b1 <= b0) continue;
        const int n_halves = job.Nw / 128, z_chunks = job.Nw / 64, kw_pad = ((job.Kw + 63) / 64) * 64;
        const int stage_bytes = (z_chunks + kw_pad / 64) * kWbBoxBytes, stages = wb_stages(z_chunks + kw_pad / 64, p.max_stages);
        int stage = 0;
        const uint32_t idesc = instr_desc_bf16(128, job.Kw, 1, 1);
        if (seg > 0) {  // the accumulators still hold the previous segment until its epilogue has read them
          mbar_wait(&bars->tmem_empty, (uint32_t)((seg - 1) & 1));
          tc_fence_after();
        }
        for (long long b = b0; b < b1; ++b) {
          mbar_wait(&bars->full[stage], (par >> stage) & 1);
          par ^= 1u << stage;
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
              umma_f16(tmem_base + h * kw_pad, ad, bd, idesc, (uint32_t)((b != b0) | (k != 0)));
            }
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == stages) stage = 0;
        }
        umma_commit(&bars->tmem_full);
        ++seg;
      }
    }
    __syncwarp();
  } else {
    // ---- warps 2..5: bias gradient on the side, then the epilogue of the segment ------------------------------
    // column sums: warp `chunk` owns the 64-column box `chunk` of Z; lane = (row group rg, 16-byte unit u): it adds the
    // 8 columns 8u..8u+7 of the rows r = 4 i + rg, the four row groups are folded with shuffles at the segment's end
    const int tid = threadIdx.x - 64;
    const int chunk = tid >> 5, u = lane & 7, rg = lane >> 3;
    const int q = warp & 3;
    int seg = 0;
    uint32_t par = 0;
    int slot = p.cta_slot0[blockIdx.x];
    for (int j = job0; j < p.n_jobs && p.start[j] < hi; ++j) {
      const WbJob job = p.jobs[j];
      long long b0, b1;
      wb_segment(lo, hi, p.start[j], job.cost, p.blocks64, &b0, &b1);
      if (b1 <= b0) continue;
      const int n_halves = job.Nw / 128, z_chunks = job.Nw / 64, kw_pad = ((job.Kw + 63) / 64) * 64;
      const bool colsum = p.job_colsum[j] != 0;
      const bool active = colsum && chunk < z_chunks;
      const int stage_bytes = (z_chunks + kw_pad / 64) * kWbBoxBytes, stages = wb_stages(z_chunks + kw_pad / 64, p.max_stages);
      int stage = 0;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (long long b = b0; b < b1; ++b) {
        mbar_wait(&bars->full[stage], (par >> stage) & 1);
        par ^= 1u << stage;
        if (active) {
          const uint8_t* base = smem + (size_t)stage * stage_bytes + (size_t)chunk * kWbBoxBytes + rg * 128;
#pragma unroll
          for (int i = 0; i < kWbRows / 4; ++i) {
            const int r = 4 * i + rg;
            const uint4 v = *reinterpret_cast<const uint4*>(base + i * 512 + ((u ^ (r & 7)) << 4));
            acc[0] += __uint_as_float(v.x << 16), acc[1] += __uint_as_float(v.x & 0xffff0000u);
            acc[2] += __uint_as_float(v.y << 16), acc[3] += __uint_as_float(v.y & 0xffff0000u);
            acc[4] += __uint_as_float(v.z << 16), acc[5] += __uint_as_float(v.z & 0xffff0000u);
            acc[6] += __uint_as_float(v.w << 16), acc[7] += __uint_as_float(v.w & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->empty[stage]);
        if (++stage == stages) stage = 0;
      }
      float* out = p.partial + (size_t)slot * kWbPartial;
      if (colsum) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
          acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
        }
        if (active && rg == 0) {
          float* o = out + 256 * 256 + chunk * 64 + u * 8;
          *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
      mbar_wait(&bars->tmem_full, (uint32_t)(seg & 1));
      tc_fence_after();
      for (int h = 0; h < n_halves; ++h) {
        const int n = h * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * kw_pad;
#pragma unroll 1
        for (int c0 = 0; c0 < job.Kw; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          float* orow = out + (size_t)n * job.Kw + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c0 + i < job.Kw) *reinterpret_cast<float4*>(orow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty);
      ++seg, ++slot;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.prof != nullptr && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[2 * blockIdx.x + 1] = t;
  }
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
        for (int c = 0; c < b.nslot; ++c) s += p.partial[(size_t)(b.slot0 + c) * kWbPartial + 256 * 256 + n];
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
      for (int c = 0; c < j.nslot[r]; ++c) s += p.partial[(size_t)(j.slot0[r] + c) * kWbPartial + i];
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

// Weight of one 64-sample block of a job in the cut = the time a CTA needs for it, in ns, fitted on a B200 from the
// per-CTA clocks of a launch (PNB_WB_PROF=1; profiles/r02_wgrad_batch_balance.md).  With both operands MN-major a
// K=16 MMA step costs ~250 cycles whatever N = Kw is, so a block costs by its Z width (1.08 us for the two 128-row
// halves of a 256-wide Z, 0.69 us for one); the X boxes add ~13 ns each, the bias column sums 0.14 us.
// PNB_WB_SPLIT=bytes weighs by operand bytes instead (measured 25-50 % slower: the narrow jobs move fewer bytes per
// unit of time, their CTAs finish last and the HBM idles with them).
static int wb_block_cost(int Nw, int Kw, bool colsum) {
  static const bool by_bytes = [] {
    const char* e = getenv("PNB_WB_SPLIT");
    return e != nullptr && e[0] == 'b';
  }();
  if (by_bytes) return 2 * (Nw + Kw);
  const int xb = (Kw + 63) / 64;
  return Nw == 256 ? 1075 + 13 * xb + (colsum ? 140 : 0) : 685 + 5 * xb + (colsum ? 70 : 0);
}

// Cut policies.  Default: kNumSMs pieces of equal weight (wb_block_cost), a piece may straddle job boundaries.
// "jobs" (PNB_WB_SPLIT=jobs, the round 1-2 split, kept for A/B runs): every job gets a whole number of CTAs in
// proportion to its boxes, the job's blocks are divided evenly between them.
static void wb_build_cuts(const long long* start, const int* cost, const int* boxes, int n_jobs, long long blocks64,
                          long long* cut) {
  static const bool by_jobs = [] {
    const char* e = getenv("PNB_WB_SPLIT");
    return e != nullptr && e[0] == 'j';
  }();
  if (!by_jobs || n_jobs > kNumSMs) {
    for (int c = 0; c <= kNumSMs; ++c) cut[c] = start[n_jobs] / kNumSMs * c + start[n_jobs] % kNumSMs * c / kNumSMs;
    cut[kNumSMs] = start[n_jobs];
    return;
  }
  long long total = 0;
  for (int j = 0; j < n_jobs; ++j) total += boxes[j];
  int ncta[kWbMaxJobs], used = 0;
  for (int j = 0; j < n_jobs; ++j) {
    ncta[j] = (int)((long long)kNumSMs * boxes[j] / total);
    if (ncta[j] < 1) ncta[j] = 1;
    used += ncta[j];
  }
  while (used > kNumSMs) {
    int best = 0;
    for (int j = 1; j < n_jobs; ++j)
      if ((double)ncta[j] / boxes[j] > (double)ncta[best] / boxes[best] && ncta[j] > 1) best = j;
    --ncta[best], --used;
  }
  while (used < kNumSMs) {
    int best = 0;
    for (int j = 1; j < n_jobs; ++j)
      if ((double)ncta[j] / boxes[j] < (double)ncta[best] / boxes[best]) best = j;
    ++ncta[best], ++used;
  }
  int c = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long per = (blocks64 + ncta[j] - 1) / ncta[j];
    for (int i = 0; i < ncta[j]; ++i, ++c) {
      long long b = (long long)i * per;
      if (b > blocks64) b = blocks64;
      cut[c] = start[j] + b * cost[j];
    }
  }
  cut[kNumSMs] = start[n_jobs];
}

extern "C" long long pnb_wgrad_batch_workspace(void) { return (long long)kWbMaxSlots * kWbPartial * 4; }

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
  // the linear work list: job after job, each 64-sample block weighted by the operand bytes it moves
  const long long blocks64 = (M + kWbRows - 1) / kWbRows;
  p.start[0] = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long* q = jobs + 8 * j;
    const int Nw = (int)q[4], Kw = (int)q[5];
    PNB_REQUIRE((Nw == 128 || Nw == 256) && Kw >= 16 && Kw <= 256 && Kw % 16 == 0,
                "wgrad_batch: need Nw in {128,256}, 16<=Kw<=256, Kw%16==0");
    PNB_REQUIRE(q[0] >= 0 && q[0] < n_maps && q[2] >= 0 && q[2] < n_maps && dW[j] != nullptr,
                "wgrad_batch: bad job");
    WbJob& jb = p.jobs[j];
    jb.zmap = (int)q[0], jb.zplane = (int)q[1], jb.xmap = (int)q[2], jb.xplane = (int)q[3];
    p.job_colsum[j] = (unsigned char)(q[7] != 0 && db[j] != nullptr);
    jb.Nw = Nw, jb.Kw = Kw, jb.cost = wb_block_cost(Nw, Kw, p.job_colsum[j]), jb.pad = 0;
    p.start[j + 1] = p.start[j] + blocks64 * jb.cost;
  }
  p.n_jobs = n_jobs, p.blocks64 = blocks64;
  const int grid = kNumSMs;
  {
    int cost[kWbMaxJobs], boxes[kWbMaxJobs];
    for (int j = 0; j < n_jobs; ++j) cost[j] = p.jobs[j].cost, boxes[j] = p.jobs[j].Nw / 64 + (p.jobs[j].Kw + 63) / 64;
    wb_build_cuts(p.start, cost, boxes, n_jobs, blocks64, p.cut);
  }
  // slot numbering: CTA after CTA, segment after segment - the segments of one job get consecutive slots
  int job_slot0[kWbMaxJobs], job_nslot[kWbMaxJobs];
  for (int j = 0; j < n_jobs; ++j) job_slot0[j] = -1, job_nslot[j] = 0;
  int slot = 0;
  for (int c = 0; c < grid; ++c) {
    const long long lo = p.cut[c], hi = p.cut[c + 1];
    p.cta_slot0[c] = (unsigned short)slot;
    int first = n_jobs;
    for (int j = 0; j < n_jobs && p.start[j] < hi; ++j) {
      long long b0, b1;
      wb_segment(lo, hi, p.start[j], p.jobs[j].cost, blocks64, &b0, &b1);
      if (b1 <= b0) continue;
      if (first == n_jobs) first = j;
      if (job_slot0[j] < 0) job_slot0[j] = slot;
      PNB_REQUIRE(job_slot0[j] + job_nslot[j] == slot && slot < kWbMaxSlots, "wgrad_batch: internal slot numbering");
      ++job_nslot[j], ++slot;
    }
    p.cta_job0[c] = (unsigned char)first;
  }
  rp.n_unique = 0, rp.n_bias = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long* q = jobs + 8 * j;
    const WbJob& jb = p.jobs[j];
    PNB_REQUIRE(job_nslot[j] >= 1, "wgrad_batch: internal (job without a segment)");
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
    rp.jobs[r].slot0[rp.jobs[r].n_ranges] = job_slot0[j], rp.jobs[r].nslot[rp.jobs[r].n_ranges] = job_nslot[j];
    ++rp.jobs[r].n_ranges;
    if (p.job_colsum[j]) {
      WbBiasJob& bj = rp.bias[rp.n_bias++];
      bj.db = reinterpret_cast<float*>(const_cast<void*>(db[j]));
      bj.Nw = jb.Nw, bj.slot0 = job_slot0[j], bj.nslot = job_nslot[j];
    }
  }
  p.partial = workspace;
  rp.partial = workspace;
  static const int max_stages = [] {
    const char* e = getenv("PNB_WB_MAXSTAGES");
    const int v = e != nullptr ? atoi(e) : kWbMaxStages;
    return v < 2 ? 2 : (v > kWbMaxStages ? kWbMaxStages : v);
  }();
  p.max_stages = max_stages;
  static const bool want_prof = getenv("PNB_WB_PROF") != nullptr;
  static long long* prof_dev = nullptr;
  if (want_prof && prof_dev == nullptr) cudaMalloc(&prof_dev, 2 * kNumSMs * sizeof(long long));
  p.prof = want_prof ? prof_dev : nullptr;
  const size_t smem_bytes = kSmemLimit;
  cudaError_t e = cudaFuncSetAttribute(wgrad_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) {
    set_error("wgrad_batch(smem attr)", e);
    return (int)e;
  }
  cudaStream_t st = as_stream(stream);
  wgrad_batch_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  int rc = finish("wgrad_batch");
  if (rc) return rc;
  if (want_prof) {  // per-CTA wall time and the jobs it worked on (stderr)
    static long long host[2 * kNumSMs];
    cudaStreamSynchronize(st);
    cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost);
    long long t0 = host[0], t1 = host[1];
    for (int c = 0; c < grid; ++c) t0 = host[2 * c] < t0 ? host[2 * c] : t0, t1 = host[2 * c + 1] > t1 ? host[2 * c + 1] : t1;
    fprintf(stderr, "[wgrad_batch] %d jobs, kernel %.1f us; per CTA: first job, us busy, MB\n", n_jobs, (t1 - t0) / 1e3);
    for (int c = 0; c < grid; ++c)
      fprintf(stderr, "  cta %3d job %2d  %7.1f us  %6.1f MB  boxes %d colsum %d\n", c, (int)p.cta_job0[c],
              (host[2 * c + 1] - host[2 * c]) / 1e3, (p.cut[c + 1] - p.cut[c]) * 64 / 1e6,
              p.jobs[p.cta_job0[c]].Nw / 64 + (p.jobs[p.cta_job0[c]].Kw + 63) / 64, (int)p.job_colsum[p.cta_job0[c]]);
  }
  wgrad_batch_reduce_kernel<<<dim3(32, rp.n_unique + 1), 256, 0, st>>>(rp);
  return finish("wgrad_batch_reduce");
}

// Test hook (host only, no GPU needed): the segment plan pnb_wgrad_batch uses for (M, jobs) - rows of
// {cta, job, first block, end block, slot}; returns the number of segments, or -1 if `max_segments` is too small.
extern "C" int pnb_wgrad_batch_plan(long long M, int n_jobs, const long long* jobs, long long* out_segments,
                                    int max_segments) {
  if (M <= 0 || n_jobs < 1 || n_jobs > kWbMaxJobs || jobs == nullptr || out_segments == nullptr) return -1;
  const long long blocks64 = (M + kWbRows - 1) / kWbRows;
  long long start[kWbMaxJobs + 1], cut[kNumSMs + 1];
  int cost[kWbMaxJobs], boxes[kWbMaxJobs];
  start[0] = 0;
  for (int j = 0; j < n_jobs; ++j) {
    cost[j] = wb_block_cost((int)jobs[8 * j + 4], (int)jobs[8 * j + 5], jobs[8 * j + 7] != 0);
    boxes[j] = (int)(jobs[8 * j + 4] / 64 + (jobs[8 * j + 5] + 63) / 64);
    start[j + 1] = start[j] + blocks64 * cost[j];
  }
  wb_build_cuts(start, cost, boxes, n_jobs, blocks64, cut);
  int n = 0;
  for (int c = 0; c < kNumSMs; ++c) {
    const long long lo = cut[c], hi = cut[c + 1];
    for (int j = 0; j < n_jobs && start[j] < hi; ++j) {
      long long b0, b1;
      wb_segment(lo, hi, start[j], cost[j], blocks64, &b0, &b1);
      if (b1 <= b0) continue;
      if (n >= max_segments) return -1;
      long long* o = out_segments + 5 * n;
      o[0] = c, o[1] = j, o[2] = b0, o[3] = b1, o[4] = n;
      ++n;
    }
  }
  return n;
}
