// K3-fused: the radiance/density MLP of models/pano_mip_nerf.py:78-114 (8x256 trunk with the skip connection,
// density / extra / view / colour heads), the density-Jacobian sweep that replaces vmap(jacrev)
// (pano_mip_nerf.py:295-302), the data-gradient chain of its backward pass and the adjoint (forward-mode) sweep of the
// Jacobian - each as ONE persistent kernel in which activations never leave the SM.
//
// Why: layer-by-layer GEMMs (gemm_tc.cu) stream every [M,256] activation through HBM and are bound by it at
// ~0.3 of the tensor roofline.  Here a CTA owns TWO 128-sample tiles at a time and ping-pongs between them:
//
//   warp 0      producer : streams the pre-swizzled bf16 weight tiles (cp.async.bulk, L2 -> SMEM ring of three 32 KB
//                          slots) and the IPE k-blocks of each tile (TMA tensor load) following a static plan of the
//                          ring (plan_ring): tile 1 walks the k-blocks of an op backwards so that it re-uses the
//                          slots tile 0 touched last (5 loads instead of 8 per layer and tile pair);
//   warp 1      MMA      : one thread issues tcgen05.mma (M=128, N<=256, K=16, bf16 -> fp32 TMEM).  A comes from the
//                          tile's activation buffer in SMEM (128B-swizzled K-major, 4 k-blocks of 64 columns), B from
//                          the ring, D is the tile's own 256-column TMEM accumulator.  The program alternates
//                          tile 0 / tile 1 op by op, so while the epilogue warps rewrite one tile's activations the
//                          tensor pipe is busy with the other tile;
//   warps 2..3  encoder  : (inference forward only) integrated positional encoding computed in the kernel: the 96
//                          features of the next tile pair (models/mip.py:394-428) are evaluated from the Gaussians
//                          (24 B per sample) into a per-CTA, double-buffered scratch that lives in L2 and is
//                          re-dirtied in place - it never reaches DRAM - from where the producer's TMA loads pick it
//                          up as before.  The [M,96] encoding array and its kernel disappear from the render path;
//   warps 4..11 epilogue : tcgen05.ld -> bias / ReLU / mask -> bf16 -> written IN PLACE into the tile's activation
//                          buffer as the next op's A operand; one mbarrier hand-shake per (op, tile) in each direction.
//
// ReLU sign bits go to a small global (L2-resident) bit-plane buffer: the Jacobian sweep
// a_{i-1} = relu'(h_{i-1}) * (a_i W_i) reads them back in the same kernel, the backward kernels read them instead of
// the activations.  With `acts` given, every activation is also written out with TMA stores straight from the
// activation buffer - that is what the weight-gradient GEMMs consume.
//
// Programs (compile-time schedules, one kernel instantiation each):
//   P_FWD   trunk + heads
//   P_FWDJ  trunk + heads + density-Jacobian sweep (d sigma / d enc)
//   P_BWD   d(raw_rgb), d(raw_sigma...) -> dz of every layer (and d enc), the dgrad chain of the backward pass
//   P_JADJ  u = J_ipe d_v -> q_i = relu'(h_i) * (q_{i-1} W_i^T): adjoint of the Jacobian sweep (second-order terms)
#include <utility>

#include "tc_common.cuh"

namespace pnb {
namespace fused {
using namespace pnb::tc;

constexpr int kTileM = 128;
constexpr int kWidth = 256, kEncDim = 96, kCondW = 128;
constexpr int kSlotBytes = 32768;  // ring slot: a weight tile of up to 256 rows x 64 bf16, or the two IPE k-blocks
constexpr int kKbBytes = 16384;    // one k-block: 128 rows x 64 bf16, 128B-swizzled
constexpr int kAbufBytes = 4 * kKbBytes;
constexpr int kFThreads = 384;  // producer, MMA, 2 encoder warps, 8 epilogue warps
constexpr int kMaxSteps = 96, kMaxOps = 24, kMaxPack = 176, kSlots = 3, kMaxLoads = 2 * 96 + 16;
constexpr int kNumParams = 12;  // weights (and biases) in state-dict order: layers 0..7, density, extra, view, colour
constexpr int kActPlanes = 18, kBwdPlanes = 10, kAdjPlanes = 8;
constexpr int kMaskPlanes = 9;  // ReLU sign bits: trunk layers 0..7, view layer
constexpr int kMaskWordsPerTile = kMaskPlanes * 8 * kTileM;
// Every CTA streams the same weight tiles in the same order; with a single copy all 148 SMs hammer the same few L2
// slices at the same time.  The packer therefore writes kReplicas copies of the blob at different addresses and
// CTA b reads copy b % replicas.
constexpr int kReplicas = 1;

// bias blob (fp32) layout
constexpr int kBiasHE = 2048, kBiasHD = 2304, kBiasC = 2320, kWDen = 2336, kWCol = kWDen + 16 * 256,
              kBiasFloats = kWCol + 4 * 128;

// Step flags.  The kernel consumes F_AENC only (A operand = the IPE tile in the ring instead of the activation buffer);
// the other three annotate the forward walk of a step list - which step needs / releases the IPE tile, which one starts
// the accumulation - and are re-derived per tile by plan_ring (tile 1 walks every op backwards).
enum : uint32_t { F_AENC = 1, F_LOADENC = 2, F_RELENC = 4, F_FIRST = 8, F_ABIAS = 16 };
// Bias as one more K = 16 MMA step (F_ABIAS): D += ones[128 x 16] * Bt[N x 16]^T with the bias of column n split into two
// bf16 terms (hi + mid + lo: all 24 mantissa bits) in row n of Bt and zeros elsewhere.  The per-column fp32 add (one FADD + half a
// uniform load per element, a third of the epilogue's instructions) leaves the epilogue warps, which are the
// bottleneck of these kernels; the tensor pipe pays 1/16 of an op for it.  Both operands are small un-swizzled K-major
// tiles (8-row x 16-byte core matrices) that travel together in one 12 KB blob entry: [Bt k 0..7 | Bt k 8..15 | ones].
// Because A is all ones the position of the three terms inside a row of Bt is irrelevant.
constexpr uint32_t kBiasTileBytes = 12288, kBiasOnesOff = 8192;
constexpr uint32_t kBiasB_LBO = 4096, kBiasB_SBO = 128, kBiasA_LBO = 2048, kBiasA_SBO = 128;
enum { P_FWD = 0, P_FWDJ = 1, P_BWD = 2, P_JADJ = 3, kNumProgs = 4 };

// One step = one ring slot = one weight tile of `nk16` K=16 MMAs.
struct Step {
  uint32_t blob_off;   // byte offset of the tile in the weight blob
  uint32_t bytes;      // tile bytes
  uint32_t idesc;      // tcgen05 instruction descriptor (M=128, N of this op, bf16 -> fp32)
  uint32_t acc_col;    // accumulator column inside the tile's 256-column TMEM region
  uint32_t a_off16;    // A operand: byte offset >> 4 from the activation buffer (or from the IPE slot with F_AENC)
  uint32_t b_kb16;     // byte stride >> 4 between the k-blocks of the weight tile inside the slot
  uint32_t flags;      // F_*
  uint32_t nk16;       // K=16 MMAs in this step
};
struct PackTile {      // one rows x 64 sub-tile: tile(r,c) = W[r0+r, c0+c] (or W[r0+c, c0+r] when transposed)
  uint32_t blob_off;
  int16_t param, transposed, r0, c0, vr, vc, rows, pad;
};
enum : uint8_t {
  E_RELU = 0,  // acc + bias -> ReLU (sign bits -> mask plane) -> abuf
  E_DEN,       // acc[0:16] + bias -> raw_den
  E_EXTRA,     // acc + bias -> abuf
  E_VIEW,      // acc + per-ray view-direction term -> ReLU -> abuf[0:128]
  E_COLOR,     // acc[128:144] + bias -> raw_rgb ; P_FWDJ: seed of the Jacobian sweep -> abuf
  E_MASK,      // acc * mask plane -> abuf
  E_GSKIP,     // acc[0:96] -> g_enc
  E_G0,        // acc[0:96] + g_enc -> g_enc
  E_BSEED,     // no MMA: relu'(hv) * (d_rgb W_col) -> abuf[0:128]
  E_LIN,       // acc -> abuf
  E_BDZ7       // (acc + d_den W_den) * mask plane 7 -> abuf
};
struct Op {
  int16_t s0, s1;      // steps [s0, s1)
  uint8_t epi, has_mma;
  int16_t acc_col;     // accumulator column the epilogue reads
  int16_t bias_off;
  int8_t mask_plane, save_plane, nunits, pad;
};
// Ring-slot plan of one step for one of the two tiles (static: see plan_ring below).
struct Use {
  int8_t w_slot, w_load, w_rel;  // weight tile: slot, wait for a fresh load first, release the slot afterwards
  int8_t e_slot, e_load, e_rel;  // IPE tile of this (op, tile) for F_AENC steps
  int8_t first, pad;             // first executed step of its op: the MMA overwrites the accumulator
};
struct LoadRec {                 // the producer's program: loads in the order the MMA warp needs them
  uint32_t blob_off, bytes;
  int8_t slot, is_enc, t, nkb;   // nkb: k-blocks stacked in the tile (0 for a bias tile)
  int16_t rows, pad;             // rows of the weight tile (N of the op)
};
struct Prog {
  Step steps[kMaxSteps];
  Op ops[kMaxOps];
  Use uses[2][kMaxSteps];
  LoadRec loads[kMaxLoads];
  int n_steps, n_ops, n_loads;
};
struct Sched {
  Prog prog[kNumProgs];
  PackTile pack[kMaxPack];
  int n_pack;
  uint32_t blob_bytes;
};
struct PackArgs {
  const float* w[kNumParams];
  const float* b[kNumParams];
  int ld[kNumParams];
  int C, n_tiles;
};

struct FusedParams {
  long long M, num_tiles, num_pairs;
  int S, C, nstages, save, debug, masks_per_tile, replicas;
  long long blob_stride;
  const uint8_t* wblob;
  const float* bblob;
  const float* row_bias;
  float* raw_den;
  float* raw_rgb;
  float* g_enc;          // P_FWDJ: d sigma / d enc out ; P_BWD: d L / d enc out (nullable)
  uint32_t* masks;       // ReLU sign bit-planes (nullable in P_FWD without `save`)
  const float* d_rgb;    // P_BWD inputs
  const float* d_den;
  __nv_bfloat16* planes;     // save != 0: activation planes [n][M][256] (same memory the TMA map covers)
  const float* means;        // in-kernel IPE (nullable): Gaussians [M,3] x 2 and the lowest degree
  const float* covs;
  __nv_bfloat16* enc_scratch;  // [grid][2 buffers][2 tiles][128][96] bf16, L2-resident
  int ipe_min_deg, vb_mod;     // vb_mod != 0: per-ray row bias of ray (m / S) % vb_mod (env rays share D directions)
  unsigned long long* prof;  // timing experiments: [grid][8] cycle counters (nullable)
};

// CTA-pair mode (template parameter C2, cta_group::2): two CTAs of a cluster - the two SMs of a TPC - run the
// schedule in lock-step; one thread of the LEADER (cluster rank 0) issues M = 256 MMAs that take each CTA's own 128-row
// activation tile as its half of A and HALF of every weight tile (N / 2 rows) from each CTA's ring as B, and write
// each CTA's accumulator into its own TMEM.  Per SM the weight traffic (L2 -> SMEM fills and the tensor cores' operand
// reads) halves, which is what bounds the single-CTA kernel (12 KB of operands per 128-cycle MMA = 96 of the 128 B/clk
// of shared-memory bandwidth, before the epilogue's stores and the ring fills).  Barrier topology: both producers'
// TMA loads count their bytes on the LEADER's `full` barriers (cp.async.bulk.tensor ... cta_group::2), both CTAs'
// epilogue warps arrive on the LEADER's `abuf_ready` (the peer through mapa / shared::cluster), and the leader's
// tcgen05.commit multicasts its arrivals to `acc_full` / `empty` of BOTH CTAs.
struct WMaps {                   // the weight blob as [bytes / 128][64] bf16, un-swizzled (tiles are pre-swizzled)
  CUtensorMap m[5];              // box rows 128, 64, 48, 16, 8
};
__host__ __device__ constexpr int wmap_index(int box_rows) {
  return box_rows == 128 ? 0 : box_rows == 64 ? 1 : box_rows == 48 ? 2 : box_rows == 16 ? 3 : box_rows == 8 ? 4 : -1;
}

struct FBarriers {
  uint64_t full[kSlots];
  uint64_t empty[kSlots];
  uint64_t abuf_ready[2];
  uint64_t acc_full[2];
  uint64_t enc_ready[2];  // in-kernel IPE: scratch buffer b holds the encodings of a tile pair
  uint64_t enc_free[2];   //                the MMA thread is done with the pair that used buffer b
  uint32_t tmem_base;
};

// ---------------------------------------------------------------------------------------------------------------
// schedules: the order of weight tiles == the order of MMA steps == the order of the producer's loads.
// Built at compile time: the MMA warp's program is fully unrolled from it (every descriptor offset, flag and
// barrier index is an immediate), the producer and epilogue warps read the same tables from constant memory.
// Identical weight tiles are shared between the programs (one blob).
// ---------------------------------------------------------------------------------------------------------------
struct Builder {
  Sched s{};
  // per-step tile identity for de-duplication
  int key[kMaxPack][8] = {};
  uint32_t key_off[kMaxPack] = {};
  int n_keys = 0;

  constexpr uint32_t tile(int param, int transposed, int r0, int c0, int vr, int vc, int rows, int nkb) {
    for (int i = 0; i < n_keys; ++i)
      if (key[i][0] == param && key[i][1] == transposed && key[i][2] == r0 && key[i][3] == c0 && key[i][4] == vr &&
          key[i][5] == vc && key[i][6] == rows && key[i][7] == nkb)
        return key_off[i];
    const uint32_t off = s.blob_bytes;
    key[n_keys][0] = param, key[n_keys][1] = transposed, key[n_keys][2] = r0, key[n_keys][3] = c0;
    key[n_keys][4] = vr, key[n_keys][5] = vc, key[n_keys][6] = rows, key[n_keys][7] = nkb;
    key_off[n_keys++] = off;
    for (int kb = 0; kb < nkb; ++kb) {  // consecutive 64-column k-blocks of the same rows
      PackTile pt{};
      pt.blob_off = s.blob_bytes;
      pt.param = (int16_t)param, pt.transposed = (int16_t)transposed;
      pt.r0 = (int16_t)(transposed ? r0 + 64 * kb : r0), pt.c0 = (int16_t)(transposed ? c0 : c0 + 64 * kb);
      pt.vr = (int16_t)vr, pt.vc = (int16_t)vc, pt.rows = (int16_t)rows;
      s.pack[s.n_pack++] = pt;
      s.blob_bytes += (uint32_t)rows * 128u;
    }
    return off;
  }
  // bias tile of parameter `param` (transposed == 2 marks it in the pack table)
  constexpr uint32_t bias_tile(int param) {
    for (int i = 0; i < n_keys; ++i)
      if (key[i][0] == param && key[i][1] == 2) return key_off[i];
    const uint32_t off = s.blob_bytes;
    key[n_keys][0] = param, key[n_keys][1] = 2;
    key_off[n_keys++] = off;
    PackTile pt{};
    pt.blob_off = off, pt.param = (int16_t)param, pt.transposed = 2, pt.rows = 256;
    s.pack[s.n_pack++] = pt;
    s.blob_bytes += kBiasTileBytes;
    return off;
  }
  constexpr void bias_step(int P, int param, int n_mma) {
    Prog& g = s.prog[P];
    Step st{};
    st.blob_off = bias_tile(param);
    st.bytes = kBiasTileBytes;
    st.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    st.acc_col = 0, st.a_off16 = kBiasOnesOff >> 4, st.b_kb16 = 0, st.flags = F_ABIAS, st.nk16 = 1;
    g.steps[g.n_steps++] = st;
  }
  constexpr void step(int P, int param, int transposed, int r0, int c0, int vr, int vc, int rows, int nkb, int n_mma,
                      int acc_col, int a_kb, int nk16, uint32_t flags) {
    Prog& g = s.prog[P];
    Step st{};
    st.blob_off = tile(param, transposed, r0, c0, vr, vc, rows, nkb);
    st.bytes = (uint32_t)(rows * 128 * nkb);
    st.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    st.acc_col = (uint32_t)acc_col;
    st.a_off16 = (uint32_t)(a_kb * kKbBytes) >> 4;
    st.b_kb16 = (uint32_t)(rows * 128) >> 4;
    st.flags = flags;
    st.nk16 = (uint32_t)nk16;
    g.steps[g.n_steps++] = st;
  }
  constexpr void op(int P, int s0, int epi, int acc_col, int bias_off, int mask_plane, int save_plane, int nunits) {
    Prog& g = s.prog[P];
    Op o{};
    o.s0 = (int16_t)s0, o.s1 = (int16_t)g.n_steps;
    o.epi = (uint8_t)epi, o.has_mma = (uint8_t)(g.n_steps > s0 ? 1 : 0);
    o.acc_col = (int16_t)acc_col, o.bias_off = (int16_t)bias_off;
    o.mask_plane = (int8_t)mask_plane, o.save_plane = (int8_t)save_plane, o.nunits = (int8_t)nunits;
    g.ops[g.n_ops++] = o;
  }
  // trunk layer i, forward direction (weights W_i[256, K]); `epi` = E_RELU (forward) or E_MASK (adjoint sweep)
  constexpr void trunk_layer(int P, int i, int epi, int save_plane) {
    const int s0 = s.prog[P].n_steps;
    if (i == 0) {
      step(P, 0, 0, 0, 0, 256, 64, 256, 1, 256, 0, 0, 4, F_AENC | F_LOADENC | F_FIRST);
      step(P, 0, 0, 0, 64, 256, 32, 256, 1, 256, 0, 1, 2, F_AENC | F_RELENC);
    } else {
      for (int kb = 0; kb < 4; ++kb) step(P, i, 0, 0, kb * 64, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
      if (i == 5) {  // skip connection: input = [h4 | enc]  (models/pano_mip_nerf.py:99-100)
        step(P, 5, 0, 0, 256, 256, 64, 256, 1, 256, 0, 0, 4, F_AENC | F_LOADENC);
        step(P, 5, 0, 0, 320, 256, 32, 256, 1, 256, 0, 1, 2, F_AENC | F_RELENC);
      }
    }
    if (epi == E_RELU) bias_step(P, i, 256);  // forward programs: + bias on the tensor pipe
    op(P, s0, epi, 0, i * 256, i, save_plane, 8);
  }
  // a_{i-1} = relu'(h_{i-1}) * (a_i W_i) for i = 7..1 with the transposed tiles, then the two contributions to the
  // gradient w.r.t. the encoding (through the skip connection and through layer 0)
  constexpr void input_gradient_chain(int P, int save_base) {
    for (int i = 7; i >= 1; --i) {
      if (i == 5) {  // skip connection: g_enc = a_5 W_5[:, 256:352]
        const int s0 = s.prog[P].n_steps;
        step(P, 5, 1, 0, 256, 96, 64, 96, 2, 96, 0, 0, 8, F_FIRST);
        step(P, 5, 1, 128, 256, 96, 64, 96, 2, 96, 0, 2, 8, 0);
        op(P, s0, E_GSKIP, 0, 0, -1, -1, 0);
      }
      const int s0 = s.prog[P].n_steps;
      for (int kb = 0; kb < 4; ++kb) step(P, i, 1, kb * 64, 0, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
      op(P, s0, E_MASK, 0, 0, i - 1, save_base + (i - 1), 8);
    }
    const int s0 = s.prog[P].n_steps;
    step(P, 0, 1, 0, 0, 96, 64, 96, 2, 96, 0, 0, 8, F_FIRST);
    step(P, 0, 1, 128, 0, 96, 64, 96, 2, 96, 0, 2, 8, 0);
    op(P, s0, E_G0, 0, 0, -1, -1, 0);
  }
};

// Static plan of the 3-slot weight ring for one pair of tiles.  Tile 0 walks the k-blocks of an op forwards, tile 1
// backwards, so the tiles that tile 0 touched last are still resident when tile 1 starts: 5 loads instead of 8 per
// 256x256 layer and pair.  Replacement is Belady's rule on the known future (dead tiles first, oldest first), every
// slot is released at its last use so the producer can refill it two steps ahead, and all slots are free again at
// the end of the pair (the plan repeats verbatim for the next pair).
#ifndef PNB_RING_MODE
#define PNB_RING_MODE 2  // 0: FIFO, no reuse, both tiles forwards; 1: reuse, both forwards; 2: reuse, tile 1 backwards
#endif
constexpr void plan_ring(Prog& g) {
  // events in execution order
  int ev_op[2 * kMaxSteps] = {}, ev_t[2 * kMaxSteps] = {}, ev_s[2 * kMaxSteps] = {};
  int n_ev = 0;
  for (int o = 0; o < g.n_ops; ++o) {
    const int s0 = g.ops[o].s0, s1 = g.ops[o].s1;
    for (int t = 0; t < 2; ++t)
      for (int k = 0; k < s1 - s0; ++k) {
        const int sidx = (t == 0 || PNB_RING_MODE < 2) ? s0 + k : s1 - 1 - k;
        ev_op[n_ev] = o, ev_t[n_ev] = t, ev_s[n_ev] = sidx;
        g.uses[t][sidx] = Use{};
        g.uses[t][sidx].first = (int8_t)(k == 0);
        ++n_ev;
      }
  }
  // content ids: weights = blob offset + 1 (> 0); IPE tile of (op, t) = -(2 * op + t + 1)
  auto wid = [&](int e) {
    return PNB_RING_MODE == 0 ? (long long)(e + 1) * (1ll << 32) : (long long)g.steps[ev_s[e]].blob_off + 1;
  };
  auto eid = [&](int e) {
    return (g.steps[ev_s[e]].flags & F_AENC) ? -(long long)(2 * ev_op[e] + ev_t[e] + 1) : 0ll;
  };
  long long content[kSlots] = {};
  int last_ev[kSlots] = {}, last_kind[kSlots] = {};  // last use of the slot's content: event, 0 = weight / 1 = IPE
  for (int k = 0; k < kSlots; ++k) last_ev[k] = -1;
  g.n_loads = 0;
  for (int e = 0; e < n_ev; ++e) {
    Use& u = g.uses[ev_t[e]][ev_s[e]];
    const long long need[2] = {eid(e), wid(e)};
    int slot_of[2] = {-1, -1};
    for (int n = 0; n < 2; ++n) {
      if (need[n] == 0) continue;
      int slot = -1;
      for (int k = 0; k < kSlots; ++k)
        if (content[k] == need[n]) slot = k;
      bool load = false;
      if (slot < 0) {
        load = true;
        // victim: empty slot, else a dead tile (no later use; oldest last use first), else the farthest next use
        int best = -1;
        long long best_key = -1;
        for (int k = 0; k < kSlots; ++k) {
          if (k == slot_of[0]) continue;  // never the IPE tile of this very step
          long long key = 0;
          if (content[k] == 0) {
            key = (1ll << 40);
          } else {
            int next = -1;
            for (int f = e; f < n_ev && next < 0; ++f)
              if (wid(f) == content[k] || eid(f) == content[k]) next = f;
            key = next < 0 ? (1ll << 30) - last_ev[k] : next;
          }
          if (key > best_key) best_key = key, best = k;
        }
        slot = best;
        if (content[slot] != 0) {  // the previous tenant leaves at its last use
          Use& pu = g.uses[ev_t[last_ev[slot]]][ev_s[last_ev[slot]]];
          if (last_kind[slot] == 0) pu.w_rel = 1;
          else pu.e_rel = 1;
        }
        content[slot] = need[n];
        LoadRec& l = g.loads[g.n_loads++];
        l.slot = (int8_t)slot, l.is_enc = (int8_t)(n == 0), l.t = (int8_t)ev_t[e];
        l.blob_off = g.steps[ev_s[e]].blob_off, l.bytes = g.steps[ev_s[e]].bytes;
        if ((g.steps[ev_s[e]].flags & F_ABIAS) != 0 || g.steps[ev_s[e]].b_kb16 == 0) {
          l.rows = 0, l.nkb = 0;
        } else {
          l.rows = (int16_t)(g.steps[ev_s[e]].b_kb16 / 8);
          l.nkb = (int8_t)(l.bytes / ((uint32_t)l.rows * 128u));
        }
      }
      slot_of[n] = slot;
      last_ev[slot] = e, last_kind[slot] = n == 0 ? 1 : 0;
      if (n == 0) u.e_slot = (int8_t)slot, u.e_load = (int8_t)load;
      else u.w_slot = (int8_t)slot, u.w_load = (int8_t)load;
    }
  }
  for (int k = 0; k < kSlots; ++k)
    if (content[k] != 0) {
      Use& pu = g.uses[ev_t[last_ev[k]]][ev_s[last_ev[k]]];
      if (last_kind[k] == 0) pu.w_rel = 1;
      else pu.e_rel = 1;
    }
}

constexpr Sched make_sched() {
  Builder b{};
  const int W_DEN = 8, W_EXTRA = 9, W_VIEW = 10, W_COL = 11;
  for (int P = P_FWD; P <= P_FWDJ; ++P) {
    for (int i = 0; i < 8; ++i) b.trunk_layer(P, i, E_RELU, i);
    // density head first (its 16 columns are read out before the extra layer overwrites the accumulator)
    int s0 = b.s.prog[P].n_steps;
    b.step(P, W_DEN, 0, 0, 0, 16, 64, 16, 4, 16, 0, 0, 16, F_FIRST);
    b.op(P, s0, E_DEN, 0, kBiasHD, -1, -1, 0);
    s0 = b.s.prog[P].n_steps;
    for (int kb = 0; kb < 4; ++kb) b.step(P, W_EXTRA, 0, 0, kb * 64, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
    b.bias_step(P, W_EXTRA, 256);
    b.op(P, s0, E_EXTRA, 0, kBiasHE, -1, 8, 8);
    // view layer, bottleneck columns (the view-direction columns are the per-ray row bias)
    s0 = b.s.prog[P].n_steps;
    for (int h = 0; h < 2; ++h) b.step(P, W_VIEW, 0, 0, h * 128, 128, 64, 128, 2, 128, 0, 2 * h, 8, h == 0 ? F_FIRST : 0);
    b.op(P, s0, E_VIEW, 0, 0, 8, 9, 4);
    s0 = b.s.prog[P].n_steps;
    b.step(P, W_COL, 0, 0, 0, 16, 64, 16, 2, 16, 128, 0, 8, F_FIRST);
    b.op(P, s0, E_COLOR, 128, kBiasC, 7, 17, 8);
    if (P == P_FWDJ) b.input_gradient_chain(P, 10);
  }
  {  // ---- backward: dgrad chain -------------------------------------------------------------------------------
    const int P = P_BWD;
    b.op(P, 0, E_BSEED, 0, 0, 8, 0, 4);  // dzv = relu'(hv) * (d_rgb W_col)
    int s0 = b.s.prog[P].n_steps;        // d_bott = dzv W_view[:, :256]
    for (int kb = 0; kb < 2; ++kb) b.step(P, W_VIEW, 1, kb * 64, 0, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
    b.op(P, s0, E_LIN, 0, 0, -1, 1, 8);
    s0 = b.s.prog[P].n_steps;            // dz_7 = relu'(h_7) * (d_bott W_extra + d_den W_den)
    for (int kb = 0; kb < 4; ++kb) b.step(P, W_EXTRA, 1, kb * 64, 0, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
    b.op(P, s0, E_BDZ7, 0, 0, 7, 2, 8);
    // dz_{i-1} -> plane 2 + (7 - (i-1)) = 9 - (i-1)
    for (int i = 7; i >= 1; --i) {
      if (i == 5) {
        const int g0 = b.s.prog[P].n_steps;
        b.step(P, 5, 1, 0, 256, 96, 64, 96, 2, 96, 0, 0, 8, F_FIRST);
        b.step(P, 5, 1, 128, 256, 96, 64, 96, 2, 96, 0, 2, 8, 0);
        b.op(P, g0, E_GSKIP, 0, 0, -1, -1, 0);
      }
      const int j0 = b.s.prog[P].n_steps;
      for (int kb = 0; kb < 4; ++kb) b.step(P, i, 1, kb * 64, 0, 256, 64, 256, 1, 256, 0, kb, 4, kb == 0 ? F_FIRST : 0);
      b.op(P, j0, E_MASK, 0, 0, i - 1, 9 - (i - 1), 8);
    }
    const int g0 = b.s.prog[P].n_steps;
    b.step(P, 0, 1, 0, 0, 96, 64, 96, 2, 96, 0, 0, 8, F_FIRST);
    b.step(P, 0, 1, 128, 0, 96, 64, 96, 2, 96, 0, 2, 8, 0);
    b.op(P, g0, E_G0, 0, 0, -1, -1, 0);
  }
  for (int i = 0; i < 8; ++i) b.trunk_layer(P_JADJ, i, E_MASK, i);
  for (int P = 0; P < kNumProgs; ++P) plan_ring(b.s.prog[P]);
  return b.s;
}
constexpr Sched kSched = make_sched();
static_assert(kSched.n_pack <= kMaxPack, "pack table");
static_assert(kSched.prog[P_FWDJ].n_steps <= kMaxSteps && kSched.prog[P_FWDJ].n_ops <= kMaxOps, "schedule tables");
static_assert(kSched.prog[P_BWD].n_steps <= kMaxSteps && kSched.prog[P_BWD].n_ops <= kMaxOps, "schedule tables");
static_assert(kSched.prog[P_FWDJ].n_loads <= kMaxLoads, "load table");
static_assert(1024 + 2 * kAbufBytes + kSlots * kSlotBytes + sizeof(FBarriers) <= (size_t)kSmemLimit, "shared memory budget");

// Bias blob staged in the constant bank before every launch (stream-ordered device-to-device copy): the epilogue's
// per-column addends are uniform across a warp, so they come through the constant cache / uniform datapath instead of
// exposing a global-load latency in front of every FADD.
__constant__ float c_bblob[kBiasFloats];

// The producer and the epilogue warps walk the same schedule at run time from constant memory.
__constant__ Prog c_prog[kNumProgs] = {kSched.prog[0], kSched.prog[1], kSched.prog[2], kSched.prog[3]};

// ---------------------------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> bf16 tiles in the 128B-swizzled K-major image tcgen05 reads from shared memory
// ---------------------------------------------------------------------------------------------------------------
struct PackTable {
  PackTile t[kMaxPack];
};
constexpr PackTable make_pack_table() {
  PackTable pt{};
  for (int i = 0; i < kSched.n_pack; ++i) pt.t[i] = kSched.pack[i];
  return pt;
}
__constant__ PackTable c_pack = make_pack_table();

__global__ void pack_tiles_kernel(const PackArgs a, uint8_t* __restrict__ wblob, long long blob_stride) {
  PackTile t = c_pack.t[blockIdx.x];
  wblob += (size_t)blockIdx.y * blob_stride;
  if (t.transposed == 2) {  // bias tile: [Bt chunk 0: (hi, lo, 0 x 6) per row | Bt chunk 1: zeros | ones]
    const float* bsrc = a.b[t.param];
    uint4* dst = reinterpret_cast<uint4*>(wblob + t.blob_off);
    for (int i = threadIdx.x; i < (int)(kBiasTileBytes / 16); i += blockDim.x) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (i < 256) {  // row n = i: core matrix n / 8 at n / 8 * 128 bytes, row n % 8 inside it
        // three bf16 terms carry all 24 mantissa bits of the fp32 bias (hi + mid + lo == bias exactly unless the
        // last term underflows), so the pre-activations - and with them the ReLU pattern the Jacobian sweep
        // depends on - equal those of an fp32 bias add to the last bit or so
        const float bv = bsrc[i];
        const __nv_bfloat16 hi = __float2bfloat16_rn(bv);
        const float r1 = bv - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
        const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
        v.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
        v.y = (uint32_t)__bfloat16_as_ushort(lo);
      } else if (i >= (int)(kBiasOnesOff / 16)) {
        v = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      }
      dst[i] = v;
    }
    return;
  }
  if (t.param == 8) t.vr = (int16_t)(t.transposed ? t.vr : a.C);  // density head: C valid rows
  if (t.param == 11) t.vr = 3;                                   // colour head: 3 valid rows
  const float* W = a.w[t.param];
  const int ld = a.ld[t.param];
  const int chunks = t.rows * 8;  // 16-byte chunks (8 bf16)
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
    const int r = i >> 3, j = i & 7;
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      float v[2];
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int c = j * 8 + e + d;
        float x = 0.f;
        if (r < t.vr && c < t.vc) x = t.transposed ? W[(size_t)(t.r0 + c) * ld + t.c0 + r] : W[(size_t)(t.r0 + r) * ld + t.c0 + c];
        v[d] = x;
      }
      h[e >> 1] = __floats2bfloat162_rn(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(wblob + t.blob_off + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(h);
  }
}

__global__ void pack_bias_kernel(const PackArgs a, float* __restrict__ bblob) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kBiasFloats; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (i < 2048) v = a.b[i >> 8][i & 255];
    else if (i < kBiasHD) v = a.b[9][i - kBiasHE];
    else if (i < kBiasC) v = (i - kBiasHD < a.C) ? a.b[8][i - kBiasHD] : 0.f;
    else if (i < kWDen) v = (i - kBiasC < 3) ? a.b[11][i - kBiasC] : 0.f;
    else if (i < kWCol) {  // density head weights [16][256]; row 0 = sigma row, the seed of the Jacobian sweep
      const int c = (i - kWDen) >> 8, k = (i - kWDen) & 255;
      v = c < a.C ? a.w[8][c * 256 + k] : 0.f;
    } else {               // colour head weights [4][128]
      const int c = (i - kWCol) >> 7, k = (i - kWCol) & 127;
      v = c < 3 ? a.w[11][c * 128 + k] : 0.f;
    }
    bblob[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
// L2 policies: the 2.2 MB weight blob is re-streamed by every CTA for every tile pair and must stay in L2
// (evict_last), while the saved activation planes are written once and read milliseconds later by the weight-gradient
// kernel (evict_first) - without the hints GBs of planes flush the weights out of L2 in the training variants.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2,
                                             uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32u(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, 128B-swizzled operand descriptor from (shared address >> 4): LBO = 16 B (unused), SBO = 1024 B, version 1
__device__ __forceinline__ uint64_t desc_from16(uint32_t addr16) {
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)kHi << 32) | (uint64_t)((addr16 & 0x3FFFu) | (1u << 16));
}

// Integrated positional encoding of one sample (models/mip.py:394-428, 355-361) written as 96 bf16 to `dst`
// (12 x 16-byte stores).  Same arithmetic as ipe_fwd_tile_kernel in rays.cu - exact fixed-point phase of 2^l * mean,
// SFU sin / cos, TwoSum-corrected second half that reproduces sin(fl32(y + fl32(pi/2))) - so the in-kernel encoding is
// bit-identical to the stand-alone kernel's bf16 output.
__device__ __forceinline__ void encode_row(const float* __restrict__ means, const float* __restrict__ covs, long long m,
                                           int min_deg, __nv_bfloat16* __restrict__ dst) {
  constexpr float kHalfPiF = 1.57079632679489661923f;
  uint32_t hi[3], lo[3];
  float mean[3], cov[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mean[c] = __ldg(means + 3 * m + c), cov[c] = __ldg(covs + 3 * m + c);
    double u = (double)mean[c] * 0.15915494309189534561;  // turns
    u -= floor(u);
    const unsigned long long U = (unsigned long long)(u * 18446744073709551616.0);
    hi[c] = (uint32_t)(U >> 32), lo[c] = (uint32_t)U;
  }
  uint32_t w[48];  // packed bf16 pairs: features 2k, 2k+1
  float f[6];      // sin then cos of (l, c = 0..2), flushed every two degrees
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    const int sh = min_deg + l;
    const float sc = __uint_as_float((uint32_t)(127 + sh) << 23);  // 2^sh
    float sn3[3], cs3[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t ph = __funnelshift_l(lo[c], hi[c], sh);
      const float r = (float)(int)ph * 1.46291807926715968e-9f;
      const float sn = __sinf(r), cs = __cosf(r);
      const float y = __fmul_rn(mean[c], sc);
      const float kl = -1.44269504088896341f * __uint_as_float((uint32_t)(127 + 2 * sh - 1) << 23);
      float ex;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(__fmul_rn(cov[c], kl)));
      const float z = __fadd_rn(y, kHalfPiF);
      const float bb = __fsub_rn(z, y);
      const float err = __fadd_rn(__fsub_rn(y, __fsub_rn(z, bb)), __fsub_rn(kHalfPiF, bb));
      const float eps = __fsub_rn(4.37113900018624283e-8f, err);
      const float e2 = __fmul_rn(eps, eps);
      const float ce = fmaf(-0.5f, e2, 1.0f);
      const float se = __fmul_rn(eps, fmaf(-0.16666667f, e2, 1.0f));
      const float c2 = fmaf(cs, ce, -__fmul_rn(sn, se));
      sn3[c] = __fmul_rn(ex, sn), cs3[c] = __fmul_rn(ex, c2);
    }
    // feature index l*3 + c (sin) and 48 + l*3 + c (cos); two degrees = 6 features = 3 packed words each
    if ((l & 1) == 0) {
      f[0] = sn3[0], f[1] = sn3[1], f[2] = sn3[2];
      f[3] = cs3[0], f[4] = cs3[1], f[5] = cs3[2];
    } else {
      const int k = (l >> 1) * 3;  // word index of feature (l-1)*3
      __nv_bfloat162 t;
      t = __floats2bfloat162_rn(f[0], f[1]); w[k] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(f[2], sn3[0]); w[k + 1] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(sn3[1], sn3[2]); w[k + 2] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(f[3], f[4]); w[24 + k] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(f[5], cs3[0]); w[24 + k + 1] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(cs3[1], cs3[2]); w[24 + k + 2] = *reinterpret_cast<uint32_t*>(&t);
    }
  }
  uint4* o = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int k = 0; k < 12; ++k) o[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

// K-major, un-swizzled operand descriptor (8-row x 16-byte core matrices; LBO: K direction, SBO: M / N direction)
__device__ __forceinline__ uint64_t desc_nosw16(uint32_t addr16, uint32_t lbo, uint32_t sbo) {
  const uint32_t hi = (sbo >> 4) | (1u << 14);
  return ((uint64_t)hi << 32) | (uint64_t)((addr16 & 0x3FFFu) | ((lbo >> 4) << 16));
}

struct EpiCtx {
  uint32_t ab_cluster;      // pair mode, peer CTA: shared::cluster address of the LEADER's abuf_ready of this tile (else 0)
  uint8_t* abuf;            // this tile's activation buffer
  uint32_t* mask;           // this tile's sign bit-planes [plane][unit][row] (nullptr: not kept)
  uint64_t* acc_full;
  uint64_t* abuf_ready;
  const CUtensorMap* tmActs;
  uint32_t tacc;            // TMEM address of this tile's accumulator, lane quadrant applied
  uint32_t acc_parity;
  uint32_t mbits[4];        // sign-bit words of this thread's units for this op (prefetched one op-tile ahead)
  long long m;              // global sample row of this thread
  int q, hf, lane, row, tile_row0;
  uint64_t store_policy;    // L2 evict_first for the saved planes
  bool row_ok, save, skip, direct, tma, tmaw;  // save: this tile stores its planes; tma / tmaw / direct: how
};

enum { M_BIAS_RELU = 0, M_BIAS, M_ROWBIAS_RELU, M_MASK, M_LIN, M_SEED, M_BSEED, M_BDZ7 };

// "this warp is done with the tile's accumulator and activation buffer": the MMA issuer waits on the leader's barrier
__device__ __forceinline__ void epi_arrive(const EpiCtx& c) {
  if (c.ab_cluster != 0) mbar_arrive_cluster(c.ab_cluster);
  else mbar_arrive(c.abuf_ready);
}

// Rewrite this warp's part of the tile's activation buffer (the next op's A operand) from the accumulator.
// The two warps of a TMEM lane quadrant split the 32-column units: warp hf handles units [hf*n/2, (hf+1)*n/2), i.e.
// whole 64-column k-blocks, so each warp can TMA-store its 32-row x 64-column boxes on its own.
//   unit u covers columns [32u, 32u+32) = k-block u/2, 16-byte chunks (u&1)*4 .. +3 of the 128-byte row.
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// HF (which of the two warps of the lane quadrant) is a template parameter so that every column offset - and with it
// every constant-bank address of the bias - is an immediate on top of a warp-uniform base.
template <int MODE, int HF>
__device__ __forceinline__ void rewrite_abuf_hf(const EpiCtx& c, const FusedParams& p, const Op& op, const float* aux,
                                                int coff) {
  constexpr bool kReadsAcc = MODE != M_SEED && MODE != M_BSEED;
  constexpr bool kAppliesMask = MODE == M_MASK || MODE == M_SEED || MODE == M_BSEED || MODE == M_BDZ7;
  constexpr bool kWritesMask = MODE == M_BIAS_RELU || MODE == M_ROWBIAS_RELU;
  // (trunk / extra-layer biases arrive through the MMA: F_ABIAS; only the per-ray view term is added here)
  constexpr int n_mine = (MODE == M_ROWBIAS_RELU || MODE == M_BSEED) ? 2 : 4;  // units per warp (op.nunits / 2)
  constexpr int ub = HF * n_mine;
  // sign-bit words of this thread's units: bit (31 - j) of word u <-> column 32u + j is positive.
  uint32_t bw[n_mine];
#pragma unroll
  for (int i = 0; i < n_mine; ++i) bw[i] = c.mbits[i];
  float rb[MODE == M_ROWBIAS_RELU ? n_mine : 1][32];
  if (MODE == M_ROWBIAS_RELU) {
#pragma unroll
    for (int i = 0; i < n_mine; ++i) {
      const float4* src = reinterpret_cast<const float4*>(aux) + (ub + i) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 t = __ldg(src + j);
        rb[i][4 * j] = t.x, rb[i][4 * j + 1] = t.y, rb[i][4 * j + 2] = t.z, rb[i][4 * j + 3] = t.w;
      }
    }
  }
  float d0 = 0.f, d1 = 0.f, d2 = 0.f;
  if (MODE == M_BSEED && c.row_ok) d0 = __ldg(p.d_rgb + c.m * 3), d1 = __ldg(p.d_rgb + c.m * 3 + 1), d2 = __ldg(p.d_rgb + c.m * 3 + 2);
  if (op.has_mma != 2) {  // (2: the caller has already waited - Jacobian seed after the colour head)
    mbar_wait(c.acc_full, c.acc_parity);
    tc_fence_after();
  }
  if (c.skip) {  // timing experiment: the epilogue costs nothing
    tc_fence_before();
    __syncwarp();
    if (c.lane == 0) epi_arrive(c);
    return;
  }
  uint32_t r[2][32];
  if (kReadsAcc) tmem_ld32u(c.tacc + op.acc_col + ub * 32, r[0]);
  if (c.tma) {
    // The four warps of this half (one per TMEM lane quadrant) store whole 128-row k-blocks together; warp q = 0
    // issues one bulk group per (op, tile).  Its group that still reads this tile's buffer (the op before the other
    // tile's) must be done; only the other tile's newer group may stay pending.
    if (c.q == 0 && c.lane == 0) bulk_wait_read<1>();
    named_bar_sync(1 + HF, 128);
  } else if (c.tmaw) {
    // per-warp variant: every warp stores its own 32 rows (4 KB boxes) and waits only for its own earlier group -
    // no rendezvous between the warps of a half
    if (c.lane == 0) bulk_wait_read<1>();
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < n_mine; ++i) {
    const int u = ub + i;
    float x[32];
    if (MODE == M_SEED) {  // sigma row of the density head
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = c_bblob[coff + u * 32 + j];
    }
    if (MODE == M_BSEED) {  // d_rgb W_col: 3 x 32 FMAs
      const float* w = c_bblob + kWCol + u * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaf(d2, w[256 + j], fmaf(d1, w[128 + j], d0 * w[j]));
    }
    if (kReadsAcc) {
      tmem_wait_ld();
      if (i + 1 < n_mine) tmem_ld32u(c.tacc + op.acc_col + (u + 1) * 32, r[(i + 1) & 1]);  // next unit in flight
      if (MODE == M_ROWBIAS_RELU) {  // + per-ray view-direction term (prefetched above)
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[i & 1][j]) + rb[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(r[i & 1][j]);
      }
    }
    if (MODE == M_BDZ7) {  // + d_den W_den (C x 32 FMAs; the density head is too narrow for its own MMA pass)
      for (int ch = 0; ch < p.C; ++ch) {
        const float* w = c_bblob + kWDen + ch * 256 + u * 32;
        const float d = c.row_ok ? __ldg(p.d_den + c.m * p.C + ch) : 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaf(d, w[j], x[j]);
      }
    }
    if (kWritesMask) {
      if (c.mask != nullptr) {  // funnel shifts collect the sign bits (positive <-> sign bit clear), 4 independent chains
        uint32_t sg[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int g = 0; g < 4; ++g) sg[g] = __funnelshift_l(__float_as_uint(x[8 * g + j]), sg[g], 1);
        }
        const uint32_t word = ~((sg[0] << 24) | (sg[1] << 16) | (sg[2] << 8) | sg[3]);
        uint32_t* mdst = c.mask + (op.mask_plane * 8 + u) * kTileM + c.row;
        if (p.masks_per_tile) __stcs(mdst, word);  // training: next read by the backward kernels, stream past L2
        else *mdst = word;                         // inference: read back by this kernel's Jacobian sweep
      }
    }
    if (kAppliesMask) {
      const uint32_t b = bw[i];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = (b & (0x80000000u >> j)) ? x[j] : 0.f;
    }
    uint32_t h[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (MODE == M_BIAS_RELU || MODE == M_ROWBIAS_RELU) {  // ReLU inside the conversion (F2FP.RELU)
        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(h[k]) : "f"(x[2 * k + 1]), "f"(x[2 * k]));
      } else {
        __nv_bfloat162 t = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);
        h[k] = *reinterpret_cast<uint32_t*>(&t);
      }
    }
    const uint32_t dst = smem_u32(c.abuf) + (u >> 1) * kKbBytes + c.row * 128;
    const int jb = (u & 1) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(dst + (((jb + j) ^ (c.row & 7)) << 4), h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    if (c.save && c.direct && op.save_plane >= 0 && c.row_ok) {
      // straight from the registers: this thread's 32 columns are 64 contiguous bytes of its row (two full sectors)
      uint4* gdst = reinterpret_cast<uint4*>(p.planes + ((size_t)op.save_plane * (size_t)p.M + (size_t)c.m) * kWidth + u * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) __stcs(gdst + j, make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
    }
  }
  fence_async_smem();
  if (c.tma) {
    // this half's k-blocks are in place for all 128 rows: one rendezvous, then the 128 x 64 boxes (16 KB each) leave
    // as one bulk group.  (Always committed - also for ops without a plane and for the phantom tile of an odd tail -
    // so that the one-group-per-(op, tile) count the wait above relies on stays exact.)
    named_bar_sync(1 + HF, 128);
    if (c.q == 0 && c.lane == 0) {
      if (c.save && op.save_plane >= 0) {
#pragma unroll
        for (int kb = ub / 2; kb < (ub + n_mine) / 2; ++kb)
          tma_store_3d(c.tmActs, c.abuf + kb * kKbBytes, kb * 64, c.tile_row0, op.save_plane, c.store_policy);
      }
      bulk_commit();
    }
  }
  if (c.tmaw) {
    __syncwarp();
    if (c.lane == 0) {
      if (c.save && op.save_plane >= 0) {
#pragma unroll
        for (int kb = ub / 2; kb < (ub + n_mine) / 2; ++kb)
          tma_store_3d(c.tmActs, c.abuf + kb * kKbBytes + c.q * 4096, kb * 64, c.tile_row0 + c.q * 32, op.save_plane,
                       c.store_policy);
      }
      bulk_commit();
    }
  }
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) epi_arrive(c);
}

template <int MODE>
__device__ __forceinline__ void rewrite_abuf(const EpiCtx& c, const FusedParams& p, const Op& op, const float* aux,
                                             int coff) {
  if (c.hf == 0) rewrite_abuf_hf<MODE, 0>(c, p, op, aux, coff);
  else rewrite_abuf_hf<MODE, 1>(c, p, op, aux, coff);
}

// Epilogues that only read a few accumulator columns (heads, encoding gradients) and leave the buffer alone.
__device__ __forceinline__ void epi_wait(const EpiCtx& c) {
  mbar_wait(c.acc_full, c.acc_parity);
  tc_fence_after();
}
__device__ __forceinline__ void epi_done(const EpiCtx& c) {
  tc_fence_before();
  __syncwarp();
  if (c.lane == 0) epi_arrive(c);
}

// ---------------------------------------------------------------------------------------------------------------
// MMA program
// ---------------------------------------------------------------------------------------------------------------
struct MmaCtx {
  FBarriers* bars;
  uint32_t full_ph, ab_ph;  // per-slot / per-tile phase bits
  uint32_t abuf16, ring16, tmem_base;
  bool prof;
  long long t_ab, t_full;
};

template <int SLOT>
__device__ __forceinline__ void mma_wait_slot(MmaCtx& c) {
  if (c.prof) {
    const long long t0 = clock64();
    mbar_wait(&c.bars->full[SLOT], (c.full_ph >> SLOT) & 1u);
    c.t_full += clock64() - t0;
  } else {
    mbar_wait(&c.bars->full[SLOT], (c.full_ph >> SLOT) & 1u);
  }
  c.full_ph ^= 1u << SLOT;
}

template <bool C2>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  if constexpr (C2) umma_f16_2cta(d, a, b, (idesc & ~(0x1Fu << 24)) | ((256u >> 4) << 24), accum);
  else umma_f16(d, a, b, idesc, accum);
}
template <bool C2>
__device__ __forceinline__ void commit_any(uint64_t* bar) {
  if constexpr (C2) umma_commit_2cta(bar, (uint16_t)3);
  else umma_commit(bar);
}

template <int P, int S, int T, bool C2>
__device__ __forceinline__ void mma_step(MmaCtx& c) {
  constexpr Step st = kSched.prog[P].steps[S];
  constexpr Use u = kSched.prog[P].uses[T][S];
  if constexpr ((st.flags & F_AENC) != 0 && u.e_load != 0) mma_wait_slot<u.e_slot>(c);
  if constexpr (u.w_load != 0) mma_wait_slot<u.w_slot>(c);
  tc_fence_after();
  const uint32_t a16 = (((st.flags & F_AENC) != 0) ? c.ring16 + (uint32_t)u.e_slot * (kSlotBytes >> 4)
                                                  : c.abuf16 + (uint32_t)T * (kAbufBytes >> 4)) +
                       st.a_off16;
  const uint32_t b16 = c.ring16 + (uint32_t)u.w_slot * (kSlotBytes >> 4);
  const uint32_t d_tmem = c.tmem_base + (uint32_t)T * 256u + st.acc_col;
  if constexpr ((st.flags & F_ABIAS) != 0) {
    // pair mode: the slot holds this CTA's 128 rows of Bt ([k 0..7: 2 KB | k 8..15: 2 KB]) and the ones tile at 4 KB
    constexpr uint32_t a_off = C2 ? (4096u >> 4) : st.a_off16;
    constexpr uint32_t b_lbo = C2 ? 2048u : kBiasB_LBO;
    umma_any<C2>(d_tmem, desc_nosw16(b16 + a_off, kBiasA_LBO, kBiasA_SBO), desc_nosw16(b16, b_lbo, kBiasB_SBO),
                 st.idesc, u.first != 0 ? 0u : 1u);
  } else
#pragma unroll
  for (int k = 0; k < (int)st.nk16; ++k) {
    const uint32_t ak = (uint32_t)((k >> 2) * (kKbBytes >> 4) + (k & 3) * 2);
    const uint32_t bk = (uint32_t)(k >> 2) * (C2 ? st.b_kb16 >> 1 : st.b_kb16) + (uint32_t)(k & 3) * 2;
    umma_any<C2>(d_tmem, desc_from16(a16 + ak), desc_from16(b16 + bk), st.idesc, (k == 0 && u.first != 0) ? 0u : 1u);
  }
  if constexpr (u.w_rel != 0) commit_any<C2>(&c.bars->empty[u.w_slot]);
  if constexpr ((st.flags & F_AENC) != 0 && u.e_rel != 0) commit_any<C2>(&c.bars->empty[u.e_slot]);
}

// tile 0 walks the steps of an op forwards, tile 1 backwards (see plan_ring)
template <int P, int S0, int N, int T, bool C2, int... K>
__device__ __forceinline__ void mma_steps(MmaCtx& c, std::integer_sequence<int, K...>) {
  (mma_step<P, ((T == 0 || PNB_RING_MODE < 2) ? S0 + K : S0 + N - 1 - K), T, C2>(c), ...);
}

template <int P, int OP, int T, bool C2>
__device__ __forceinline__ void mma_op_tile(MmaCtx& c) {
  constexpr Op op = kSched.prog[P].ops[OP];
  // the tile's previous epilogue has drained the accumulator and rewritten the activation buffer
  if (c.prof) {
    const long long t0 = clock64();
    mbar_wait(&c.bars->abuf_ready[T], (c.ab_ph >> T) & 1u);
    c.t_ab += clock64() - t0;
  } else {
    mbar_wait(&c.bars->abuf_ready[T], (c.ab_ph >> T) & 1u);
  }
  c.ab_ph ^= 1u << T;
  if constexpr (op.has_mma != 0) {
    tc_fence_after();
    mma_steps<P, op.s0, op.s1 - op.s0, T, C2>(c, std::make_integer_sequence<int, op.s1 - op.s0>{});
    commit_any<C2>(&c.bars->acc_full[T]);
  } else {
    // Epilogue-only op: hand the tile back at once.  The epilogue still waits for this arrival, so it can never
    // complete two phases of abuf_ready ahead of this thread (a parity wait cannot tell phase n from phase n + 2).
    mbar_arrive(&c.bars->acc_full[T]);
    if constexpr (C2) mbar_arrive_cluster(mapa_u32(smem_u32(&c.bars->acc_full[T]), 1u));  // ... in the peer CTA too
  }
}

template <int P, int OP, bool C2>
__device__ __forceinline__ void mma_op(MmaCtx& c) {
  mma_op_tile<P, OP, 0, C2>(c);
  mma_op_tile<P, OP, 1, C2>(c);
}

template <int P, bool C2, int... OPS>
__device__ __forceinline__ void mma_program(MmaCtx& c, std::integer_sequence<int, OPS...>) {
  (mma_op<P, OPS, C2>(c), ...);
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <int P, bool C2>
__global__ void __launch_bounds__(kFThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmEnc, const __grid_constant__ CUtensorMap tmActs,
                 const __grid_constant__ WMaps wmaps, const FusedParams p) {
  constexpr int kOps = kSched.prog[P].n_ops;
  const uint32_t rank = C2 ? cluster_ctarank() : 0u;   // pair mode: 0 = leader (issues the MMAs), 1 = peer
  const long long pair_first = (long long)blockIdx.x - rank;  // both CTAs of a pair run the same number of rounds
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* abuf = smem;                       // [2 tiles][4 k-blocks]
  uint8_t* ring = abuf + 2 * kAbufBytes;
  FBarriers* bars = reinterpret_cast<FBarriers*>(ring + (size_t)kSlots * kSlotBytes);

  // shuffled so that the compiler knows the warp index (and everything derived from it) is warp-uniform
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmEnc);
    if (p.save) tma_prefetch_desc(&tmActs);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bars->abuf_ready[t], C2 ? 16 : 8);  // pair mode: the leader's barrier collects both CTAs' warps
      mbar_init(&bars->acc_full[t], 1);
      mbar_init(&bars->enc_ready[t], 2);
      mbar_init(&bars->enc_free[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (C2) tmem_alloc_2cta(&bars->tmem_base, 512);
    else tmem_alloc(&bars->tmem_base, 512);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (C2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them remotely
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ================================ producer ==================================================================
    if (lane == 0) {
      uint32_t eph = 0;  // per-slot phase bits of the empty barriers
      const uint8_t* wblob = p.wblob + (size_t)(blockIdx.x % p.replicas) * p.blob_stride;
      const uint64_t wpol = l2_policy_evict_last();
      const uint64_t epol = l2_policy_evict_first();  // IPE tiles are read once per use
      const int n_loads = c_prog[P].n_loads;
      const bool ipe = p.means != nullptr;
      uint32_t it = 0;  // this CTA's pair counter: scratch buffer it & 1
      for (long long pair0 = pair_first; pair0 < p.num_pairs; pair0 += gridDim.x, ++it) {
        const long long pair = pair0 + rank;
        bool enc_waited = false;
        for (int l = 0; l < n_loads; ++l) {
          const LoadRec ld = c_prog[P].loads[l];
          const int slot = ld.slot;
          mbar_wait(&bars->empty[slot], ((eph >> slot) & 1u) ^ 1u);
          eph ^= 1u << slot;
          if constexpr (C2) {
            // Pair mode: every load is a 2-D TMA whose bytes are counted on the LEADER's `full` barrier; the leader
            // arms it with the bytes of both CTAs (the peer's completions may arrive first: the transaction count
            // may go negative inside a phase).
            uint8_t* dst = ring + (size_t)slot * kSlotBytes;
            const uint32_t bar = mapa_u32(smem_u32(&bars->full[slot]), 0u);
            if (ld.is_enc) {
              long long tile = pair * 2 + ld.t;
              if (tile >= p.num_tiles) tile = p.num_tiles - 1;
              int row0 = (int)(tile * kTileM);
              if (ipe) {
                if (!enc_waited) {
                  mbar_wait(&bars->enc_ready[it & 1u], (it >> 1) & 1u);
                  enc_waited = true;
                }
                row0 = (int)((((long long)blockIdx.x * 2 + (it & 1u)) * 2 + ld.t) * kTileM);
              }
              if (rank == 0) mbar_expect_tx(&bars->full[slot], 2 * 2 * kKbBytes);
              tma_load_2d_2cta(dst, &tmEnc, bar, 0, row0, epol);
              tma_load_2d_2cta(dst + kKbBytes, &tmEnc, bar, 64, row0, epol);
            } else if (p.debug & 1) {  // experiment: no weight traffic (the MMAs read whatever the slot holds)
              if (rank == 0) mbar_arrive(&bars->full[slot]);
            } else if (ld.nkb == 0) {  // bias tile: this CTA's 128 rows of Bt (two 2 KB halves of K) + the ones tile
              const int r0 = (int)(ld.blob_off >> 7);
              if (rank == 0) mbar_expect_tx(&bars->full[slot], 2 * 8192);
              const CUtensorMap* m16 = &wmaps.m[wmap_index(16)];
              tma_load_2d_2cta(dst, m16, bar, 0, r0 + (int)rank * 16, wpol);
              tma_load_2d_2cta(dst + 2048, m16, bar, 0, r0 + 32 + (int)rank * 16, wpol);
              tma_load_2d_2cta(dst + 4096, m16, bar, 0, r0 + 64, wpol);
              tma_load_2d_2cta(dst + 6144, m16, bar, 0, r0 + 80, wpol);
            } else {                   // this CTA's half (N / 2 rows) of every k-block of the weight tile
              const int half = ld.rows >> 1;
              const CUtensorMap* m = &wmaps.m[wmap_index(half)];
              if (rank == 0) mbar_expect_tx(&bars->full[slot], ld.bytes);
              for (int kb = 0; kb < ld.nkb; ++kb)
                tma_load_2d_2cta(dst + (size_t)kb * half * 128, m, bar, 0,
                                 (int)(ld.blob_off >> 7) + kb * ld.rows + (int)rank * half, wpol);
            }
            continue;
          }
          if (ld.is_enc) {
            long long tile = pair * 2 + ld.t;
            if (tile >= p.num_tiles) tile = p.num_tiles - 1;  // phantom tile of an odd tail: recompute the last one
            int row0 = (int)(tile * kTileM);
            if (ipe) {  // the encoder warps have filled scratch buffer it & 1 with this pair's encodings
              if (!enc_waited) {
                mbar_wait(&bars->enc_ready[it & 1u], (it >> 1) & 1u);
                enc_waited = true;
              }
              row0 = (int)((((long long)blockIdx.x * 2 + (it & 1u)) * 2 + ld.t) * kTileM);
            }
            mbar_expect_tx(&bars->full[slot], 2 * kKbBytes);
            tma_load_2d_hint(ring + (size_t)slot * kSlotBytes, &tmEnc, &bars->full[slot], 0, row0, epol);
            tma_load_2d_hint(ring + (size_t)slot * kSlotBytes + kKbBytes, &tmEnc, &bars->full[slot], 64, row0, epol);
          } else if (p.debug & 1) {  // experiment: no weight traffic (the MMAs read whatever the slot holds)
            mbar_arrive(&bars->full[slot]);
          } else {
            mbar_expect_tx(&bars->full[slot], ld.bytes);
            bulk_load_1d(ring + (size_t)slot * kSlotBytes, wblob + ld.blob_off, ld.bytes, &bars->full[slot], wpol);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================================================
    // One thread runs the schedule, fully unrolled at compile time (mma_step<P,S>): descriptor offsets, instruction
    // descriptors, accumulate flags and barrier indices are immediates; only the ring position and the tile are
    // dynamic.
    if (lane == 0) {
      MmaCtx mc;
      mc.bars = bars, mc.full_ph = 0, mc.ab_ph = 0;
      mc.abuf16 = smem_u32(abuf) >> 4, mc.ring16 = smem_u32(ring) >> 4, mc.tmem_base = tmem_base;
      mc.prof = p.prof != nullptr, mc.t_ab = 0, mc.t_full = 0;
      const long long t_start = p.prof != nullptr ? clock64() : 0;
      uint32_t it = 0;
      if (rank == 0)  // (pair mode: the peer's MMA warp only allocates / frees its half of the tensor memory)
      for (long long pair0 = pair_first; pair0 < p.num_pairs; pair0 += gridDim.x, ++it) {
        mma_program<P, C2>(mc, std::make_integer_sequence<int, kOps>{});
        // every TMA load of this pair's encodings has landed (their `full` barriers were waited for above): the
        // encoder may overwrite scratch buffer it & 1 for the pair after next
        if (p.means != nullptr) {
          mbar_arrive(&bars->enc_free[it & 1u]);
          if constexpr (C2) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->enc_free[it & 1u]), 1u));
        }
      }
      if (p.prof != nullptr && rank == 0) {
        p.prof[blockIdx.x * 8 + 0] = (unsigned long long)(clock64() - t_start);
        p.prof[blockIdx.x * 8 + 1] = (unsigned long long)mc.t_ab;
        p.prof[blockIdx.x * 8 + 2] = (unsigned long long)mc.t_full;
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================ encoder warps (in-kernel IPE) ==============================================
    if ((P == P_FWD || P == P_FWDJ) && p.means != nullptr) {
      const int e = (warp - 2) * 32 + lane;  // 0..63: rows e, e + 64, e + 128, e + 192 of the pair's 256
      uint32_t it = 0;
      for (long long pair0 = pair_first; pair0 < p.num_pairs; pair0 += gridDim.x, ++it) {
        const long long pair = pair0 + rank;
        const uint32_t b = it & 1u;
        if (it >= 2) mbar_wait(&bars->enc_free[b], ((it >> 1) - 1u) & 1u);
        __nv_bfloat16* dst = p.enc_scratch + (((size_t)blockIdx.x * 2 + b) * 2 * kTileM) * kEncDim;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int row = e + 64 * j;  // tile row / 128, row % 128
          const long long m = pair * 2 * kTileM + row;
          if (p.debug & 4) continue;  // timing experiment: hand-shakes only (garbage encodings)
          encode_row(p.means, p.covs, m < p.M ? m : p.M - 1, p.ipe_min_deg, dst + (size_t)row * kEncDim);
        }
        // generic-proxy global writes -> async-proxy (TMA) reads by the producer thread
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->enc_ready[b]);
      }
    }
  } else {
    // ================================ epilogue warps ============================================================
    EpiCtx c;
    c.tmActs = &tmActs;
    c.q = warp & 3;            // TMEM lane quadrant (hardware rule: warp id % 4)
    c.hf = (warp - 4) >> 2;    // which of the two warps of the quadrant
    c.lane = lane, c.row = c.q * 32 + lane, c.save = p.save != 0, c.skip = (p.debug & 2) != 0;
    c.direct = p.save == 2, c.tma = p.save == 1, c.tmaw = p.save == 3;
    c.store_policy = l2_policy_evict_first();
    const uint32_t tlane = tmem_base + ((uint32_t)(c.q * 32) << 16);
    // pair mode: every `abuf_ready` arrival of the peer CTA goes to the leader's barrier
    const uint32_t ab_cluster0 = (C2 && rank != 0) ? mapa_u32(smem_u32(&bars->abuf_ready[0]), 0u) : 0u;
    const uint32_t ab_cluster1 = (C2 && rank != 0) ? mapa_u32(smem_u32(&bars->abuf_ready[1]), 0u) : 0u;
    if (lane == 0) {           // both activation buffers / accumulators start out free
      c.ab_cluster = ab_cluster0, c.abuf_ready = &bars->abuf_ready[0];
      epi_arrive(c);
      c.ab_cluster = ab_cluster1, c.abuf_ready = &bars->abuf_ready[1];
      epi_arrive(c);
    }
    uint32_t acc_ph = 0;
    long long t_wait = 0;
    // Sign-bit words are fetched one (op, tile) ahead of their use: the L2 round trip hides behind the previous
    // epilogue.  (In P_FWDJ the words were written by this very thread several ops earlier.)
    auto fetch_bits = [&](long long pair, int e, int t) -> uint4 {
      uint4 out = make_uint4(0u, 0u, 0u, 0u);
      if (p.masks == nullptr || pair >= p.num_pairs) return out;
      const int epi = c_prog[P].ops[e].epi;
      const bool applies = epi == E_MASK || epi == E_BSEED || epi == E_BDZ7 || (epi == E_COLOR && P == P_FWDJ);
      if (!applies) return out;
      const int n_mine = c_prog[P].ops[e].nunits >> 1;
      const long long slot = p.masks_per_tile ? pair * 2 + t : (long long)blockIdx.x * 2 + t;
      const uint32_t* mp = p.masks + (size_t)slot * kMaskWordsPerTile +
                           (c_prog[P].ops[e].mask_plane * 8 + c.hf * n_mine) * kTileM + c.row;
      out.x = mp[0], out.y = mp[kTileM];
      if (n_mine > 2) out.z = mp[2 * kTileM], out.w = mp[3 * kTileM];
      return out;
    };
    uint4 nbits = fetch_bits(blockIdx.x, 0, 0);
    const long long t_epi0 = p.prof != nullptr ? clock64() : 0;
    for (long long pair0 = pair_first; pair0 < p.num_pairs; pair0 += gridDim.x) {
      const long long pair = pair0 + rank;
      for (int e = 0; e < kOps; ++e) {
        const Op op = c_prog[P].ops[e];
        for (int t = 0; t < 2; ++t) {
          const long long tile = pair * 2 + t;
          const bool tile_ok = tile < p.num_tiles;
          c.m = tile * kTileM + c.row;
          c.row_ok = tile_ok && c.m < p.M;
          const long long m_safe = c.row_ok ? c.m : p.M - 1;
          c.tile_row0 = (int)(tile * kTileM);
          c.save = p.save != 0 && tile_ok;
          c.abuf = abuf + t * kAbufBytes;
          c.tacc = tlane + (uint32_t)t * 256u;
          c.acc_full = &bars->acc_full[t], c.abuf_ready = &bars->abuf_ready[t];
          c.ab_cluster = t == 0 ? ab_cluster0 : ab_cluster1;
          c.acc_parity = (acc_ph >> t) & 1u;
          acc_ph ^= 1u << t;
          if (p.prof != nullptr) {  // (the real wait inside the epilogue then returns at once)
            const long long t0 = clock64();
            mbar_wait(c.acc_full, c.acc_parity);
            t_wait += clock64() - t0;
          }
          c.mask = p.masks == nullptr
                       ? nullptr
                       : p.masks + (size_t)(p.masks_per_tile ? tile : (long long)blockIdx.x * 2 + t) * kMaskWordsPerTile;
          c.mbits[0] = nbits.x, c.mbits[1] = nbits.y, c.mbits[2] = nbits.z, c.mbits[3] = nbits.w;
          if (t == 0) nbits = fetch_bits(pair, e, 1);
          else if (e + 1 < kOps) nbits = fetch_bits(pair, e + 1, 0);
          else nbits = fetch_bits(pair + gridDim.x, 0, 0);
          constexpr bool kFwd = P == P_FWD || P == P_FWDJ;
          constexpr bool kChain = P == P_FWDJ || P == P_BWD;
          const int epi = op.epi;
          if (kFwd && epi == E_RELU) {
            rewrite_abuf<M_BIAS_RELU>(c, p, op, nullptr, op.bias_off);
          } else if (kFwd && epi == E_EXTRA) {
            rewrite_abuf<M_BIAS>(c, p, op, nullptr, op.bias_off);
          } else if (kFwd && epi == E_VIEW) {
            // (M < 2^31 is checked at launch: 32-bit divisions, not the ~100-instruction 64-bit subroutine, on the epilogue warps)
            const unsigned ray = (unsigned)m_safe / (unsigned)p.S;
            rewrite_abuf<M_ROWBIAS_RELU>(c, p, op, p.row_bias + (size_t)(p.vb_mod ? ray % (unsigned)p.vb_mod : ray) * kCondW, 0);
          } else if (P != P_FWD && epi == E_MASK) {
            rewrite_abuf<M_MASK>(c, p, op, nullptr, 0);
          } else if (P == P_BWD && epi == E_LIN) {
            rewrite_abuf<M_LIN>(c, p, op, nullptr, 0);
          } else if (P == P_BWD && epi == E_BSEED) {
            rewrite_abuf<M_BSEED>(c, p, op, nullptr, 0);
          } else if (P == P_BWD && epi == E_BDZ7) {
            rewrite_abuf<M_BDZ7>(c, p, op, nullptr, 0);
          } else if (kFwd) switch (epi) {
            case E_DEN:
              epi_wait(c);
              if (c.hf == 0) {
                float v[16];
                tmem_ld16(c.tacc + op.acc_col, v);
                if (c.row_ok) {
#pragma unroll
                  for (int ch = 0; ch < 16; ++ch)
                    if (ch < p.C) p.raw_den[c.m * p.C + ch] = v[ch] + __ldg(p.bblob + kBiasHD + ch);
                }
              }
              epi_done(c);
              break;
            case E_COLOR:
              if (P == P_FWDJ) {
                // raw_rgb first (the accumulator is complete once acc_full fires), then the seed of the Jacobian
                // sweep a_7 = relu'(h_7) * w_sigma replaces the view activations
                epi_wait(c);
                if (c.hf == 0) {
                  float v[16];
                  tmem_ld16(c.tacc + op.acc_col, v);
                  if (c.row_ok) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) p.raw_rgb[c.m * 3 + ch] = v[ch] + __ldg(p.bblob + kBiasC + ch);
                  }
                }
                Op o2 = op;
                o2.has_mma = 2;
                rewrite_abuf<M_SEED>(c, p, o2, nullptr, kWDen);
              } else {
                epi_wait(c);
                if (c.hf == 0) {
                  float v[16];
                  tmem_ld16(c.tacc + op.acc_col, v);
                  if (c.row_ok) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) p.raw_rgb[c.m * 3 + ch] = v[ch] + __ldg(p.bblob + kBiasC + ch);
                  }
                }
                epi_done(c);
              }
              break;
            default:
              break;
          }
          if (kChain && (epi == E_GSKIP || epi == E_G0)) {
            {  // gradient w.r.t. the encoding, 96 fp32 columns: chunks 0,1 -> hf 0, chunk 2 -> hf 1
              epi_wait(c);
              if (p.g_enc != nullptr) {
#pragma unroll 1
                for (int ch = c.hf * 2; ch < (c.hf == 0 ? 2 : 3); ++ch) {
                  uint32_t r[32];
                  tmem_ld32u(c.tacc + op.acc_col + ch * 32, r);
                  tmem_wait_ld();
                  if (c.row_ok) {
                    float4* dst = reinterpret_cast<float4*>(p.g_enc + c.m * kEncDim + ch * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      float4 o = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                             __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
                      if (op.epi == E_G0) {
                        // += the skip-connection part this same thread stored earlier: a vector reduction at L2
                        // instead of a load (no L2 round trip in front of the store)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(o.x), "f"(o.y),
                                     "f"(o.z), "f"(o.w)
                                     : "memory");
                      } else {
                        dst[i] = o;
                      }
                    }
                  }
                }
              }
              epi_done(c);
            }
          }
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    if (p.prof != nullptr && lane == 0 && warp == 4) {
      p.prof[blockIdx.x * 8 + 3] = (unsigned long long)(clock64() - t_epi0);
      p.prof[blockIdx.x * 8 + 4] = (unsigned long long)t_wait;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (C2) cluster_sync_all();  // neither CTA leaves (or frees tensor memory) while the pair's MMAs may still run
  if (warp == 1) {
    tc_fence_after();
    if constexpr (C2) tmem_dealloc_2cta(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

static bool make_map_enc(CUtensorMap* out, const void* base, unsigned long long rows, unsigned long long ld) {
  return make_map(out, base, rows, kEncDim, ld, 64, kTileM);
}

// PNB_FUSED_SAVE: unset / "t" = 16 KB TMA stores of whole 128-row k-blocks (4 warps rendezvous), "w" = per-warp 4 KB
// stores (no rendezvous), "d" = st.global from the epilogue registers
static int save_mode() {
  static const int m = [] {
    const char* e = getenv("PNB_FUSED_SAVE");
    return e == nullptr ? 1 : e[0] == 'd' ? 2 : e[0] == 'w' ? 3 : 1;
  }();
  return m;
}

static bool make_map_acts(CUtensorMap* out, const void* base, unsigned long long planes, unsigned long long rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error_msg("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)kWidth, rows, planes};
  cuuint64_t strides[2] = {(cuuint64_t)kWidth * 2, rows * (cuuint64_t)kWidth * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)(save_mode() == 3 ? 32 : kTileM), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error_msg("cuTensorMapEncodeTiled failed for the activation planes");
    return false;
  }
  return true;
}

// copies are 4 KB-aligned plus an odd number of 256-byte lines apart, so that equal tile offsets land on different slices
static long long blob_stride() { return ((long long)kSched.blob_bytes + 4095) / 4096 * 4096 + 256 * 37; }

// The weight blob as a 2-D tensor of 128-byte rows for the pair-mode producer (TMA with cta_group::2); one map per box
// height.  Encoding is host-only work (no driver call reaches the stream), cached per blob address.
static bool make_wmaps(WMaps* out, const void* blob) {
  static thread_local const void* cached_blob = nullptr;
  static thread_local WMaps cached;
  if (cached_blob == blob) {
    *out = cached;
    return true;
  }
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error_msg("cuTensorMapEncodeTiled not available from the driver");
    return false;
  }
  static const unsigned box_rows[5] = {128, 64, 48, 16, 8};
  const cuuint64_t rows = (cuuint64_t)kSched.blob_bytes / 128;
  for (int i = 0; i < 5; ++i) {
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, box_rows[i]};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&out->m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(blob), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error_msg("cuTensorMapEncodeTiled failed for the weight blob");
      return false;
    }
  }
  cached = *out, cached_blob = blob;
  return true;
}

// PNB_FUSED_2CTA=1 selects the CTA-pair kernels, =0 the single-CTA ones.
static bool pair_mode() {
  static const int v = [] {
    const char* e = getenv("PNB_FUSED_2CTA");
    return e == nullptr ? 0 : atoi(e);
  }();
  return v != 0;
}

struct LaunchEnv {  // experiment switches, read once
  int debug, save_direct, replicas, prof;
  LaunchEnv() {
    const char* e = getenv("PNB_FUSED_DEBUG");
    debug = e ? atoi(e) : 0;
    e = getenv("PNB_FUSED_SAVE");
    save_direct = (e != nullptr && e[0] == 'd');
    e = getenv("PNB_FUSED_REPLICAS");
    replicas = e ? atoi(e) : kReplicas;
    if (replicas < 1 || replicas > kReplicas) replicas = kReplicas;
    prof = getenv("PNB_FUSED_PROF") != nullptr;
  }
};

template <int P, bool C2>
static int launch_impl(const CUtensorMap& tmEnc, const CUtensorMap& tmActs, FusedParams& p, cudaStream_t st,
                       const char* what) {
  static const LaunchEnv env;
  const size_t fixed = 1024 + 2 * kAbufBytes + sizeof(FBarriers);
  p.nstages = kSlots;
  const size_t smem_bytes = fixed + (size_t)kSlots * kSlotBytes;
  p.debug = env.debug;  // timing experiments only (wrong results)
  if (p.save != 0) {
    // default: 16 KB TMA stores from the activation buffer.  "direct" (st.global from the epilogue registers) was
    // measured 25-35 % slower: 16-byte pieces at a 512-byte lane stride saturate the LSU store path.
    p.save = save_mode();
  }
  p.replicas = C2 ? 1 : env.replicas, p.blob_stride = blob_stride();
  long long gx = p.num_pairs < kNumSMs ? p.num_pairs : kNumSMs;
  if (C2) gx = (gx + 1) / 2 * 2;  // whole pairs of CTAs (a pair without a second tile pair recomputes the last one)
  WMaps wm{};
  if (C2 && !make_wmaps(&wm, p.wblob)) return PNB_ERR_ARG;
  static unsigned long long* prof_buf = nullptr;
  if (env.prof) {
    if (prof_buf == nullptr) cudaMalloc(&prof_buf, sizeof(unsigned long long) * 8 * kNumSMs);
    cudaMemsetAsync(prof_buf, 0, sizeof(unsigned long long) * 8 * kNumSMs, st);
    p.prof = prof_buf;
  }
  cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<P, C2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_bytes);
  if (e != cudaSuccess) {
    set_error("mlp_fused(smem attr)", e);
    return (int)e;
  }
  if (p.bblob != nullptr) {
    e = cudaMemcpyToSymbolAsync(c_bblob, p.bblob, sizeof(float) * kBiasFloats, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
      set_error("mlp_fused(bias staging)", e);
      return (int)e;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)gx), cfg.blockDim = dim3(kFThreads), cfg.dynamicSmemBytes = smem_bytes, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C2 ? 2 : 1, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, mlp_fused_kernel<P, C2>, tmEnc, tmActs, wm, p);
  if (e != cudaSuccess) {
    set_error(what, e);
    return (int)e;
  }
  if (env.prof) {  // timing experiments only: per-CTA cycle counters, averaged over the CTAs that issue MMAs
    unsigned long long h[8 * kNumSMs];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    double a[5] = {0, 0, 0, 0, 0};
    const int stride = C2 ? 2 : 1;
    for (long long b = 0; b < gx; b += stride)
      for (int i = 0; i < 5; ++i) a[i] += (double)h[b * 8 + i] / (double)(gx / stride);
    fprintf(stderr, "[%s%s] cycles/CTA: mma total %.0f (wait abuf %.0f, wait ring %.0f) | epilogue warp total %.0f (wait acc %.0f)\n",
            what, C2 ? ", CTA pairs" : "", a[0], a[1], a[2], a[3], a[4]);
  }
  return finish(what);
}

template <int P>
static int launch(const CUtensorMap& tmEnc, const CUtensorMap& tmActs, FusedParams& p, cudaStream_t st,
                  const char* what) {
  if (pair_mode()) return launch_impl<P, true>(tmEnc, tmActs, p, st, what);
  return launch_impl<P, false>(tmEnc, tmActs, p, st, what);
}

}  // namespace fused
}  // namespace pnb

using namespace pnb;
using namespace pnb::fused;

extern "C" long long pnb_mlp_fused_wblob_bytes(void) { return blob_stride() * kReplicas; }
extern "C" long long pnb_mlp_fused_bblob_floats(void) { return kBiasFloats; }
extern "C" int pnb_mlp_fused_act_planes(void) { return kActPlanes; }
extern "C" int pnb_mlp_fused_bwd_planes(void) { return kBwdPlanes; }
extern "C" int pnb_mlp_fused_adj_planes(void) { return kAdjPlanes; }

extern "C" long long pnb_mlp_fused_mask_words(long long M, int per_tile) {
  const long long tiles = (M + kTileM - 1) / kTileM;
  const long long slots = per_tile ? (tiles + 3) / 4 * 4 : 2ll * kNumSMs;  // (a CTA pair walks 4 tiles per round)
  return slots * kMaskWordsPerTile;
}

extern "C" int pnb_mlp_fused_pack(const void* const* params_host, int C, void* wblob, float* bblob, void* stream) {
  PNB_REQUIRE(params_host != nullptr && wblob != nullptr && bblob != nullptr, "mlp_fused_pack: null argument");
  PNB_REQUIRE(C >= 1 && C <= 16, "mlp_fused_pack: need 1 <= C <= 16 density-head channels");
  const Sched& s = kSched;
  static const int in_features[kNumParams] = {96, 256, 256, 256, 256, 352, 256, 256, 256, 256, 283, 128};
  PackArgs a{};
  for (int i = 0; i < kNumParams; ++i) {
    a.w[i] = reinterpret_cast<const float*>(params_host[2 * i]);
    a.b[i] = reinterpret_cast<const float*>(params_host[2 * i + 1]);
    a.ld[i] = in_features[i];
    PNB_REQUIRE(a.w[i] != nullptr && a.b[i] != nullptr, "mlp_fused_pack: null parameter pointer");
  }
  a.C = C;
  a.n_tiles = s.n_pack;
  cudaStream_t st = as_stream(stream);
  pack_tiles_kernel<<<dim3(s.n_pack, kReplicas), 256, 0, st>>>(a, reinterpret_cast<uint8_t*>(wblob), blob_stride());
  int rc = finish("mlp_fused_pack(tiles)");
  if (rc) return rc;
  pack_bias_kernel<<<8, 256, 0, st>>>(a, bblob);
  return finish("mlp_fused_pack(bias)");
}

extern "C" int pnb_mlp_fused_fwd(long long M, int S, int C, const void* enc, int ld_enc, const void* wblob,
                                 const float* bblob, const float* row_bias, int vb_mod, float* raw_den, float* raw_rgb,
                                 void* acts, float* g_enc, void* masks, int masks_per_tile, void* stream) {
  PNB_REQUIRE(M >= 0 && S >= 1 && C >= 1 && C <= 16 && vb_mod >= 0, "mlp_fused_fwd: bad sizes");
  PNB_REQUIRE(M == 0 || (enc && wblob && bblob && row_bias && raw_den && raw_rgb), "mlp_fused_fwd: null argument");
  PNB_REQUIRE(ld_enc % 8 == 0 && ld_enc >= kEncDim && ((uintptr_t)enc % 16 == 0) && ((uintptr_t)wblob % 16 == 0) &&
                  ((uintptr_t)bblob % 16 == 0) && ((uintptr_t)row_bias % 16 == 0),
              "mlp_fused_fwd: enc / blobs / row_bias must be 16-byte aligned, ld_enc % 8 == 0");
  PNB_REQUIRE(g_enc == nullptr || (uintptr_t)g_enc % 16 == 0, "mlp_fused_fwd: g_enc must be 16-byte aligned");
  PNB_REQUIRE(acts == nullptr || (uintptr_t)acts % 128 == 0, "mlp_fused_fwd: acts must be 128-byte aligned");
  PNB_REQUIRE(g_enc == nullptr || masks != nullptr, "mlp_fused_fwd: the Jacobian sweep needs the sign-bit buffer");
  PNB_REQUIRE(M < (1ll << 31) - 2 * kTileM, "mlp_fused_fwd: M too large for 32-bit TMA coordinates");
  if (M == 0) return 0;
  FusedParams p{};
  p.M = M, p.num_tiles = (M + kTileM - 1) / kTileM, p.num_pairs = (p.num_tiles + 1) / 2;
  p.S = S, p.C = C, p.save = acts != nullptr, p.vb_mod = vb_mod;
  p.wblob = reinterpret_cast<const uint8_t*>(wblob), p.bblob = bblob, p.row_bias = row_bias;
  p.raw_den = raw_den, p.raw_rgb = raw_rgb, p.g_enc = g_enc;
  p.masks = reinterpret_cast<uint32_t*>(masks), p.masks_per_tile = masks_per_tile;
  p.planes = reinterpret_cast<__nv_bfloat16*>(acts);
  CUtensorMap tmEnc, tmActs;
  if (!make_map_enc(&tmEnc, enc, (unsigned long long)M, (unsigned long long)ld_enc)) return PNB_ERR_ARG;
  if (p.save) {
    if (!make_map_acts(&tmActs, acts, kActPlanes, (unsigned long long)M)) return PNB_ERR_ARG;
  } else {
    tmActs = tmEnc;
  }
  cudaStream_t st = as_stream(stream);
  if (g_enc != nullptr) return launch<P_FWDJ>(tmEnc, tmActs, p, st, "mlp_fused_fwd(jac)");
  return launch<P_FWD>(tmEnc, tmActs, p, st, "mlp_fused_fwd");
}

extern "C" long long pnb_mlp_fused_scratch_bytes(void) {
  return (long long)kNumSMs * 2 * 2 * kTileM * kEncDim * (long long)sizeof(__nv_bfloat16);
}

// Inference forward with the integrated positional encoding computed inside the kernel (no [M,96] encoding array):
// means / covs [M,3] fp32 in, raw outputs (and d sigma / d enc when g_enc is given) out.
extern "C" int pnb_mlp_fused_fwd_ipe(long long M, int S, int C, const float* means, const float* covs, int min_deg,
                                     const void* wblob, const float* bblob, const float* row_bias, int vb_mod,
                                     float* raw_den, float* raw_rgb, float* g_enc, void* masks, void* scratch,
                                     void* stream) {
  PNB_REQUIRE(M >= 0 && S >= 1 && C >= 1 && C <= 16 && vb_mod >= 0, "mlp_fused_fwd_ipe: bad sizes");
  PNB_REQUIRE(M == 0 || (means && covs && wblob && bblob && row_bias && raw_den && raw_rgb && scratch),
              "mlp_fused_fwd_ipe: null argument");
  PNB_REQUIRE(min_deg >= 0 && min_deg + 16 <= 31, "mlp_fused_fwd_ipe: IPE degrees must lie in [0, 31)");
  PNB_REQUIRE(((uintptr_t)wblob % 16 == 0) && ((uintptr_t)bblob % 16 == 0) && ((uintptr_t)row_bias % 16 == 0) &&
                  ((uintptr_t)scratch % 128 == 0),
              "mlp_fused_fwd_ipe: blobs / row_bias must be 16-byte, scratch 128-byte aligned");
  PNB_REQUIRE(g_enc == nullptr || ((uintptr_t)g_enc % 16 == 0 && masks != nullptr),
              "mlp_fused_fwd_ipe: g_enc must be 16-byte aligned and needs the sign-bit buffer");
  PNB_REQUIRE(M < (1ll << 31) - 2 * kTileM, "mlp_fused_fwd_ipe: M too large for 32-bit TMA coordinates");
  if (M == 0) return 0;
  FusedParams p{};
  p.M = M, p.num_tiles = (M + kTileM - 1) / kTileM, p.num_pairs = (p.num_tiles + 1) / 2;
  p.S = S, p.C = C, p.save = 0, p.vb_mod = vb_mod;
  p.wblob = reinterpret_cast<const uint8_t*>(wblob), p.bblob = bblob, p.row_bias = row_bias;
  p.raw_den = raw_den, p.raw_rgb = raw_rgb, p.g_enc = g_enc;
  p.masks = reinterpret_cast<uint32_t*>(masks), p.masks_per_tile = 0;
  p.means = means, p.covs = covs, p.ipe_min_deg = min_deg;
  p.enc_scratch = reinterpret_cast<__nv_bfloat16*>(scratch);
  CUtensorMap tmEnc;
  if (!make_map_enc(&tmEnc, scratch, (unsigned long long)kNumSMs * 4 * kTileM, kEncDim)) return PNB_ERR_ARG;
  cudaStream_t st = as_stream(stream);
  if (g_enc != nullptr) return launch<P_FWDJ>(tmEnc, tmEnc, p, st, "mlp_fused_fwd_ipe(jac)");
  return launch<P_FWD>(tmEnc, tmEnc, p, st, "mlp_fused_fwd_ipe");
}

extern "C" int pnb_mlp_fused_bwd(long long M, int C, const void* wblob, const float* bblob, const float* d_rgb,
                                 const float* d_den, const void* masks, void* dz_planes, float* d_enc, void* stream) {
  PNB_REQUIRE(M >= 0 && C >= 1 && C <= 16, "mlp_fused_bwd: bad sizes");
  PNB_REQUIRE(M == 0 || (wblob && bblob && d_rgb && d_den && masks && dz_planes), "mlp_fused_bwd: null argument");
  PNB_REQUIRE(((uintptr_t)wblob % 16 == 0) && ((uintptr_t)bblob % 16 == 0) && ((uintptr_t)dz_planes % 128 == 0) &&
                  (d_enc == nullptr || (uintptr_t)d_enc % 16 == 0),
              "mlp_fused_bwd: misaligned argument");
  PNB_REQUIRE(M < (1ll << 31) - 2 * kTileM, "mlp_fused_bwd: M too large for 32-bit TMA coordinates");
  if (M == 0) return 0;
  FusedParams p{};
  p.M = M, p.num_tiles = (M + kTileM - 1) / kTileM, p.num_pairs = (p.num_tiles + 1) / 2;
  p.S = 1, p.C = C, p.save = 1;
  p.wblob = reinterpret_cast<const uint8_t*>(wblob), p.bblob = bblob;
  p.g_enc = d_enc, p.d_rgb = d_rgb, p.d_den = d_den;
  p.masks = reinterpret_cast<uint32_t*>(const_cast<void*>(masks)), p.masks_per_tile = 1;
  p.planes = reinterpret_cast<__nv_bfloat16*>(dz_planes);
  CUtensorMap tmActs;
  if (!make_map_acts(&tmActs, dz_planes, kBwdPlanes, (unsigned long long)M)) return PNB_ERR_ARG;
  return launch<P_BWD>(tmActs, tmActs, p, as_stream(stream), "mlp_fused_bwd");
}

extern "C" int pnb_mlp_fused_jadj(long long M, const void* u, int ld_u, const void* wblob, const void* masks,
                                  void* q_planes, void* stream) {
  PNB_REQUIRE(M >= 0, "mlp_fused_jadj: bad sizes");
  PNB_REQUIRE(M == 0 || (u && wblob && masks && q_planes), "mlp_fused_jadj: null argument");
  PNB_REQUIRE(ld_u % 8 == 0 && ld_u >= kEncDim && ((uintptr_t)u % 16 == 0) && ((uintptr_t)wblob % 16 == 0) &&
                  ((uintptr_t)q_planes % 128 == 0),
              "mlp_fused_jadj: misaligned argument");
  PNB_REQUIRE(M < (1ll << 31) - 2 * kTileM, "mlp_fused_jadj: M too large for 32-bit TMA coordinates");
  if (M == 0) return 0;
  FusedParams p{};
  p.M = M, p.num_tiles = (M + kTileM - 1) / kTileM, p.num_pairs = (p.num_tiles + 1) / 2;
  p.S = 1, p.C = 1, p.save = 1;
  p.wblob = reinterpret_cast<const uint8_t*>(wblob);
  p.masks = reinterpret_cast<uint32_t*>(const_cast<void*>(masks)), p.masks_per_tile = 1;
  p.planes = reinterpret_cast<__nv_bfloat16*>(q_planes);
  CUtensorMap tmEnc, tmActs;
  if (!make_map_enc(&tmEnc, u, (unsigned long long)M, (unsigned long long)ld_u)) return PNB_ERR_ARG;
  if (!make_map_acts(&tmActs, q_planes, kAdjPlanes, (unsigned long long)M)) return PNB_ERR_ARG;
  return launch<P_JADJ>(tmEnc, tmActs, p, as_stream(stream), "mlp_fused_jadj");
}

// Introspection for the tests: the static ring plan of a program.  out[0] = loads per tile pair, out[1] = MMA step
// executions per tile pair, then per load {slot, is_enc, t, blob_off}, then per executed step (in MMA order)
// {tile, step, blob_off, is_aenc, w_slot, w_load, w_rel, e_slot, e_load, e_rel, first, op}.
extern "C" int pnb_mlp_fused_plan(int prog, long long* out, int cap) {
  PNB_REQUIRE(prog >= 0 && prog < kNumProgs && out != nullptr && cap >= 2, "mlp_fused_plan: bad arguments");
  const Prog& g = kSched.prog[prog];
  out[0] = g.n_loads, out[1] = 2 * g.n_steps;
  int n = 2;
  for (int l = 0; l < g.n_loads && n + 4 <= cap; ++l) {
    out[n++] = g.loads[l].slot, out[n++] = g.loads[l].is_enc, out[n++] = g.loads[l].t, out[n++] = g.loads[l].blob_off;
  }
  for (int o = 0; o < g.n_ops; ++o)
    for (int t = 0; t < 2; ++t)
      for (int k = 0; k < g.ops[o].s1 - g.ops[o].s0 && n + 12 <= cap; ++k) {
        const int sidx = (t == 0 || PNB_RING_MODE < 2) ? g.ops[o].s0 + k : g.ops[o].s1 - 1 - k;
        const Use& u = g.uses[t][sidx];
        out[n++] = t, out[n++] = sidx, out[n++] = g.steps[sidx].blob_off, out[n++] = (g.steps[sidx].flags & F_AENC) ? 1 : 0;
        out[n++] = u.w_slot, out[n++] = u.w_load, out[n++] = u.w_rel, out[n++] = u.e_slot, out[n++] = u.e_load;
        out[n++] = u.e_rel, out[n++] = u.first, out[n++] = o;
      }
  return 0;
}
