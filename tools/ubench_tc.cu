// Micro-benchmarks that size the fused-MLP design (tools only, not part of the library):
//   mode 0: tcgen05.mma alone (M=128, N=256, K=16 bf16, operands in SMEM)      -> cycles per MMA
//   mode 1: tcgen05.ld alone from `nw` epilogue warps (32x32b.x32)              -> TMEM read bytes / cycle / SM
//   mode 2: both at once                                                       -> do they share a port?
//   mode 3: tcgen05.ld + convert to bf16 + 16-byte st.shared (epilogue body)   -> cycles per 128x256 accumulator
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I panonerf_b200/csrc tools/ubench_tc.cu -o ubench_tc
#include <cstdio>
#include <vector>

#include "tc_common.cuh"

namespace pnb {
void set_error(const char*, cudaError_t) {}
void set_error_msg(const char*) {}
void count_launch(int) {}
}  // namespace pnb
using namespace pnb;
using namespace pnb::tc;

__device__ __forceinline__ void ld32u(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t desc16(uint32_t addr16) {
  constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)kHi << 32) | (uint64_t)((addr16 & 0x3FFFu) | (1u << 16));
}

struct Bars {
  uint64_t done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(320, 1) ubench(int mode, int nw, int n_mma, int n_ld, int mma_n, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  Bars* bars = reinterpret_cast<Bars*>(smem + 128 * 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bars->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  long long t0 = clock64(), t1 = t0;
  if (warp == 1) {
    if (mode == 0 || mode == 2 || mode == 4) {
      const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 65536) >> 4;
      const uint32_t idesc = instr_desc_bf16(128, mma_n, 0, 0);
      if (lane == 0) {
        for (int i = 0; i < n_mma; ++i)
          umma_f16(tmem_base + (i & 1) * 256, desc16(a16 + 2 * (i & 3)), desc16(b16 + 2 * (i & 3)), idesc, i > 1);
        umma_commit(&bars->done);
      }
      __syncwarp();
      mbar_wait(&bars->done, 0);
      t1 = clock64();
    }
  } else if (warp >= 2 && warp < 2 + nw) {
    if (mode >= 1) {
      const int q = warp & 3, hf = (warp - 2) >> 2;
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
      uint32_t acc = 0;
      uint32_t r[2][32];
      const int row = q * 32 + lane;
      if (mode >= 3) {
        // epilogue body: 4 units of 32 columns per warp and accumulator, software-pipelined loads
        for (int it = 0; it < n_ld; ++it) {
          ld32u(tl + hf * 32, r[0]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int u = hf + 2 * i;
            wait_ld();
            if (i + 1 < 4) ld32u(tl + (u + 2) * 32, r[(i + 1) & 1]);
            uint8_t* dst = smem + (u >> 1) * 16384 + row * 128;
            const int jb = (u & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float a = fmaxf(__uint_as_float(r[i & 1][8 * j + 2 * k]) + 1.0f, 0.f);
                float b = fmaxf(__uint_as_float(r[i & 1][8 * j + 2 * k + 1]) + 1.0f, 0.f);
                h[k] = __floats2bfloat162_rn(a, b);
              }
              *reinterpret_cast<uint4*>(dst + (((jb + j) ^ (row & 7)) << 4)) = *reinterpret_cast<uint4*>(h);
            }
            fence_async_smem();
          }
        }
      } else {
        for (int it = 0; it < n_ld; ++it) {
          ld32u(tl + ((it * 2 + hf) & 15) * 32, r[0]);
          wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[0][j];
        }
      }
      t1 = clock64();
      if (acc == 0x12345678u) out[1000] = acc;
    }
  }
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 16 * 8 + 16384);
  const size_t smem_bytes = 1024 + 128 * 1024 + 256;
  cudaFuncSetAttribute(ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  struct Cfg { int mode, nw, n_mma, n_ld; const char* name; int mma_n = 256; };
  std::vector<Cfg> cfgs = {
      {0, 0, 4096, 0, "mma only"},          {1, 4, 0, 4096, "ld only, 4 warps"},
      {1, 8, 0, 4096, "ld only, 8 warps"},  {2, 8, 4096, 4096, "mma + ld (8 warps)"},
      {2, 4, 4096, 4096, "mma + ld (4 warps)"}, {3, 8, 0, 512, "epilogue body, 8 warps"},
      {3, 4, 0, 512, "epilogue body, 4 warps (half the columns)"},
      {0, 0, 8192, 0, "mma only N=128", 128},
      {0, 0, 16384, 0, "mma only N=64", 64},
      {4, 8, 2048, 512, "mma N=256 + epilogue body (8 warps)", 256},
      {4, 8, 4096, 512, "mma N=128 + epilogue body (8 warps)", 128},
  };
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_out, 0, 148 * 16 * 8);
      ubench<<<148, 320, smem_bytes>>>(c.mode, c.nw, c.n_mma, c.n_ld, c.mma_n, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", c.name, cudaGetErrorString(e));
        return 1;
      }
    }
    std::vector<long long> h(148 * 16);
    cudaMemcpy(h.data(), d_out, 148 * 16 * 8, cudaMemcpyDeviceToHost);
    long long mma = h[1], ldmax = 0;
    for (int w = 2; w < 10; ++w) ldmax = h[w] > ldmax ? h[w] : ldmax;
    printf("%-45s mma warp %8lld cyc", c.name, mma);
    if (c.n_mma) printf(" (%.1f cyc/MMA)", (double)mma / c.n_mma);
    printf("   ld warps %8lld cyc", ldmax);
    if (c.mode == 1 || c.mode == 2)
      printf(" (%.1f cyc per 4 KB ld per warp; %.1f B/cyc/SM)", (double)ldmax / c.n_ld, 4096.0 * c.n_ld * c.nw / ldmax);
    if (c.mode >= 3) printf(" (%.1f cyc per 128x256 accumulator rewrite)", (double)ldmax / c.n_ld);
    printf("\n");
  }
  return 0;
}
