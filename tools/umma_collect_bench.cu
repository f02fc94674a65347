// Micro-benchmark + correctness probe of the tcgen05.mma A-operand collector (`.collector::a::fill/use/lastuse`,
// SASS: UTCHMMA ...A_KEEP / A_REUSE) for the split MMAs the row-stacked conv issues when an accumulator window
// wraps the TMEM ring: one logical N=192 MMA becomes N=128 + N=64 (or 3 x N=64) on the SAME A tile.  Question:
// does re-using A from the collector save the second A fetch from shared memory (4 KB = 32 cycles at 128 B/clk)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_collect_bench tools/umma_collect_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"
using namespace pnp;

enum Variant { kN192 = 0, kSplit3 = 1, kSplit3Collect = 2 };

template <int kCta, int kColl>   // kColl: 0 none, 1 fill, 2 use, 3 lastuse
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
#define PNP_MMA(str)                                                                                            \
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"   \
               "setp.ne.b32 p, %6, 0;\n\t" str " [%0], da, db, %5, p;\n\t}" ::"r"(d),                            \
               "r"(a_lo), "r"(kDescHiSw128), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"(acc)                  \
               : "memory")
  if constexpr (kCta == 1) {
    if constexpr (kColl == 0) PNP_MMA("tcgen05.mma.cta_group::1.kind::f16");
    if constexpr (kColl == 1) PNP_MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::fill");
    if constexpr (kColl == 2) PNP_MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::use");
    if constexpr (kColl == 3) PNP_MMA("tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse");
  } else {
    if constexpr (kColl == 0) PNP_MMA("tcgen05.mma.cta_group::2.kind::f16");
    if constexpr (kColl == 1) PNP_MMA("tcgen05.mma.cta_group::2.kind::f16.collector::a::fill");
    if constexpr (kColl == 2) PNP_MMA("tcgen05.mma.cta_group::2.kind::f16.collector::a::use");
    if constexpr (kColl == 3) PNP_MMA("tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse");
  }
#undef PNP_MMA
}

__device__ __forceinline__ uint32_t idesc_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// one logical (M x 192 x 16) product on A tile `a_lo`, B region at `b` (layout depends on the variant)
template <int kCta>
__device__ __forceinline__ void logical_mma(int variant, uint32_t d, uint32_t a_lo, uint32_t b192, uint32_t b64,
                                            uint32_t acc) {
  constexpr uint32_t M = kCta == 2 ? 256u : 128u;
  // descriptor units (16 B): a 64-row sub-block of a single-CTA B is 64*128/16 = 512 units, with cta_group::2 each
  // CTA holds half of the rows of every MMA's B
  constexpr uint32_t sb = kCta == 2 ? 256u : 512u;
  switch (variant) {
    case kN192:
      mma<kCta, 0>(d, a_lo, b192, idesc_mn(M, 192), acc);
      break;
    case kSplit3:       // three N=64 MMAs out of the per-sub-block layout
      mma<kCta, 0>(d, a_lo, b64, idesc_mn(M, 64), acc);
      mma<kCta, 0>(d + 64, a_lo, b64 + sb, idesc_mn(M, 64), acc);
      mma<kCta, 0>(d + 128, a_lo, b64 + 2 * sb, idesc_mn(M, 64), acc);
      break;
    case kSplit3Collect:
      mma<kCta, 1>(d, a_lo, b64, idesc_mn(M, 64), acc);
      mma<kCta, 2>(d + 64, a_lo, b64 + sb, idesc_mn(M, 64), acc);
      mma<kCta, 3>(d + 128, a_lo, b64 + 2 * sb, idesc_mn(M, 64), acc);
      break;
  }
}

__device__ __forceinline__ float hashf(uint32_t x) {     // small integers: exact in bf16 and in the fp32 sums
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return (float)((int)(x % 7u) - 3);
}

// write element (row, k) of a K-major 128-byte-row SWIZZLE_128B tile that starts at 1024-aligned `base`
__device__ __forceinline__ void put(uint8_t* base, int row, int k, float v) {
  const int chunk = (k >> 3) ^ (row & 7);
  reinterpret_cast<__nv_bfloat16*>(base + row * 128 + chunk * 16)[k & 7] = __float2bfloat16(v);
}

constexpr int kAOff = 0;                // A: 128 rows x 128 B
constexpr int kB192Off = 32 * 1024;     // N=192 layout (24 KB single CTA; 12 KB per CTA of a pair)
constexpr int kB64Off = 64 * 1024;      // per-sub-block layout: three 64-row blocks (pair: three 32-row halves)
constexpr int kSmem = 100 * 1024;

template <int kCta>
__global__ void __launch_bounds__(128, 1) bench(int variant, int per_group, int reps, long long* out_cycles, int* bad) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = kCta == 2 ? cluster_ctarank() : 0u;
  // logical operands: A_r (128 x 64) per CTA, B (192 x 64) shared
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int row = i >> 6, k = i & 63;
    put(sgen + kAOff, row, k, hashf(0x1000u + rank * 77777u + i));
  }
  for (int i = threadIdx.x; i < 192 * 64; i += blockDim.x) {
    const int n = i >> 6, k = i & 63;
    const float v = hashf(0x9000000u + i);
    if (kCta == 1) {
      put(sgen + kB192Off, n, k, v);
      put(sgen + kB64Off, n, k, v);
    } else {
      if (n / 96 == (int)rank) put(sgen + kB192Off, n % 96, k, v);
      const int s = n / 64, half = (n % 64) / 32;
      if (half == (int)rank) put(sgen + kB64Off, s * 32 + (n % 32), k, v);
    }
  }
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
      mbar_fence_init();
    }
    __syncwarp();
    if (kCta == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(smem_u32(&tmem_slot), 512);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t a_lo0 = umma_desc_lo(sbase + kAOff), b192 = umma_desc_lo(sbase + kB192Off), b64 = umma_desc_lo(sbase + kB64Off);
  uint32_t phase = 0;
  auto commit = [&]() {
    if (kCta == 2)
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    else
      umma_commit(smem_u32(&bar));
  };
  // ---- correctness: K = 64 product with the variant into columns [256, 448), with plain N=192 into [0, 192)
  if (warp == 1 && rank == 0) {
    if (elect_one()) {
      for (int k = 0; k < 4; ++k) logical_mma<kCta>(kN192, tmem, a_lo0 + 2 * k, b192 + 2 * k, b64 + 2 * k, k > 0);
      for (int k = 0; k < 4; ++k) logical_mma<kCta>(variant, tmem + 256, a_lo0 + 2 * k, b192 + 2 * k, b64 + 2 * k, k > 0);
      commit();
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), phase, 98);
  phase ^= 1;
  tc_fence_after();
  {
    int nbad = 0;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 192; c += 16) {
      float r0[16], r1[16];
      tmem_ld16(lane_base + c, r0);
      tmem_ld16(lane_base + 256 + c, r1);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) nbad += (r0[j] != r1[j]) ? 1 : 0;
      if (c == 0 && threadIdx.x == 5 && blockIdx.x < 2 && variant == kN192) {
        // scalar reference of D[row 5][col 0..1] for this CTA's A
        for (int col = 0; col < 2; ++col) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += hashf(0x1000u + rank * 77777u + 5 * 64 + k) * hashf(0x9000000u + col * 64 + k);
          if (ref != r0[col]) nbad += 1000;
        }
      }
    }
    if (nbad) atomicAdd(bad, nbad);
  }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  tc_fence_after();
  // ---- timing
  if (warp == 1 && rank == 0) {
    long long t0 = 0;
    for (int r = -2; r < reps; ++r) {
      if (r == 0) t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < per_group; ++i) {
          // twelve distinct A slices (3 dx shifts x 4 K slices) and B blocks per "step", like the conv kernel
          const uint32_t a_lo = a_lo0 + (uint32_t)((i % 3) * 8 + ((i / 3) & 3) * 2);
          const uint32_t bo = (uint32_t)(((i / 3) & 3) * 2);
          logical_mma<kCta>(variant, tmem + (uint32_t)((i & 1) * 256), a_lo, b192 + bo, b64 + bo, 1);
        }
        commit();
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, 99);
      phase ^= 1;
    }
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x / kCta] = clock64() - t0;
  } else if (kCta == 2 && warp == 1) {
    // the peer's barrier receives the multicast commits as well: keep its phase in step (nothing to do, it is not waited on)
  }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (kCta == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else
      tmem_dealloc(tmem, 512);
  }
}

template <int kCta>
void run(int sms, const char* name, int variant, long long* d, int* dbad) {
  const int reps = 100, per_group = 144;
  const int grid = kCta == 2 ? (sms / 2) * 2 : sms;
  cudaMemset(d, 0, sizeof(long long) * 256);
  cudaMemset(dbad, 0, sizeof(int));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = kSmem + 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, bench<kCta>, variant, per_group, reps, d, dbad);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("cta_group::%d %-28s: %s\n", kCta, name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[256];
  int bad = 0;
  const int units = grid / kCta;
  cudaMemcpy(h, d, sizeof(long long) * units, cudaMemcpyDeviceToHost);
  cudaMemcpy(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < units; ++i) mean += (double)h[i];
  mean /= units;
  printf("cta_group::%d %-28s: %6.1f cycles per logical N=192 MMA, mismatching accumulator values %d\n", kCta, name,
         mean / ((double)reps * per_group), bad);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  int* dbad;
  cudaMalloc(&d, sizeof(long long) * 256);
  cudaMalloc(&dbad, sizeof(int));
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem + 1024);
  cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem + 1024);
  const char* names[3] = {"N=192", "3 x N=64", "3 x N=64 fill/use/lastuse"};
  for (int v : {0, 1, 2}) run<1>(sms, names[v], v, d, dbad);
  for (int v : {0, 1, 2}) run<2>(sms, names[v], v, d, dbad);
  return 0;
}
