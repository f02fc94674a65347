// Probe of tcgen05.mma with the A operand in TMEM (".ts" form): is a bf16 A tile written by tcgen05.st as "lane = row,
// 32-bit column c = elements (2c, 2c+1), K-slice j at column offset 8*j" what the MMA reads?  Compares D = A * B^T computed
// with A from shared memory (SS) and from TMEM (TS) for M=128, N=64, K=64, plus cycles per TS MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_ts_probe tools/umma_ts_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"
using namespace pnp;

__device__ __forceinline__ float hashf(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return (float)((int)(x % 7u) - 3);
}
__device__ __forceinline__ void put(uint8_t* base, int row, int k, float v) {
  const int chunk = (k >> 3) ^ (row & 7);
  reinterpret_cast<__nv_bfloat16*>(base + row * 128 + chunk * 16)[k & 7] = __float2bfloat16(v);
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo),
               "r"(kDescHiSw128), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int reps, long long* cycles, int* bad) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  uint8_t* sa = sgen;                 // A: 128 x 64 bf16
  uint8_t* sb = sgen + 16384;         // B: 64 x 64 bf16
  for (int i = tid; i < 128 * 64; i += 128) put(sa, i >> 6, i & 63, hashf(0x1000u + i));
  for (int i = tid; i < 64 * 64; i += 128) put(sb, i >> 6, i & 63, hashf(0x9000000u + i));
  if (warp == 0) {
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // every thread = one row: write its 64 bf16 of A into TMEM columns [256, 288) of its lane
  {
    uint32_t r[32];
    for (int c = 0; c < 32; ++c)
      r[c] = pack_bf16x2(hashf(0x1000u + tid * 64 + 2 * c), hashf(0x1000u + tid * 64 + 2 * c + 1));
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = umma_idesc_bf16(128, 64);
  uint32_t phase = 0;
  if (warp == 1) {
    if (elect_one()) {
      for (int j = 0; j < 4; ++j)      // SS reference into columns [0, 64)
        umma_bf16_lo(tmem, umma_desc_lo(smem_u32(sa)) + 2 * j, kDescHiSw128, umma_desc_lo(smem_u32(sb)) + 2 * j, kDescHiSw128, idesc, j > 0);
      for (int j = 0; j < 4; ++j)      // TS into columns [64, 128)
        umma_ts(tmem + 64, tmem + 256 + 8 * j, umma_desc_lo(smem_u32(sb)) + 2 * j, idesc, j > 0);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), phase, 1);
  phase ^= 1;
  tc_fence_after();
  {
    int nbad = 0;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 64; c += 16) {
      float r0[16], r1[16];
      tmem_ld16(lane_base + c, r0);
      tmem_ld16(lane_base + 64 + c, r1);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) nbad += (r0[j] != r1[j]) ? 1 : 0;
      if (c == 0 && tid == 5) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += hashf(0x1000u + 5 * 64 + k) * hashf(0x9000000u + k);
        if (ref != r0[0]) nbad += 1000;
        printf("row 5 col 0: scalar %g, SS %g, TS %g\n", ref, r0[0], r1[0]);
      }
    }
    if (nbad) atomicAdd(bad, nbad);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    long long t0 = 0;
    for (int r = -2; r < reps; ++r) {
      if (r == 0) t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < 144; ++i) umma_ts(tmem + 64 * (i & 3), tmem + 256 + 8 * (i & 3), umma_desc_lo(smem_u32(sb)) + 2 * (i & 3), idesc, 1);
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, 2);
      phase ^= 1;
    }
    if ((tid & 31) == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; int* dbad;
  cudaMalloc(&d, sizeof(long long) * 256); cudaMalloc(&dbad, sizeof(int));
  cudaMemset(dbad, 0, sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<148, 128, 64 * 1024>>>(100, d, dbad);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[148]; int bad = 0;
  cudaMemcpy(h, d, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
  cudaMemcpy(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < 148; ++i) mean += (double)h[i]; mean /= 148;
  printf("TS (A in TMEM) M=128 N=64 K=16: %.1f cycles per MMA; mismatching accumulator values vs SS over 148 CTAs: %d\n",
         mean / (100.0 * 144), bad);
  return 0;
}
