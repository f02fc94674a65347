// Micro-benchmark of tcgen05.mma.cta_group::2 (M = 256 across a CTA pair) next to cta_group::1: cycles per MMA for the
// N values the conv kernels use, issued back to back by the leader CTA into one accumulator (diagnostic for round 2:
// with cta_group::2 each CTA fetches only N/2 rows of the B operand from its own shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma2_bench tools/umma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"
using namespace pnp;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
bench2(int n, int per_group, int reps, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw))[i] = 0x3c003c00u;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((256u >> 4) << 24);
    const uint32_t a_lo0 = umma_desc_lo(sbase), b_lo0 = umma_desc_lo(sbase + 96 * 1024);
    long long t0 = 0;
    uint32_t phase = 0;
    for (int r = -2; r < reps; ++r) {
      if (r == 0) t0 = clock64();
      if (elect_one()) {
#pragma unroll 4
        for (int i = 0; i < per_group; ++i) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((i * 2) & 4095), b_lo = b_lo0 + (uint32_t)((i * 2) & 2047);
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
              "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
              "setp.ne.b32 p, %6, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
              ::"r"(tmem), "r"(a_lo), "r"(kDescHiSw128), "r"(b_lo), "r"(kDescHiSw128), "r"(idesc), "r"((uint32_t)(i > 0))
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, 99);
      phase ^= 1;
    }
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 256);
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 100, per_group = 144;
  for (int n : {64, 96, 128, 192, 256}) {
    cudaMemset(d, 0, sizeof(long long) * 256);
    bench2<<<(sms / 2) * 2, 128, 200 * 1024>>>(n, per_group, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 0; }
    long long h[128];
    cudaMemcpy(h, d, sizeof(long long) * (sms / 2), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms / 2; ++i) mean += (double)h[i];
    mean /= (sms / 2);
    const double per = mean / ((double)reps * per_group);
    printf("cta_group::2  M=256 N=%3d: %6.1f cycles per MMA  (%5.0f MAC/cycle/SM, %4.1f %% of 4096; per-CTA smem fetch %4.0f B = %4.1f cycles at 128 B/clk)\n",
           n, per, 256.0 * n * 16 / per / 2, 100.0 * 256.0 * n * 16 / per / 2 / 4096.0, 4096.0 + 16.0 * n, (4096.0 + 16.0 * n) / 128.0);
  }
  return 0;
}
