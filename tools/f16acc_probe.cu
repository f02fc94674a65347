// Probe: does tcgen05.mma kind::f16 accept bf16 operands with an F16 accumulator (c_format = 0), and how is D packed?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/f16acc_probe tools/f16acc_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"
using namespace pnp;

__global__ void __launch_bounds__(128, 1) probe(int cfmt, uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint16_t* a = reinterpret_cast<uint16_t*>(smem_raw + (sbase - raw));            // 128 rows x 128 B
  uint16_t* b = a + 128 * 64;                                                     // 64 rows x 128 B
  for (int i = threadIdx.x; i < 128 * 64; i += 128) a[i] = 0x3F00;                // bf16 0.5
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i / 64;
    const float v = 0.25f * (n % 8);
    b[i] = (uint16_t)(__float_as_uint(v) >> 16);
  }
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x / 32 == 1) {
    if (elect_one()) {
      const uint32_t idesc = ((uint32_t)cfmt << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
      for (int k = 0; k < 4; ++k)
        umma_bf16_lo(tmem, umma_desc_lo(sbase) + 2 * k, kDescHiSw128, umma_desc_lo(sbase + 128 * 128) + 2 * k, kDescHiSw128,
                     idesc, k > 0);
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0, 1);
  tc_fence_after();
  float v[16];
  const uint32_t lane_base = tmem + ((uint32_t)((threadIdx.x / 32) * 32) << 16);
  for (int c = 0; c < 64; c += 16) {
    tmem_ld16(lane_base + c, v);
    tmem_ld_wait();
    if (threadIdx.x == 5)
      for (int j = 0; j < 16; ++j) out[c + j] = __float_as_uint(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int cfmt : {1, 0}) {
    cudaMemset(d, 0xff, 64 * 4);
    probe<<<1, 128, 64 * 1024>>>(cfmt, d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("c_format=%d (%s accumulator): %s\n", cfmt, cfmt ? "F32" : "F16", cudaGetErrorString(e));
    if (e != cudaSuccess) return 0;
    uint32_t h[64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("  expected D[m][n] = 16 * 0.5 * 0.25 * (n %% 8) = n %% 8\n  columns 0..15 raw:");
    for (int j = 0; j < 16; ++j) printf(" %08x", h[j]);
    printf("\n  as f32:");
    for (int j = 0; j < 16; ++j) { float f; memcpy(&f, &h[j], 4); printf(" %g", f); }
    printf("\n  as 2 x f16:");
    for (int j = 0; j < 16; ++j) { __half2 hh; memcpy(&hh, &h[j], 4); printf(" (%g,%g)", __half2float(hh.x), __half2float(hh.y)); }
    printf("\n  columns 32..39 raw:");
    for (int j = 32; j < 40; ++j) printf(" %08x", h[j]);
    printf("\n");
  }
  return 0;
}
