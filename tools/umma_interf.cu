// Micro-benchmark: what slows tcgen05.mma down inside the conv kernels?  (diagnostic, not product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_interf tools/umma_interf.cu
// One CTA per SM; warp 1 issues row-stacked steps (12 x N=192, M=128, K=16, deferred commit) like the conv
// kernels do, while other warps generate the side traffic of a real epilogue / producer, paced at one
// "row" per `pace` cycles.  Interference mask:
//   1  operands are pseudo-random bf16 instead of a constant (datapath toggling / power)
//   2  8 warps: 128 TMEM values per thread per row (tcgen05.ld x16 x 8 + wait)
//   4  8 warps: 64 B per thread per row of st.shared.v4 + fence.proxy.async (staging an output row)
//   8  1 thread: 16.6 KB global -> shared bulk copy per row (source-row producer)
//  16  8 warps: 8 broadcast ld.shared.v4 per row
//  32  8 warps: one mbarrier try_wait poll per ~100 cycles
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"

using namespace pnp;

__global__ void __launch_bounds__(640, 1)
interf_kernel(int steps, int mask, int pace, int nread, const uint8_t* gsrc, long long* out_cycles, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar, done_bar[8], ld_bar, never_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t stop_flag;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: A rows at 0 (4 slots x 17 KB), B at 96 KB (72 KB); scratch staging at 170 KB; bulk dst at 188 KB
  for (int i = threadIdx.x; i < 170 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (mask & 1) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      // two bf16 in (-1,1): sign + exponent 0x3e..0x3f + random mantissa
      const uint32_t lo = (h & 0x807fu) | 0x3f00u, hi = ((h >> 16) & 0x807fu) | 0x3e80u;
      v = lo | (hi << 16);
    }
    reinterpret_cast<uint32_t*>(sgen)[i] = v;
  }
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
      mbar_init(smem_u32(&ld_bar), 1);
      mbar_init(smem_u32(&never_bar), 1);
      for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&done_bar[i]), 1);
      stop_flag = 0;
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t stop_addr = smem_u32(&stop_flag);
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t id64 = umma_idesc_bf16(128, 64), id128 = umma_idesc_bf16(128, 128),
                     id192 = umma_idesc_bf16(128, 192);
      const uint32_t a_lo0 = umma_desc_lo(sbase), b_lo0 = umma_desc_lo(sbase + 96 * 1024);
      const long long t0 = clock64();
      bool pend = false;
      uint32_t pend_bar = 0;
      if (mask & 64) {
        while (clock64() - t0 < 400000) {
        }
      }
      for (int s = 0; s < ((mask & 64) ? 0 : steps); ++s) {
        const uint32_t slot = (uint32_t)(s & 3) * 64;
        const uint32_t a_row = a_lo0 + (uint32_t)(s & 3) * 1088;
        umma_bf16_lo(tmem + slot, a_row, kDescHiSw128, b_lo0, kDescHiSw128, id128, 1);
        umma_bf16_lo(tmem + slot + 128, a_row, kDescHiSw128, b_lo0 + 1024, kDescHiSw128, id64, 0);
        if (pend) umma_commit(pend_bar);
#pragma unroll
        for (int i = 1; i < 12; ++i)
          umma_bf16_lo(tmem + slot, a_row + (i >> 2) * 8 + 2 * (i & 3), kDescHiSw128,
                       b_lo0 + (i >> 2) * 1536 + 2 * (i & 3), kDescHiSw128, id192, 1);
        pend = true;
        pend_bar = smem_u32(&done_bar[s & 7]);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, 96);
      out_cycles[blockIdx.x] = clock64() - t0;
      st_release_shared(stop_addr, 1);
    }
  } else if (warp == 2) {
    if ((mask & 8) && elect_one()) {
      uint32_t phase = 0;
      long long t_next = clock64();
      while (ld_acquire_shared(stop_addr) == 0) {
        while (clock64() < t_next) {
        }
        t_next += pace;
        mbar_arrive_expect_tx(smem_u32(&ld_bar), 16640);
        bulk_load_1d(sbase + 188 * 1024, gsrc + (size_t)blockIdx.x * 16640, 16640, smem_u32(&ld_bar));
        mbar_wait(smem_u32(&ld_bar), phase, 95);
        phase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 4 + nread) {
    const int q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float acc = 0.f;
    long long t_next = clock64();
    long long ld_cycles = 0;
    int ld_batches = 0;
    uint8_t* rowp = sgen + 170 * 1024 + (q * 32 + lane) * 128;
    const uint32_t swz = (uint32_t)(lane & 7);
    const int half = (warp - 4) >> 2;
    if (mask & 128) {                       // pure ALU pressure on every sub-partition (no memory ops)
      float x0 = acc + 1.f, x1 = 2.f, x2 = 3.f, x3 = 4.f;
      int it = 0;
      do {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          x0 = fmaf(x0, 1.0001f, 0.5f);
          x1 = fmaf(x1, 1.0002f, 0.25f);
          x2 = fmaf(x2, 0.9999f, 0.125f);
          x3 = fmaf(x3, 0.9998f, 0.0625f);
        }
        ++it;
      } while ((it & 7) != 0 || ld_acquire_shared(stop_addr) == 0);
      acc += x0 + x1 + x2 + x3;
    }
    while (ld_acquire_shared(stop_addr) == 0) {
      while (clock64() < t_next) {
        if (mask & 32) {
          if (lane == 0) acc += mbar_try_wait(smem_u32(&never_bar), 0) ? 1.f : 0.f;
          __nanosleep(50);
        }
      }
      t_next += pace;
      if (mask & 2) {
        const long long l0 = clock64();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float v[64];
          tmem_ld16(lane_base + r * 64, v);
          tmem_ld16(lane_base + r * 64 + 16, v + 16);
          tmem_ld16(lane_base + r * 64 + 32, v + 32);
          tmem_ld16(lane_base + r * 64 + 48, v + 48);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 64; ++j) acc += v[j];
        }
        ld_cycles += clock64() - l0;
        ++ld_batches;
      }
      if (mask & 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float bx, by, bz, bw;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(bx), "=f"(by), "=f"(bz), "=f"(bw)
                       : "r"(sbase + 96 * 1024 + j * 16));
          acc += bx + bw;
        }
      }
      if (mask & 4) {
        const uint32_t w = __float_as_uint(acc);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(rowp + (((4 * half + c) ^ swz) << 4)) = make_uint4(w, w + 1, w + 2, w + 3);
        fence_proxy_async_smem();
      }
    }
    if (acc == 123.456f) sink[threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 128 && ld_batches > 0) out_cycles[255] = ld_cycles / ld_batches;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  float* sink;
  uint8_t* gsrc;
  cudaMalloc(&d, sizeof(long long) * 256);
  cudaMemset(d, 0, sizeof(long long) * 256);
  cudaMalloc(&sink, 4096);
  cudaMalloc(&gsrc, (size_t)sms * 16640);
  cudaMemset(gsrc, 0x3c, (size_t)sms * 16640);
  cudaFuncSetAttribute(interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int steps = 400;
  const int masks[] = {0, 128};
  for (int nread : {4, 8, 16}) {
    const int pace = 1200;
    for (int mask : masks) {
      double best = 1e30;
      long long ld = 0;
      for (int rep = 0; rep < 3; ++rep) {
        interf_kernel<<<sms, 640, 210 * 1024>>>(steps, mask, pace, nread, gsrc, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mask %d: %s\n", mask, cudaGetErrorString(e));
          return 1;
        }
        long long h[256];
        cudaMemcpy(h, d, sizeof(long long) * 256, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < sms; ++i) mean += (double)h[i];
        mean /= sms;
        if (mean < best) best = mean;
        ld = h[255];
      }
      printf("readers %2d mask %2d: %7.1f cycles/step  (%5.1f per MMA; ideal 12 x 98.6 = 1183); 128-value TMEM read batch: %lld cycles\n",
             nread, mask, best / steps, best / steps / 12.0, ld);
    }
  }
  return 0;
}
