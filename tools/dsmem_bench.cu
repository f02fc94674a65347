// Micro-benchmark: pushing 16 KB tiles into the partner CTA's shared memory (cluster of 2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/dsmem_bench tools/dsmem_bench.cu
// Diagnostic only (design input for the fused block kernel): cycles per 16 KB tile for
//   mode 0: st.shared::cluster.v4, thread = 128-byte pixel row (8 x 16 B at 128 B stride, swizzled)
//   mode 1: st.shared::cluster.v4, lanes contiguous (512 B per warp instruction)
//   mode 2: cp.async.bulk.shared::cluster.shared::cta of the whole tile, complete_tx on the remote mbarrier
//   mode 3: mode 0 with a local st.shared (no DSMEM) for reference
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"

using namespace pnp;

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

constexpr int kTile = 16384;
constexpr int kSlots = 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
dsmem_kernel(int mode, int reps, int writers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const uint32_t rank = cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < (kSlots + 1) * kTile / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw))[i] = i;
  __syncthreads();
  cluster_sync();
  const uint32_t remote = mapa(sbase, rank ^ 1);
  const uint32_t remote_bar = mapa(smem_u32(&bar), rank ^ 1);
  long long t0 = clock64();
  if (rank == 0) {
    if (mode == 2) {
      if (threadIdx.x == 32) {
        for (int r = 0; r < reps; ++r) {
          const uint32_t dst = remote + (r % kSlots) * kTile;
          asm volatile(
              "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
              "r"(sbase + kSlots * kTile), "r"(kTile), "r"(remote_bar)
              : "memory");
        }
      }
    } else if (warp >= 1 && warp <= writers) {
      const int w = warp - 1;
      const uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
      for (int r = 0; r < reps; ++r) {
        const uint32_t base = (mode == 3 ? sbase : remote) + (r % kSlots) * kTile;
        // rows of this warp: with `writers` warps, warp w covers rows w*32.. in steps
        for (int row0 = w * 32; row0 < 128; row0 += writers * 32) {
          if (mode == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) st_cluster_v4(base + row0 * 128 + k * 512 + lane * 16, v);
          } else {
            const int row = row0 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) st_cluster_v4(base + row * 128 + ((k ^ (row & 7)) << 4), v);
          }
        }
      }
    }
  } else {
    if (mode == 2 && threadIdx.x == 0) {
      for (int r = 0; r < reps; ++r) {
        mbar_arrive_expect_tx(smem_u32(&bar), kTile);
        mbar_wait(smem_u32(&bar), r & 1, 1);
      }
    }
  }
  cluster_sync();   // release/acquire at cluster scope: all remote stores have landed
  long long t1 = clock64();
  if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / 2] = t1 - t0;
}

int main(int argc, char** argv) {
  const int reps = 2000;
  long long* d;
  cudaMalloc(&d, 74 * sizeof(long long));
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int grid : {2, 148}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int writers : {4, 2}) {
        if (mode == 2 && writers != 4) continue;
        dsmem_kernel<<<grid, 160, 100 * 1024>>>(mode, reps, writers, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d: %s\n", mode, cudaGetErrorString(e));
          return 1;
        }
        long long h[74];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1LL << 60;
        for (int i = 0; i < grid / 2; ++i) {
          mx = h[i] > mx ? h[i] : mx;
          mn = h[i] < mn ? h[i] : mn;
        }
        printf("grid %3d mode %d writers %d: cycles per 16 KB tile min %.1f max %.1f  (%.1f B/cycle)\n", grid, mode,
               writers, (double)mn / reps, (double)mx / reps, 16384.0 * reps / mx);
      }
    }
  }
  return 0;
}
