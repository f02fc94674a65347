// Micro-benchmark of tcgen05.mma issue patterns on sm_100a (diagnostic, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_bench tools/umma_bench.cu
// For each pattern: one CTA per SM, `reps` groups of `per_group` MMAs (M=128, K=16, bf16) issued by one
// elected thread, cycles per MMA from clock64 around issue + commit + wait.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../pnpvcve_b200/csrc/pnp_ptx.cuh"

using namespace pnp;

// pattern: n = MMA N; n_acc = accumulators rotated round-robin; a_step/b_step = descriptor advance per MMA
// (in 16-byte units) so successive MMAs read different smem (like conv taps / K steps)
__global__ void __launch_bounds__(128, 1)
bench_kernel(int n, int n_acc, int per_group, int reps, int a_span, int b_span, int a_shift, int commit_every,
             int acc_off, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar2[0]), 1);
    mbar_init(smem_u32(&bar2[1]), 1);
  }
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw))[i] = 0x3c003c00u;  // bf16 ~0.0078
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
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
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a_lo0 = umma_desc_lo(sbase) + (uint32_t)a_shift;   // a_shift in 16-byte units
    const uint32_t b_lo0 = umma_desc_lo(sbase + 96 * 1024);
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    for (int r = -2; r < reps; ++r) {
      if (r == 0) t0 = clock64();
      if (elect_one()) {
        const int inner = commit_every > 0 ? commit_every : per_group;
        for (int i0 = 0; i0 < per_group; i0 += inner) {
#pragma unroll 4
          for (int i = i0; i < i0 + inner; ++i) {   // spans / n_acc are powers of two: mask, no division
            const uint32_t acc = (uint32_t)(i & (n_acc - 1)) * n + (uint32_t)acc_off;
            umma_bf16_lo(tmem + acc, a_lo0 + (uint32_t)((i * 2) & (a_span - 1)), kDescHiSw128,
                         b_lo0 + (uint32_t)((i * 2) & (b_span - 1)), kDescHiSw128, idesc, i >= n_acc);
          }
          if (commit_every > 0) {                   // un-waited commits, like the conv kernel's per tile
            umma_commit(smem_u32(&bar2[0]));
            umma_commit(smem_u32(&bar2[1]));
          }
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, 99);
      phase ^= 1;
    }
    t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Replica of one CTA of the row-stacked conv kernel's MMA thread: per step [N=128 acc + N=64 overwrite],
// optional deferred commit, 11 x N=192, optional `gap_waits` already-complete mbarrier waits between steps.
__global__ void __launch_bounds__(128, 1)
step_kernel(int steps, int use_commit, int gap_waits, int spin_warps, long long* out_cycles, int fill = 0) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar, done_bar[8], ready_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
      mbar_init(smem_u32(&ready_bar), 1);
      for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&done_bar[i]), 1);
      mbar_fence_init();
      mbar_arrive(smem_u32(&ready_bar));      // phase 0 complete: waits on parity 0 return at once
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  // fill: 0 = whatever the SM holds, 1 = constant, 2 = pseudo-random bf16 in (-2,2) (data-dependent power?)
  if (fill)
    for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) {
      uint32_t h = (uint32_t)i * 2654435761u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      const uint32_t v = (h & 0x807F807Fu) | 0x3F003F00u | ((h >> 3) & 0x00800080u);
      reinterpret_cast<uint32_t*>(smem_raw + (sbase - raw))[i] = fill == 2 ? v : 0x3c003c00u;
    }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t id64 = umma_idesc_bf16(128, 64), id128 = umma_idesc_bf16(128, 128),
                     id192 = umma_idesc_bf16(128, 192);
      const uint32_t a_lo0 = umma_desc_lo(sbase), b_lo0 = umma_desc_lo(sbase + 96 * 1024);
      const long long t0 = clock64();
      bool pend = false;
      uint32_t pend_bar = 0;
      for (int sIdx = 0; sIdx < steps; ++sIdx) {
        for (int g = 0; g < gap_waits; ++g) mbar_wait(smem_u32(&ready_bar), 0, 97);
        tc_fence_after();
        const uint32_t slot = (uint32_t)(sIdx & 3) * 64;      // keep N=192 inside 512 columns
        const uint32_t a_row = a_lo0 + (uint32_t)(sIdx & 3) * 1088;
        umma_bf16_lo(tmem + slot, a_row, kDescHiSw128, b_lo0, kDescHiSw128, id128, 1);
        umma_bf16_lo(tmem + slot + 128, a_row, kDescHiSw128, b_lo0 + 1024, kDescHiSw128, id64, 0);
        if (pend && use_commit) umma_commit(pend_bar);
#pragma unroll
        for (int i = 1; i < 12; ++i)
          umma_bf16_lo(tmem + slot, a_row + (i >> 2) * 8 + 2 * (i & 3), kDescHiSw128,
                       b_lo0 + (i >> 2) * 1536 + 2 * (i & 3), kDescHiSw128, id192, 1);
        pend = true;
        pend_bar = smem_u32(&done_bar[sIdx & 7]);
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0, 96);
      out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (warp >= 2 && warp < 2 + spin_warps) {
    // emulate epilogue warps polling a barrier that never completes until the end
    while (!mbar_try_wait(smem_u32(&bar), 0)) {
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Issue-queue probe: timestamps after each of 24 back-to-back MMA issues (no waits in between).
__global__ void __launch_bounds__(128, 1) queue_probe(int n, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (threadIdx.x == 0) {
      mbar_init(smem_u32(&bar), 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_slot), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t a_lo0 = umma_desc_lo(sbase), b_lo0 = umma_desc_lo(sbase + 96 * 1024);
    long long ts[25];
    uint32_t phase = 0;
    for (int r = 0; r < 3; ++r) {
      if (elect_one()) {
        ts[0] = clock64();
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          umma_bf16_lo(tmem, a_lo0 + 2 * (i & 3), kDescHiSw128, b_lo0 + 2 * (i & 3), kDescHiSw128, idesc, i > 0);
          ts[i + 1] = clock64();
        }
        umma_commit(smem_u32(&bar));
        if (r == 2 && blockIdx.x == 0)
          for (int i = 0; i < 25; ++i) out[i] = ts[i] - ts[0];
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, 98);
      phase ^= 1;
      if (r == 2 && blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[25] = clock64();
    }
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
  cudaMalloc(&d, sizeof(long long) * sms);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Pat { int n, n_acc, per_group, a_span, b_span; const char* what; int a_shift = 0; int commit_every = 0; int acc_off = 0; };
  const Pat pats[] = {
      {64, 1, 36, 4096, 4096, "N=64  1 accumulator, 36 per group (drain each tile)"},
      {64, 1, 144, 4096, 4096, "N=64  1 accumulator, 144 per group"},
      {64, 1, 144, 4096, 4096, "N=64  144 per group, 2 un-waited commits every 36", 0, 36},
      {64, 1, 144, 4096, 4096, "N=64  144 per group, 2 un-waited commits every 12", 0, 12},
      {64, 2, 72, 4096, 4096, "N=64  2 accumulators interleaved"},
      {64, 4, 144, 4096, 4096, "N=64  4 accumulators interleaved"},
      {64, 1, 144, 2, 2, "N=64  1 accumulator, same operands every MMA"},
      {192, 1, 144, 4096, 2048, "N=192 1 accumulator"},
      {192, 1, 144, 4096, 2048, "N=192, 1 un-waited commit every 12", 0, 12},
      {192, 1, 144, 4096, 2048, "N=192 D at column 64", 0, 0, 64},
      {192, 1, 144, 4096, 2048, "N=192 D at column 128", 0, 0, 128},
      {192, 1, 144, 4096, 2048, "N=192 D at column 192", 0, 0, 192},
      {192, 1, 144, 4096, 2048, "N=192 D at column 320", 0, 0, 320},
      {128, 1, 144, 4096, 4096, "N=128 D at column 64", 0, 0, 64},
      {128, 1, 144, 4096, 4096, "N=128 D at column 192", 0, 0, 192},
      {64, 1, 144, 4096, 4096, "N=64  D at column 448", 0, 0, 448},
      {96, 1, 144, 4096, 4096, "N=96  1 accumulator"},
      {160, 1, 144, 4096, 2048, "N=160 1 accumulator"},
      {128, 1, 144, 4096, 4096, "N=128 1 accumulator"},
      {128, 1, 144, 4096, 4096, "N=128, 2 un-waited commits every 12", 0, 12},
      {128, 2, 72, 4096, 4096, "N=128 2 accumulators interleaved"},
      {256, 1, 144, 4096, 2048, "N=256 1 accumulator"},
      {256, 2, 72, 4096, 2048, "N=256 2 accumulators interleaved"},
      {64, 1, 144, 4096, 4096, "N=64  A start shifted by 1 pixel  (+128 B)", 8},
      {64, 1, 144, 4096, 4096, "N=64  A start shifted by 2 pixels (+256 B)", 16},
      {64, 1, 144, 4096, 4096, "N=64  A start shifted by 8 pixels (+1024 B)", 64},
      {128, 1, 144, 4096, 4096, "N=128 A start shifted by 1 pixel  (+128 B)", 8},
      {256, 1, 144, 4096, 2048, "N=256 A start shifted by 1 pixel  (+128 B)", 8},
      {16, 1, 144, 4096, 4096, "N=16  1 accumulator (conv_last)"},
      {32, 1, 144, 4096, 4096, "N=32  1 accumulator"},
  };
  cudaFuncSetAttribute(queue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n : {64, 192, 256}) {
    cudaMemset(d, 0, sizeof(long long) * 32);
    queue_probe<<<sms, 128, 200 * 1024>>>(n, d);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, d, sizeof(long long) * 32, cudaMemcpyDeviceToHost);
    printf("issue timestamps N=%d:", n);
    for (int i = 1; i <= 24; ++i) printf(" %lld", h[i]);
    printf("\n");
  }
  cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct SP { int commit, gaps, spin; const char* what; int fill; };
  const SP sps[] = {{0, 0, 0, "steps: no commit, no gap"},       {1, 0, 0, "steps: deferred commit, no gap"},
                    {1, 0, 0, "steps: deferred commit, no gap, CONSTANT operands", 1},
                    {1, 0, 0, "steps: deferred commit, no gap, RANDOM operands", 2},
                    {1, 0, 2, "steps: deferred commit, no gap, random, 2 polling warps", 2},
                    {0, 4, 0, "steps: no commit, 4 waits gap"},  {1, 4, 0, "steps: deferred commit, 4 waits gap"},
                    {1, 4, 2, "steps: commit + gap + 2 polling warps"}, {1, 2, 0, "steps: deferred commit, 2 waits gap"}};
  for (const SP& sp : sps) {
    step_kernel<<<sms, 128, 200 * 1024>>>(200, sp.commit, sp.gaps, sp.spin, d, sp.fill);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", sp.what, cudaGetErrorString(e)); return 1; }
    long long h[256];
    cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    printf("%-58s %7.0f cycles/step (ideal 12 x 98.6 + 18 = 1201)\n", sp.what, mean / sms / 200.0);
  }
  const int reps = 200;
  for (const Pat& p : pats) {
    bench_kernel<<<sms, 128, 200 * 1024>>>(p.n, p.n_acc, p.per_group, reps, p.a_span, p.b_span, p.a_shift, p.commit_every, p.acc_off, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", p.what, cudaGetErrorString(e));
      return 1;
    }
    long long h[256];
    cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    const double per_mma = mean / ((double)reps * p.per_group);
    const double macs = 128.0 * p.n * 16;
    printf("%-58s %7.1f cyc/MMA  %6.0f MAC/cyc/SM  (%4.1f%% of 4096)\n", p.what, per_mma, macs / per_mma,
           100.0 * macs / per_mma / 4096.0);
  }
  return 0;
}
