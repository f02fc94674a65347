// Parameters of the CTA-pair fused residual-block kernel (pnp_block.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace pnp {

constexpr int kBlockOutPx = 126;                       // output pixels per strip row (128 minus the x halo of conv1)
constexpr int kBlockW0Bytes = (9 + 3) * 64 * 128;      // conv2 row-stacked (72 KB) + three stacked 1x1 (24 KB)
constexpr int kBlockW1Bytes = 9 * 64 * 128;            // conv1 row-stacked (72 KB)

struct BlockParams {
  CUtensorMap tm_src;   // x   (64, W, H, N) bf16, box (64,130,1,1), SWIZZLE_128B
  CUtensorMap tm_out;   // out (64, W, H, N) bf16, box (64,126,1,1)
  const void* w0;       // kBlockW0Bytes: pnp_pack_conv3x3_rowstack(conv2 mix) followed by 192 pnp_pack_rows rows
  const void* w1;       // kBlockW1Bytes: pnp_pack_conv3x3_rowstack(conv1)
  const float* bias0;   // [64] bias of the first stage (SE gain * mixed conv2 bias), or null
  const float* bias1;   // [64] conv1 bias, or null
  const float* par;     // partition map (N,3,H,W) fp32 view
  long long par_sn, par_sc, par_sy;
  const void* x;        // same tensor as tm_src, for the identity add (read through the global path)
  int H, W, N, strips;
  int tiles_total, tiles_per_pair;
  int s_a;              // source-row ring slots of the first-stage CTA
  int n_t;              // intermediate-row ring slots of the second-stage CTA
  long long* trace;     // optional clock64 stamps of cluster 0 (diagnostics)
  int debug_skip;       // what-if bits (results WRONG): 1 skip source-row loads, 2 skip stores, 4 skip TMEM loads
};

size_t block_smem_bytes(const BlockParams& p);
cudaError_t launch_block(const BlockParams& p, int pairs, cudaStream_t stream);

}  // namespace pnp
