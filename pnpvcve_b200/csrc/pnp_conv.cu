// 3x3 convolution (64 -> 64 | 16 channels) as an implicit GEMM on the 5th-gen tensor cores.
//
// Replaces, per launch, one F.conv2d of the reference plus everything elementwise around it:
//   * ResidualBlockNoBNDynamic_drt.forward  (mmedit/models/common/sr_backbone_utils.py:304-333):
//       launch A:  t = relu( gamma * (conv3x3(x, Wmix) + bmix) + sum_k par_k * conv1x1_k(x) )
//       launch B:  x = x + conv3x3(t, W1) + b1
//   * input_conv + LeakyReLU (basicvsr_net.py:484,515), one source tensor per launch, partial sums
//     chained through the `id` operand; the 3-channel LR frame rides along as an im2col'd 1x1 (aux)
//   * conv_hr + LeakyReLU, conv_last + `out += lq`  (iconvsr_ipb_par.py:144-146)
//
// Layout: activations are NHWC bf16 with 64 channels = one 128-byte row per pixel, which is exactly
// one SWIZZLE_128B row for TMA and for the UMMA shared-memory descriptor.
//
// Work decomposition: an output tile is 128 consecutive pixels of one image row (UMMA M = 128).
// A persistent CTA walks down a 128-pixel-wide strip; source rows (130 pixels incl. x halo) are
// streamed once by TMA into a ring, and the nine taps of a tile are nine *views* of three ring rows:
// the A descriptor start is moved by (dx+1) pixels = (dx+1)*128 bytes and dy selects the ring slot.
// TMA zero fill outside the image implements the conv padding.  Weights stay resident in smem.
// Accumulators live in TMEM (two 256-column buffers) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue.
#include "pnp_conv.cuh"
#include "pnp_ptx.cuh"

namespace pnp {

namespace {

struct SmemLayout {
  uint32_t w, a, aux, io, misc, total;
};

__host__ __device__ inline SmemLayout make_layout(int n_wchunks, int s_a, int has_aux, int n_io) {
  SmemLayout l;
  l.w = 0;
  l.a = l.w + n_wchunks * kWChunkBytes;
  l.aux = l.a + s_a * kASlotBytes;
  l.io = l.aux + (has_aux ? 2 * kTileBytes : 0);
  l.misc = l.io + n_io * kTileBytes;
  l.total = l.misc + 1024;
  return l;
}

// misc region: 64 floats scale, 64 floats bias, then barriers
struct Misc {
  float scale[64];
  float bias[64];
  uint64_t w_full;
  uint64_t a_full[kMaxASlots];
  uint64_t a_empty[kMaxASlots];
  uint64_t aux_full[2];
  uint64_t aux_empty[2];
  uint64_t id_full[kMaxIoSlots];
  uint64_t io_empty[kMaxIoSlots];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  uint32_t go_tile;                // scout -> MMA: tiles whose barriers have all completed
};
static_assert(sizeof(Misc) <= 1024, "misc region overflow");

// Walks the tiles [t_begin, t_end) of a CTA in (image, strip, row) order.  One division at
// construction, none per tile: the per-tile critical path of the MMA warp is only ~2000 cycles long
// and an integer division costs ~100 of them.
struct TileIter {
  int t, t_begin, t_end, H, strips, n, strip, y;
  __device__ TileIter(const ConvParams& p, int b, int e) : t(b), t_begin(b), t_end(e), H(p.H), strips(p.strips) {
    const int col = b / p.H;
    y = b - col * p.H;
    n = col / p.strips;
    strip = col - n * p.strips;
  }
  __device__ __forceinline__ bool valid() const { return t < t_end; }
  __device__ __forceinline__ bool first() const { return t == t_begin || y == 0; }
  __device__ __forceinline__ bool last() const { return t == t_end - 1 || y == H - 1; }
  __device__ __forceinline__ int x0() const { return strip * kTilePx; }
  __device__ __forceinline__ void next() {
    ++t;
    if (++y == H) {
      y = 0;
      if (++strip == strips) {
        strip = 0;
        ++n;
      }
    }
  }
};

// Slot / phase-parity cursor of a ring of mbarrier-guarded buffers.
struct Ring {
  uint32_t slot, phase, size;
  __device__ explicit Ring(uint32_t n) : slot(0), phase(0), size(n) {}
  __device__ __forceinline__ void advance() {
    if (++slot == size) {
      slot = 0;
      phase ^= 1;
    }
  }
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == kActLrelu) return v > 0.f ? v : 0.1f * v;
  if (act == kActRelu) return fmaxf(v, 0.f);
  return v;
}

}  // namespace

template <bool kPar, bool kScale>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;            // 1024-byte aligned window
  uint8_t* sgen = smem_raw + (sbase - raw);
  const SmemLayout L = make_layout(p.n_wchunks, p.s_a, p.aux_k16 > 0, p.n_io);
  Misc* misc = reinterpret_cast<Misc*>(sgen + L.misc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_cta);
  const int s_a = p.s_a;
  const int n_io = p.n_io;

  // ---------------------------------------------------------------- setup
  if (threadIdx.x < 64) {
    misc->scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.0f;
    const int nb = (p.mode == kModeLast) ? 3 : 64;
    misc->bias[threadIdx.x] = (p.bias && threadIdx.x < nb) ? p.bias[threadIdx.x] : 0.0f;
  }
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&misc->w_full), 1);
      for (int i = 0; i < kMaxASlots; ++i) {
        mbar_init(smem_u32(&misc->a_full[i]), 1);
        mbar_init(smem_u32(&misc->a_empty[i]), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(smem_u32(&misc->aux_full[i]), 1);
        mbar_init(smem_u32(&misc->aux_empty[i]), 1);
        mbar_init(smem_u32(&misc->acc_full[i]), 1);
        mbar_init(smem_u32(&misc->acc_empty[i]), kEpilogueWarps);
      }
      for (int i = 0; i < kMaxIoSlots; ++i) {
        mbar_init(smem_u32(&misc->id_full[i]), 1);
        mbar_init(smem_u32(&misc->io_empty[i]), 1);
      }
      misc->go_tile = 0;
      mbar_fence_init();
      tma_prefetch_desc(&p.tm_src);
      if (p.mode != kModeLast) tma_prefetch_desc(&p.tm_out);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&misc->tmem_base), kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;
  // Programmatic dependent launch: launch latency, block scheduling and the prologue above overlap
  // the tail of the previous kernel in the stream; nothing below touches global memory before the
  // previous kernel has completed (packed weights may have been written by the kernel just before).
  griddep_launch_dependents();
  const uint32_t w_smem_early = sbase + L.w;
  auto load_weights = [&]() {                 // one elected lane of warp 0
    const uint32_t wbar = smem_u32(&misc->w_full);
    mbar_arrive_expect_tx(wbar, p.n_wchunks * kWChunkBytes);
    for (int c = 0; c < p.n_wchunks; ++c)
      bulk_load_1d(w_smem_early + c * kWChunkBytes,
                   reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)c * kWChunkBytes, kWChunkBytes, wbar);
  };
  // stable weights (packed long before this launch) are fetched while the previous kernel drains
  if (p.w_stable && warp == 0) {
    if (elect_one()) load_weights();
    __syncwarp();
  }
  griddep_wait();

  // patient waits for threads that merely wait for work (see pnp_conv_rows.cu); debug bit 32 = patient polling
  const bool hot_waits = (p.debug_skip & 32) == 0;
  const uint32_t hint_ns = (p.debug_skip & 64) ? 40u : 200u;
  auto pwait = [&](uint32_t bar, uint32_t parity, int tag) {      // one elected lane
    if (hot_waits) mbar_wait(bar, parity, tag); else mbar_wait_patient(bar, parity, tag, hint_ns);
  };
  auto ewait = [&](uint32_t bar, uint32_t parity, int tag) {      // whole (converged) warp
    if (hot_waits) mbar_wait(bar, parity, tag); else mbar_wait_warp(bar, parity, tag, hint_ns);
  };
  const uint32_t w_smem = sbase + L.w;
  const uint32_t a_smem = sbase + L.a;
  const uint32_t aux_smem = sbase + L.aux;
  const uint32_t io_smem = sbase + L.io;

  if (warp == 0) {
    // ============================================================ TMA producer (one elected lane)
    if (elect_one()) {
      if (!p.w_stable) load_weights();
      Ring ar(s_a), ior(n_io);
      uint32_t loads = 0;
      int it = 0;
      for (TileIter c(p, t_begin, t_end); c.valid(); c.next(), ++it) {
        for (int r = c.first() ? c.y - 1 : c.y + 1; r <= c.y + 1; ++r, ++loads, ar.advance()) {
          pwait(smem_u32(&misc->a_empty[ar.slot]), ar.phase ^ 1, 1);
          const uint32_t fb = smem_u32(&misc->a_full[ar.slot]);
          if ((p.debug_skip & 1) && loads >= (uint32_t)s_a) {
            mbar_arrive(fb);
          } else {
            mbar_arrive_expect_tx(fb, kRowBytes);
            tma_load_4d(a_smem + ar.slot * kASlotBytes, &p.tm_src, fb, 0, c.x0() - 1, r, c.n);
          }
        }
        if (p.aux_k16 > 0) {
          const uint32_t s = it & 1, ph = (it >> 1) & 1;
          pwait(smem_u32(&misc->aux_empty[s]), ph ^ 1, 2);
          const uint32_t fb = smem_u32(&misc->aux_full[s]);
          mbar_arrive_expect_tx(fb, kTileBytes);
          tma_load_4d(aux_smem + s * kTileBytes, &p.tm_aux, fb, 0, c.x0(), c.y, c.n);
        }
        if (p.has_id) {
          pwait(smem_u32(&misc->io_empty[ior.slot]), ior.phase ^ 1, 3);
          const uint32_t fb = smem_u32(&misc->id_full[ior.slot]);
          mbar_arrive_expect_tx(fb, kTileBytes);
          tma_load_4d(io_smem + ior.slot * kTileBytes, &p.tm_id, fb, 0, c.x0(), c.y, c.n);
          ior.advance();
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    // The whole warp walks the tile list and waits on the barriers (warp-uniform control flow keeps
    // addresses in uniform registers); one elected lane issues the 36..48 tcgen05.mma of a tile as
    // straight-line code: descriptor = per-row base word + compile-time offset.
    if (elect_one()) {
      const uint32_t idesc_center = umma_idesc_bf16(128, p.center_n);
      const uint32_t idesc_tap = umma_idesc_bf16(128, p.tap_n);
      const uint32_t center_chunks = (p.center_n == 256) ? 4 : 1;
      const uint32_t bo_mul = (p.base_off_mode == 1) ? (1u << 17) : 0u;   // diagnostic only
      const uint32_t w_lo = umma_desc_lo(w_smem);
      const uint32_t tap_lo = w_lo + center_chunks * (kWChunkBytes >> 4);
      mbar_wait(smem_u32(&misc->w_full), 0, 4);

      // Everything a tile's issue code needs.  The barriers of tile i+1 are checked in the middle
      // of tile i (an already-complete mbarrier wait costs ~200 cycles; the tensor pipe rides out
      // only ~300 cycles of silence from this thread -- tools/umma_bench.cu).
      struct TileCtx {
        bool valid, last;
        uint32_t s0, s1, s2, b, it;
      };
      Ring ar(s_a);
      TileIter ti(p, t_begin, t_end);
      const uint32_t go_tile = smem_u32(&misc->go_tile);
      auto acquire = [&](TileCtx& c, const TileCtx& prev) {   // ring slots of the tile at `ti`
        c.valid = ti.valid();
        if (!c.valid) return;
        if (ti.first()) {
          c.s0 = ar.slot;
          ar.advance();
          c.s1 = ar.slot;
          ar.advance();
        } else {
          c.s0 = prev.s1;
          c.s1 = prev.s2;
        }
        c.s2 = ar.slot;
        ar.advance();
        c.it = prev.it + 1;
        c.b = c.it & 1;
        c.last = ti.last();
        // every mbarrier this tile depends on (source rows, accumulator buffer, aux tile) has been
        // waited for by the scout warp; an acquire load of its counter costs ~30 cycles, an
        // already-complete mbarrier wait 220-290 in this kernel
        spin_until_ge(go_tile, c.it + 1, 5);
      };
      auto commit_tile = [&](const TileCtx& c) {
        umma_commit(smem_u32(&misc->a_empty[c.s0]));
        if (c.last) {
          umma_commit(smem_u32(&misc->a_empty[c.s1]));
          umma_commit(smem_u32(&misc->a_empty[c.s2]));
        }
        if (p.aux_k16 > 0) umma_commit(smem_u32(&misc->aux_empty[c.b]));
        umma_commit(smem_u32(&misc->acc_full[c.b]));
      };
      TileCtx none{false, false, 0, 0, 0, 0, 0xFFFFFFFFu};   // it + 1 == 0 for the first tile
      TileCtx cur;
      acquire(cur, none);
      tc_fence_after();
      bool pend = false;
      TileCtx pend_ctx = none;
      while (cur.valid) {
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && cur.it < 64;
        if (tr) p.trace[cur.it * 8 + 0] = clock64();
        const uint32_t d = tmem_base + cur.b * kAccStride;
        const uint32_t row_lo[3] = {umma_desc_lo(a_smem + cur.s0 * kASlotBytes),
                                    umma_desc_lo(a_smem + cur.s1 * kASlotBytes),
                                    umma_desc_lo(a_smem + cur.s2 * kASlotBytes)};
        TileCtx nxt;
        nxt.valid = false;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          const int tap = (j == 0) ? 4 : (j <= 4 ? j - 1 : j);   // centre first, then row-major
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const uint32_t a_lo = row_lo[dy + 1] + (dx + 1) * (128 >> 4);
          const uint32_t a_hi = kDescHiSw128 + bo_mul * (uint32_t)(dx + 1);
          const uint32_t b_lo = (j == 0) ? w_lo : tap_lo + (j - 1) * (kWChunkBytes >> 4);
          const uint32_t idesc = (j == 0) ? idesc_center : idesc_tap;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_lo(d, a_lo + 2 * k, a_hi, b_lo + 2 * k, kDescHiSw128, idesc, (j | k) != 0);
          if (j == 0 && pend) {
            // commits of the PREVIOUS tile ride behind this tile's first MMAs
            commit_tile(pend_ctx);
            pend = false;
          }
          if (j == 5 && !cur.last) {
            // barriers of the NEXT tile while ~8 MMAs of this one are queued.  (A tile that closes a
            // strip segment must first hand its three rows back -- the ring may be only 5 deep -- so
            // its look-ahead happens after the commit below.)
            ti.next();
            acquire(nxt, cur);
            tc_fence_after();
          }
        }
        if (p.aux_k16 > 0) {
          const uint32_t a_lo = umma_desc_lo(aux_smem + cur.b * kTileBytes);
          const uint32_t b_lo = tap_lo + 8 * (kWChunkBytes >> 4);
          for (int k = 0; k < p.aux_k16; ++k)
            umma_bf16_lo(d, a_lo + 2 * k, kDescHiSw128, b_lo + 2 * k, kDescHiSw128, idesc_tap, 1);
        }
        if (tr) p.trace[cur.it * 8 + 1] = clock64();
        if (cur.last) {
          // segment end: commit now (frees the three rows), then look ahead
          commit_tile(cur);
          pend = false;
          ti.next();
          acquire(nxt, cur);
          tc_fence_after();
        } else {
          pend = true;
          pend_ctx = cur;
        }
        cur = nxt;
      }
      if (pend) commit_tile(pend_ctx);
    }
  } else if (warp == 10) {
    // ============================================================ barrier scout (one elected lane)
    if (elect_one()) {
      const uint32_t go_tile = smem_u32(&misc->go_tile);
      Ring ar(s_a);
      uint32_t it = 0;
      for (TileIter c(p, t_begin, t_end); c.valid(); c.next(), ++it) {
        const int rows = c.first() ? 3 : 1;
        for (int r = 0; r < rows; ++r, ar.advance())
          mbar_wait(smem_u32(&misc->a_full[ar.slot]), ar.phase, 5);
        const uint32_t b = it & 1;
        mbar_wait(smem_u32(&misc->acc_empty[b]), ((it >> 1) & 1) ^ 1, 6);
        if (p.aux_k16 > 0) mbar_wait(smem_u32(&misc->aux_full[b]), (it >> 1) & 1, 7);
        st_release_shared(go_tile, it + 1);
      }
    }
  } else {
    // ============================================================ epilogue (8 warps, 256 threads)
    // A warp may only read the TMEM lane quarter (warp % 4); the two warps of a quarter split the
    // 64 output channels in halves.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;       // 0: channels 0..31, 1: channels 32..63
    const int row = q * 32 + lane;          // pixel inside the tile == TMEM lane
    const bool store_warp = (warp == 2);
    const uint32_t sw = (uint32_t)(row & 7);
    // per-channel epilogue constants of this warp's 32 channels live in registers for the whole
    // kernel (re-reading them from shared memory per tile costs as many smem wavefronts as the
    // output staging itself, and the shared-memory pipe is what bounds this kernel)
    float bias_r[32], scale_r[kScale ? 32 : 1];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      bias_r[j] = misc->bias[half * 32 + j];
      if (kScale) scale_r[j] = misc->scale[half * 32 + j];
    }
    Ring ior(n_io), rel(n_io);               // staging slot of this tile / slot to hand back
    int it = 0;
    for (TileIter c(p, t_begin, t_end); c.valid(); c.next(), ++it) {
      const int x = c.x0() + row;
      const bool valid = x < p.W;
      const uint32_t b = it & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * kAccStride;
      if (p.mode == kModeLast) {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (valid && half == 0) {
          const float* lp = p.lq + (long long)c.n * p.lq_sn + (long long)c.y * p.lq_sy + x;
          r0 = __ldg(lp);
          r1 = __ldg(lp + p.lq_sc);
          r2 = __ldg(lp + 2 * p.lq_sc);
        }
        ewait(smem_u32(&misc->acc_full[b]), (it >> 1) & 1, 9);
        tc_fence_after();
        float v[16];
        if (half == 0) {
          tmem_ld16(taddr, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        warp_arrive(smem_u32(&misc->acc_empty[b]));
        if (valid && half == 0) {
          float* op = p.outf + (long long)c.n * p.of_sn + (long long)c.y * p.of_sy + x;
          op[0] = v[0] + misc->bias[0] + r0;
          op[p.of_sc] = v[1] + misc->bias[1] + r1;
          op[2 * p.of_sc] = v[2] + misc->bias[2] + r2;
        }
        continue;
      }
      float p0 = 0.f, p1 = 0.f, p2 = 0.f;
      if (kPar && valid) {
        const float* pp = p.par + (long long)c.n * p.par_sn + (long long)c.y * p.par_sy + x;
        p0 = __ldg(pp);
        p1 = __ldg(pp + p.par_sc);
        p2 = __ldg(pp + 2 * p.par_sc);
      }
      const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && it < 64 && threadIdx.x == 64;
      if (tr) p.trace[it * 8 + 2] = clock64();
      const uint32_t s_io = ior.slot;
      // Staging-slot recycling.  All stores but the most recent one have finished READING shared
      // memory once wait_group.read<1> returns, i.e. the slots of tiles <= it-2 are free: hand the
      // slot of tile it-2 back to the producer now, so the identity tile of tile it-2+n_io is in
      // flight while this tile and the next are computed (its HBM latency is ~2 tiles long).
      if (store_warp) {
        if (elect_one()) {
          tma_store_wait_read<1>();
          if (p.has_id && it >= 2) mbar_arrive(smem_u32(&misc->io_empty[rel.slot]));
        }
        __syncwarp();
      }
      if (it >= 2) rel.advance();
      if (p.has_id) {
        ewait(smem_u32(&misc->id_full[s_io]), ior.phase, 8);
      } else {
        named_bar_sync(1, 256);             // n_io == 2: tile it-2's store has drained this slot
      }
      ewait(smem_u32(&misc->acc_full[b]), (it >> 1) & 1, 9);
      tc_fence_after();
      // only now touch the partition values: consuming them right after the loads would expose their global-memory
      // latency at the top of every tile instead of hiding it behind the wait above (measured: launch A +10 us)
      if (kPar && p.par_sparse) par_sparse_select(p0, p1, p2);
      if (tr) p.trace[it * 8 + 3] = clock64();
      uint8_t* rowp = sgen + L.io + s_io * kTileBytes + row * 128;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int g = half * 2 + gg;        // 16-channel group
        float v[16];
        if (p.debug_skip & 4) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        } else {
          tmem_ld16(taddr + g * 16, v);
        }
        if (kPar) {
          float a1[16], a2[16], a3[16];
          tmem_ld16(taddr + 64 + g * 16, a1);
          tmem_ld16(taddr + 128 + g * 16, a2);
          tmem_ld16(taddr + 192 + g * 16, a3);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = kScale ? fmaf(v[j], scale_r[gg * 16 + j], bias_r[gg * 16 + j]) : v[j] + bias_r[gg * 16 + j];
            v[j] = fmaf(p0, a1[j], v[j]);
            v[j] = fmaf(p1, a2[j], v[j]);
            v[j] = fmaf(p2, a3[j], v[j]);
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = kScale ? fmaf(v[j], scale_r[gg * 16 + j], bias_r[gg * 16 + j]) : v[j] + bias_r[gg * 16 + j];
        }
        uint4* c0 = reinterpret_cast<uint4*>(rowp + (((2 * g) ^ sw) << 4));
        uint4* c1 = reinterpret_cast<uint4*>(rowp + (((2 * g + 1) ^ sw) << 4));
        if (p.has_id) {
          const uint4 i0 = *c0, i1 = *c1;
          const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[2 * j] += bf16_lo(iw[j]);
            v[2 * j + 1] += bf16_hi(iw[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], p.act);
        uint4 o0, o1;
        o0.x = pack_bf16x2(v[0], v[1]);
        o0.y = pack_bf16x2(v[2], v[3]);
        o0.z = pack_bf16x2(v[4], v[5]);
        o0.w = pack_bf16x2(v[6], v[7]);
        o1.x = pack_bf16x2(v[8], v[9]);
        o1.y = pack_bf16x2(v[10], v[11]);
        o1.z = pack_bf16x2(v[12], v[13]);
        o1.w = pack_bf16x2(v[14], v[15]);
        if (!(p.debug_skip & 2)) {
          *c0 = o0;
          *c1 = o1;
        }
      }
      tc_fence_before();
      warp_arrive(smem_u32(&misc->acc_empty[b]));
      if (tr) p.trace[it * 8 + 4] = clock64();
      fence_proxy_async_smem();
      named_bar_sync(2, 256);
      if (tr) p.trace[it * 8 + 5] = clock64();
      if (store_warp) {
        if (elect_one()) {
          if (!(p.debug_skip & 2)) {
            tma_store_4d(&p.tm_out, io_smem + s_io * kTileBytes, 0, c.x0(), c.y, c.n);
            tma_store_commit();
          }
        }
        __syncwarp();
      }
      ior.advance();
    }
    if (store_warp) {
      if (elect_one()) tma_store_wait_all<0>();
      __syncwarp();
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

size_t conv_smem_bytes(const ConvParams& p) {
  return make_layout(p.n_wchunks, p.s_a, p.aux_k16 > 0, p.n_io).total + 1024;  // + alignment slack
}

namespace {
template <bool kPar, bool kScale>
cudaError_t launch_variant(const ConvParams& p, int grid, size_t smem, cudaStream_t stream) {
  // opt in to the full 227 KB once per device (one device per process is the deployment model,
  // but several are tolerated)
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(conv3x3_umma_kernel<kPar, kScale>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             232448);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kRowsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<kPar, kScale>, p);
}
}  // namespace

cudaError_t launch_conv(const ConvParams& p, int grid, cudaStream_t stream) {
  const size_t smem = conv_smem_bytes(p);
  const bool par = (p.center_n == 256), scale = (p.scale != nullptr);
  if (par) return scale ? launch_variant<true, true>(p, grid, smem, stream)
                        : launch_variant<true, false>(p, grid, smem, stream);
  return scale ? launch_variant<false, true>(p, grid, smem, stream)
               : launch_variant<false, false>(p, grid, smem, stream);
}

}  // namespace pnp
