// Row-stacked variant of the tcgen05 3x3 convolution: one source row feeds THREE output rows.
//
// Why (measured, profiles/r01_notes.md): an M=128,K=16 tcgen05.mma costs max(N/2, (4096+32N)/128)
// cycles, so the N=64 MMAs of the tap-major kernel (pnp_conv.cu) are bound by the shared-memory
// operand pipe at 64 % of the tensor peak.  Here the A operand is still "128 pixels of source row r,
// shifted by dx", but B stacks the three dy weight blocks [W(+1,dx); W(0,dx); W(-1,dx)] so a single
// N=192 MMA (~98 cycles instead of 3 x 50) adds row r's contribution to the accumulators of output
// rows r-1, r and r+1 at once.  Accumulators of eight consecutive output rows live in a TMEM ring
// (8 x 64 columns = all 512), each source row is staged in shared memory for exactly one step, and
// an output row is finished -- and handed to the epilogue -- one step after its own source row.
//
// Same reference semantics as pnp_conv.cu (F.conv2d + bias/activation/identity/LR-aux/+lq fused);
// the partition-modulated block launch (center_n == 256) stays on the tap-major kernel for now.
#include "pnp_conv.cuh"
#include "pnp_ptx.cuh"

namespace pnp {

namespace {

constexpr int kAccRing = 8;      // output-row accumulators in TMEM
constexpr int kStepRing = 8;     // step-completion barriers

struct RowsLayout {
  uint32_t w, a, aux, io, misc, total;
};

__host__ __device__ inline RowsLayout rows_layout(int w_bytes, int s_a, int has_aux, int n_io) {
  RowsLayout l;
  l.w = 0;
  l.a = l.w + ((w_bytes + 1023) & ~1023);
  l.aux = l.a + s_a * kASlotBytes;
  l.io = l.aux + (has_aux ? 2 * kTileBytes : 0);
  l.misc = l.io + n_io * kTileBytes;
  l.total = l.misc + 1024;
  return l;
}

struct RowsMisc {
  float scale[64];
  float bias[64];
  uint64_t w_full;
  uint64_t a_full[kMaxASlots];
  uint64_t step_done[kStepRing];   // tcgen05.commit after every step (one source row)
  uint64_t acc_free[kAccRing];     // epilogue -> MMA: accumulator slot drained
  uint64_t aux_full[2];
  uint64_t id_full[kMaxIoSlots];
  uint64_t io_empty[kMaxIoSlots];
  uint32_t tmem_base;
};
static_assert(sizeof(RowsMisc) <= 1024, "misc region overflow");

// A CTA owns the tiles [t_begin, t_end) in (image, strip, row) order; a segment is a maximal run of
// consecutive rows of one strip.  Steps j = j_first..j_last are the in-image source rows y_b + j.
struct Segment {
  int n, strip, y_b, len, j_first, j_last;
};

struct SegIter {
  int t, t_end, H, strips, n, strip, y_b;
  __device__ SegIter(const ConvParams& p, int b, int e) : t(b), t_end(e), H(p.H), strips(p.strips) {
    const int col = b / p.H;
    y_b = b - col * p.H;
    n = col / p.strips;
    strip = col - n * p.strips;
  }
  __device__ __forceinline__ bool valid() const { return t < t_end; }
  __device__ __forceinline__ Segment get() const {
    Segment s;
    s.n = n;
    s.strip = strip;
    s.y_b = y_b;
    s.len = min(H - y_b, t_end - t);
    s.j_first = (y_b > 0) ? -1 : 0;
    s.j_last = (y_b + s.len < H) ? s.len : s.len - 1;
    return s;
  }
  __device__ __forceinline__ void next(const Segment& s) {
    t += s.len;
    y_b = 0;
    if (++strip == strips) {
      strip = 0;
      ++n;
    }
  }
};

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

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == kActLrelu) return v > 0.f ? v : 0.1f * v;
  if (act == kActRelu) return fmaxf(v, 0.f);
  return v;
}

}  // namespace

template <bool kScale>
__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_rows_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int tap_n = p.tap_n;                            // 64, or 16 for the 64->3 tail
  const int dx_block_bytes = 3 * tap_n * 128;           // [3 dy sub-blocks][tap_n rows][128 B]
  const int w_bytes = 3 * dx_block_bytes + (p.aux_k16 > 0 ? kWChunkBytes : 0);
  const RowsLayout L = rows_layout(w_bytes, p.s_a, p.aux_k16 > 0, p.n_io);
  RowsMisc* misc = reinterpret_cast<RowsMisc*>(sgen + L.misc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_cta);
  const int s_a = p.s_a;
  const int n_io = p.n_io;
  const bool last_mode = (p.mode == kModeLast);

  if (threadIdx.x < 64) {
    misc->scale[threadIdx.x] = p.scale ? p.scale[threadIdx.x] : 1.0f;
    const int nb = last_mode ? 3 : 64;
    misc->bias[threadIdx.x] = (p.bias && threadIdx.x < nb) ? p.bias[threadIdx.x] : 0.0f;
  }
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&misc->w_full), 1);
      for (int i = 0; i < kMaxASlots; ++i) mbar_init(smem_u32(&misc->a_full[i]), 1);
      for (int i = 0; i < kStepRing; ++i) mbar_init(smem_u32(&misc->step_done[i]), 1);
      for (int i = 0; i < kAccRing; ++i) mbar_init(smem_u32(&misc->acc_free[i]), kConvThreads - 64);
      for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&misc->aux_full[i]), 1);
      for (int i = 0; i < kMaxIoSlots; ++i) {
        mbar_init(smem_u32(&misc->id_full[i]), 1);
        mbar_init(smem_u32(&misc->io_empty[i]), 1);
      }
      mbar_fence_init();
      tma_prefetch_desc(&p.tm_src);
      if (!last_mode) tma_prefetch_desc(&p.tm_out);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&misc->tmem_base), kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;
  const uint32_t w_smem = sbase + L.w;
  const uint32_t a_smem = sbase + L.a;
  const uint32_t aux_smem = sbase + L.aux;
  const uint32_t io_smem = sbase + L.io;

  if (warp == 0) {
    // ============================================================ TMA producer (one elected lane)
    if (elect_one()) {
      const uint32_t wbar = smem_u32(&misc->w_full);
      mbar_arrive_expect_tx(wbar, w_bytes);
      for (int off = 0; off < w_bytes; off += kWChunkBytes) {
        const int n = min(kWChunkBytes, w_bytes - off);
        bulk_load_1d(w_smem + off, reinterpret_cast<const uint8_t*>(p.wpack) + off, n, wbar);
      }
      Ring ar(s_a), ior(n_io);
      uint32_t sc = 0, ord = 0;            // step counter, output-row ordinal
      uint32_t aux_step[2] = {0, 0};       // step in which each aux slot was last consumed
      for (SegIter it(p, t_begin, t_end); it.valid();) {
        const Segment s = it.get();
        const int x0 = s.strip * kTilePx;
        for (int j = s.j_first; j <= s.j_last; ++j, ++sc, ar.advance()) {
          if (sc >= (uint32_t)s_a) {       // slot last used by step sc - s_a
            const uint32_t ps = sc - s_a;
            mbar_wait(smem_u32(&misc->step_done[ps & (kStepRing - 1)]), (ps >> 3) & 1, 1);
          }
          const uint32_t fb = smem_u32(&misc->a_full[ar.slot]);
          if ((p.debug_skip & 1) && sc >= (uint32_t)s_a) {
            mbar_arrive(fb);
          } else {
            mbar_arrive_expect_tx(fb, kRowBytes);
            tma_load_4d(a_smem + ar.slot * kASlotBytes, &p.tm_src, fb, 0, x0 - 1, s.y_b + j, s.n);
          }
          if (j >= 0 && j < s.len) {       // per-output-row operands of row y_b + j
            if (p.aux_k16 > 0) {
              const uint32_t as = ord & 1;
              if (ord >= 2) {
                const uint32_t ps = aux_step[as];
                mbar_wait(smem_u32(&misc->step_done[ps & (kStepRing - 1)]), (ps >> 3) & 1, 2);
              }
              aux_step[as] = sc;           // consumed in this very step (centre row)
              const uint32_t ab = smem_u32(&misc->aux_full[as]);
              mbar_arrive_expect_tx(ab, kTileBytes);
              tma_load_4d(aux_smem + as * kTileBytes, &p.tm_aux, ab, 0, x0, s.y_b + j, s.n);
            }
            if (p.has_id) {
              mbar_wait(smem_u32(&misc->io_empty[ior.slot]), ior.phase ^ 1, 3);
              const uint32_t ib = smem_u32(&misc->id_full[ior.slot]);
              mbar_arrive_expect_tx(ib, kTileBytes);
              tma_load_4d(io_smem + ior.slot * kTileBytes, &p.tm_id, ib, 0, x0, s.y_b + j, s.n);
              ior.advance();
            }
            ++ord;
          }
        }
        it.next(s);
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc0 = umma_idesc_bf16(128, 0);                  // N field added per MMA
      const uint32_t idesc_step = ((uint32_t)tap_n >> 3) << 17;         // one dy sub-block of N
      const uint32_t w_lo = umma_desc_lo(w_smem);
      const uint32_t dxb = (uint32_t)dx_block_bytes >> 4;      // descriptor units per dx block
      const uint32_t sbb = (uint32_t)(tap_n * 128) >> 4;       // ... per dy sub-block
      const uint32_t aux_w_lo = w_lo + 3 * dxb;
      mbar_wait(smem_u32(&misc->w_full), 0, 4);
      // One step = one in-image source row.  `StepCtx` carries everything the issue code needs, so
      // the barriers of step s+1 can be checked in the MIDDLE of step s: an already-complete
      // mbarrier wait costs ~200 cycles and the tensor pipe only rides out ~300 cycles of silence
      // from this thread (tools/umma_bench.cu, "steps" rows), so waits between steps would stall it.
      struct StepCtx {
        bool valid;
        int j, len, j_first;
        uint32_t ord0, sc, a_slot, a_phase;
      };
      SegIter seg_it(p, t_begin, t_end);
      Segment seg = seg_it.valid() ? seg_it.get() : Segment{0, 0, 0, 0, 0, -1};
      Ring ar(s_a);
      StepCtx cur{seg_it.valid(), seg.j_first, seg.len, seg.j_first, 0u, 0u, ar.slot, ar.phase};
      auto advance = [&](StepCtx& c) {        // next step in program order (crosses segments)
        ar.advance();
        c.sc += 1;
        c.a_slot = ar.slot;
        c.a_phase = ar.phase;
        if (c.j < seg.j_last) {
          c.j += 1;
          return;
        }
        const uint32_t next_ord0 = c.ord0 + (uint32_t)seg.len;
        seg_it.next(seg);
        c.valid = seg_it.valid();
        if (c.valid) {
          seg = seg_it.get();
          c.j = seg.j_first;
          c.len = seg.len;
          c.j_first = seg.j_first;
          c.ord0 = next_ord0;
        }
      };
      auto ranges = [](const StepCtx& c, int& lo, int& cnt, int& old_cnt) {
        lo = max(c.j - 1, 0);
        const int hi = min(c.j + 1, c.len - 1);
        cnt = hi - lo + 1;
        const int new_from = (c.j == c.j_first) ? lo : c.j + 1;   // rows first touched in this step
        old_cnt = min(max(new_from - lo, 0), cnt);
      };
      auto wait_for = [&](const StepCtx& c) {
        mbar_wait(smem_u32(&misc->a_full[c.a_slot]), c.a_phase, 5);
        int lo, cnt, old_cnt;
        ranges(c, lo, cnt, old_cnt);
        for (int o = lo + old_cnt; o < lo + cnt; ++o) {          // new rows: slot must be drained
          const uint32_t od = c.ord0 + o;
          mbar_wait(smem_u32(&misc->acc_free[od & (kAccRing - 1)]), ((od >> 3) & 1) ^ 1, 6);
        }
        if (p.aux_k16 > 0 && c.j >= 0 && c.j < c.len) {
          const uint32_t od = c.ord0 + c.j;
          mbar_wait(smem_u32(&misc->aux_full[od & 1]), (od >> 1) & 1, 7);
        }
      };
      bool pend = false;
      uint32_t pend_bar = 0;

      // MMAs of one (dx,k) over `cnt` consecutive accumulator slots starting at slot_lo (wraps)
      auto mma_range = [&](uint32_t slot_lo, int cnt, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
        const int n1 = min(cnt, kAccRing - (int)slot_lo);
        umma_bf16_lo(tmem_base + slot_lo * tap_n, a_lo, kDescHiSw128, b_lo, kDescHiSw128, idesc0 + n1 * idesc_step, acc);
        if (cnt > n1)
          umma_bf16_lo(tmem_base, a_lo, kDescHiSw128, b_lo + n1 * sbb, kDescHiSw128, idesc0 + (cnt - n1) * idesc_step, acc);
      };

      if (cur.valid) wait_for(cur);
      tc_fence_after();
      while (cur.valid) {
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && cur.sc < 64;
        if (tr) p.trace[cur.sc * 8 + 0] = clock64();
        int lo, cnt, old_cnt;
        ranges(cur, lo, cnt, old_cnt);
        const int new_cnt = cnt - old_cnt;
        const bool centre = (cur.j >= 0 && cur.j < cur.len);
        const uint32_t slot_lo = (cur.ord0 + lo) & (kAccRing - 1);
        const uint32_t a_row = umma_desc_lo(a_smem + cur.a_slot * kASlotBytes);
        const uint32_t b_row = w_lo + (uint32_t)(lo - (cur.j - 1)) * sbb;   // first dy sub-block in range
        const uint32_t cur_sc = cur.sc;
        const uint32_t cur_od = cur.ord0 + (uint32_t)max(cur.j, 0);
        StepCtx nxt = cur;
        advance(nxt);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          if (dx == 2) {
            // barriers of the NEXT step, checked while ~8 MMAs of this step are still queued
            if (nxt.valid) wait_for(nxt);
            tc_fence_after();
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t a_lo = a_row + dx * 8 + 2 * k;
            const uint32_t b_lo = b_row + dx * dxb + 2 * k;
            if (dx == 0 && k == 0) {
              // first MMA of the step: rows touched before accumulate, new rows are overwritten
              if (old_cnt > 0) mma_range(slot_lo, old_cnt, a_lo, b_lo, 1);
              if (new_cnt > 0)
                mma_range((slot_lo + old_cnt) & (kAccRing - 1), new_cnt, a_lo, b_lo + old_cnt * sbb, 0);
              if (pend) {
                // the previous step's commit rides behind this step's first MMA
                umma_commit(pend_bar);
                pend = false;
              }
            } else {
              mma_range(slot_lo, cnt, a_lo, b_lo, 1);
            }
          }
        }
        if (p.aux_k16 > 0 && centre) {
          const uint32_t a_lo = umma_desc_lo(aux_smem + (cur_od & 1) * kTileBytes);
          for (int k = 0; k < p.aux_k16; ++k)
            umma_bf16_lo(tmem_base + (cur_od & (kAccRing - 1)) * tap_n, a_lo + 2 * k, kDescHiSw128,
                         aux_w_lo + 2 * k, kDescHiSw128, idesc0 + idesc_step, 1);
        }
        pend = true;
        pend_bar = smem_u32(&misc->step_done[cur_sc & (kStepRing - 1)]);
        if (tr) p.trace[cur_sc * 8 + 1] = clock64();
        cur = nxt;
      }
      if (pend) umma_commit(pend_bar);
    }
  } else {
    // ============================================================ epilogue (8 warps, 256 threads)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool store_warp = (warp == 2);
    const uint32_t sw = (uint32_t)(row & 7);
    float bias_r[32], scale_r[kScale ? 32 : 1];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      bias_r[j] = misc->bias[half * 32 + j];
      if (kScale) scale_r[j] = misc->scale[half * 32 + j];
    }
    Ring ior(n_io), rel(n_io);
    uint32_t ord = 0, sc0 = 0;
    for (SegIter it(p, t_begin, t_end); it.valid();) {
      const Segment s = it.get();
      const int x = s.strip * kTilePx + row;
      const bool valid = x < p.W;
      for (int o = 0; o < s.len; ++o, ++ord) {
        const int y = s.y_b + o;
        const uint32_t sc_last = sc0 + (uint32_t)(min(o + 1, s.j_last) - s.j_first);
        const uint32_t slot = ord & (kAccRing - 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * tap_n;
        if (last_mode) {
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          if (valid && half == 0) {
            const float* lp = p.lq + (long long)s.n * p.lq_sn + (long long)y * p.lq_sy + x;
            r0 = __ldg(lp);
            r1 = __ldg(lp + p.lq_sc);
            r2 = __ldg(lp + 2 * p.lq_sc);
          }
          mbar_wait(smem_u32(&misc->step_done[sc_last & (kStepRing - 1)]), (sc_last >> 3) & 1, 9);
          tc_fence_after();
          float v[16];
          if (half == 0) {
            tmem_ld16(taddr, v);
            tmem_ld_wait();
          }
          tc_fence_before();
          mbar_arrive(smem_u32(&misc->acc_free[slot]));
          if (valid && half == 0) {
            float* op = p.outf + (long long)s.n * p.of_sn + (long long)y * p.of_sy + x;
            op[0] = v[0] + misc->bias[0] + r0;
            op[p.of_sc] = v[1] + misc->bias[1] + r1;
            op[2 * p.of_sc] = v[2] + misc->bias[2] + r2;
          }
          continue;
        }
        const bool tr = (p.trace != nullptr) && blockIdx.x == 0 && ord < 64 && threadIdx.x == 64;
        if (tr) p.trace[ord * 8 + 2] = clock64();
        const uint32_t s_io = ior.slot;
        if (store_warp) {
          if (elect_one()) {
            tma_store_wait_read<1>();      // stores of output rows <= ord-2 no longer read smem
            if (p.has_id && ord >= 2) mbar_arrive(smem_u32(&misc->io_empty[rel.slot]));
          }
          __syncwarp();
        }
        if (ord >= 2) rel.advance();
        if (p.has_id) {
          mbar_wait(smem_u32(&misc->id_full[s_io]), ior.phase, 8);
        } else {
          named_bar_sync(1, 256);
        }
        mbar_wait(smem_u32(&misc->step_done[sc_last & (kStepRing - 1)]), (sc_last >> 3) & 1, 9);
        tc_fence_after();
        if (tr) p.trace[ord * 8 + 3] = clock64();
        uint8_t* rowp = sgen + L.io + s_io * kTileBytes + row * 128;
        float v[32];
        if (p.debug_skip & 4) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
          tmem_ld16(taddr + half * 32, v);
          tmem_ld16(taddr + half * 32 + 16, v + 16);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&misc->acc_free[slot]));   // accumulator is in registers: slot reusable
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = half * 2 + gg;
          float* vv = v + gg * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            vv[j] = kScale ? fmaf(vv[j], scale_r[gg * 16 + j], bias_r[gg * 16 + j]) : vv[j] + bias_r[gg * 16 + j];
          uint4* c0 = reinterpret_cast<uint4*>(rowp + (((2 * g) ^ sw) << 4));
          uint4* c1 = reinterpret_cast<uint4*>(rowp + (((2 * g + 1) ^ sw) << 4));
          if (p.has_id) {
            const uint4 i0 = *c0, i1 = *c1;
            const uint32_t iw[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              vv[2 * j] += bf16_lo(iw[j]);
              vv[2 * j + 1] += bf16_hi(iw[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) vv[j] = act_fn(vv[j], p.act);
          uint4 o0, o1;
          o0.x = pack_bf16x2(vv[0], vv[1]);
          o0.y = pack_bf16x2(vv[2], vv[3]);
          o0.z = pack_bf16x2(vv[4], vv[5]);
          o0.w = pack_bf16x2(vv[6], vv[7]);
          o1.x = pack_bf16x2(vv[8], vv[9]);
          o1.y = pack_bf16x2(vv[10], vv[11]);
          o1.z = pack_bf16x2(vv[12], vv[13]);
          o1.w = pack_bf16x2(vv[14], vv[15]);
          if (!(p.debug_skip & 2)) {
            *c0 = o0;
            *c1 = o1;
          }
        }
        if (tr) p.trace[ord * 8 + 4] = clock64();
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (tr) p.trace[ord * 8 + 5] = clock64();
        if (store_warp) {
          if (elect_one()) {
            if (!(p.debug_skip & 2)) {
              tma_store_4d(&p.tm_out, io_smem + s_io * kTileBytes, 0, s.strip * kTilePx, y, s.n);
              tma_store_commit();
            }
          }
          __syncwarp();
        }
        ior.advance();
      }
      sc0 += (uint32_t)(s.j_last - s.j_first + 1);
      it.next(s);
    }
    if (store_warp) {
      if (elect_one()) tma_store_wait_all<0>();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

size_t conv_rows_smem_bytes(const ConvParams& p) {
  const int w_bytes = 3 * 3 * p.tap_n * 128 + (p.aux_k16 > 0 ? kWChunkBytes : 0);
  return rows_layout(w_bytes, p.s_a, p.aux_k16 > 0, p.n_io).total + 1024;
}

namespace {
template <bool kScale>
cudaError_t launch_rows_variant(const ConvParams& p, int grid, size_t smem, cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(conv3x3_rows_kernel<kScale>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  conv3x3_rows_kernel<kScale><<<grid, kConvThreads, smem, stream>>>(p);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_conv_rows(const ConvParams& p, int grid, cudaStream_t stream) {
  const size_t smem = conv_rows_smem_bytes(p);
  return p.scale != nullptr ? launch_rows_variant<true>(p, grid, smem, stream)
                            : launch_rows_variant<false>(p, grid, smem, stream);
}

}  // namespace pnp
