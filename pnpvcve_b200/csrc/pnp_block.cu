// One BAE residual block per launch, fused across a CTA PAIR (thread-block cluster of 2).
//
// Replaces ResidualBlockNoBNDynamic_drt.forward (mmedit/models/common/sr_backbone_utils.py:304-333):
//     t   = relu( conv3x3(x, g*Wmix) + g*bmix + sum_k par_k * conv1x1_k(x) )        (:307-311)
//     out = x + conv3x3(t, W1) + b1                                                  (:313-316)
// which pnp_conv.cu / pnp_conv_rows.cu run as two launches with `t` round-tripping through HBM (the
// layer-by-layer form is HBM-bound: 590 MB per block at 720p).  Here `t` never leaves the chip:
//
//   * the two convolutions need 72 + 24 + 72 KB of resident weights and 5 + 3 + 5 TMEM accumulator
//     slots -- more than one SM has -- so the block is split over the two SMs of a cluster:
//       rank 0 ("stage 1"): TMA-loads source rows of x, row-stacked conv2 (N=192 MMAs into a ring of five
//                64-column accumulators) + the three partition 1x1 convs (one N=192 group into a 192-column
//                region), epilogue = bias + partition blend + ReLU -> bf16 row of `t` staged in its own
//                shared memory and pushed into the partner's ring with cp.async.bulk (DSMEM);
//       rank 1 ("stage 2"): row-stacked conv1 on the rows of `t` arriving in its ring, epilogue =
//                bias + identity (x re-read through L2) -> TMA store of the output row.
//   * a strip is 126 output pixels wide: stage 1 computes 128 pixels of `t` (x0-1 .. x0+126) from 130
//     source pixels, stage 2 computes 128 outputs of which 126 are valid; pixels of `t` outside the image
//     are forced to zero (they are conv1's zero padding).  At the top/bottom of a CTA's row range stage 1
//     computes one extra row of `t` (halo recompute, ~2 % at 720p).
//   * both CTAs walk the same (image, strip, row) ranges; hand-offs are mbarriers: t_full (bulk-copy
//     complete_tx on the consumer), stage_free / t_free (remote arrives back to the producer).
//   * each CTA keeps the warp roles of pnp_conv_rows.cu (TMA producer / MMA issuer / barrier scout /
//     8 epilogue warps), but the epilogue is two independent groups of four warps working on alternate
//     rows, so a row's epilogue may take two MMA steps.
#include "pnp_block.cuh"
#include "pnp_conv.cuh"
#include "pnp_ptx.cuh"

namespace pnp {

namespace {

// Eight epilogue warps, not sixteen: measured (tools/umma_interf.cu, profiles/r01_umma_interf_alu.log) the warp
// scheduler does not favour the MMA-issuing thread -- four ALU-busy warps on its sub-partition slow MMA issue
// 4.2x, two by 19 %, one not at all.
constexpr int kEpiWarps = 8;
constexpr int kEpiCh = 64 / (kEpiWarps / 4);   // channels per epilogue thread
constexpr int kBlockThreads = 128 + 32 * kEpiWarps;   // warps 0 TMA / relay, 1 MMA, 2 scout, 3 copy / store issuer, 4.. epilogue
constexpr int kCtrlRegs = 128;       // setmaxnreg: the control warpgroup (warps 0-3) ...
constexpr int kEpiRegs = 184;        // ... vs the epilogue warpgroups (launch allocation: 168 per thread)
constexpr int kRing0 = 5;        // stage-1 accumulator ring (5 x 64 TMEM columns)
constexpr int kRing1 = 8;        // stage-2 accumulator ring (8 x 64 TMEM columns)
constexpr int kParCol = 320;     // stage 1: TMEM column of the partition 1x1 accumulators (3 x 64)
constexpr int kStepRing = 8;
constexpr int kMaxSlots = 8;
constexpr uint32_t kDxb = 3 * 64 * 128 / 16;   // descriptor units per dx block of a row-stacked pack
constexpr uint32_t kSbb = 64 * 128 / 16;       // ... per dy sub-block

struct PairMisc {
  float bias[64];                  // stage 1: g*bmix; stage 2: conv1 bias
  uint64_t w_full;
  uint64_t a_full[kMaxSlots];      // stage 1: source rows (TMA); stage 2: rows of t (partner's bulk copies)
  uint64_t step_done[kStepRing];   // tcgen05.commit after every step (one input row)
  uint64_t acc_free[kMaxSlots];    // epilogue -> MMA: accumulator slot drained
  uint64_t t_free[kMaxSlots];      // stage 1: the partner's MMAs are done with ring slot i (remote arrive)
  uint64_t stage_free[2];          // stage 1: the copy out of staging tile b has landed (remote arrive)
  uint64_t staged[2];              // stage 1: all eight epilogue warps have written staging tile b
  uint64_t par_done;               // stage 1: the partition accumulators of a row are complete (own commit)
  uint32_t tmem_base;
  uint32_t go_step;
  uint32_t go_par;                 // stage 1: epilogue-warp reads of the partition region (4 per row of t)
};
static_assert(sizeof(PairMisc) <= 1024, "misc region overflow");

struct PairLayout {
  uint32_t misc, w, ring, stage, total;
};

__host__ __device__ inline PairLayout pair_layout(int role, int s_a, int n_t) {
  PairLayout l;
  l.misc = 0;
  l.w = 1024;
  l.ring = l.w + (role == 0 ? kBlockW0Bytes : kBlockW1Bytes);
  l.stage = l.ring + (role == 0 ? s_a : n_t) * kASlotBytes;
  l.total = l.stage + 2 * kTileBytes;
  return l;
}

// A pair owns the output tiles [t_begin, t_end) in (image, strip, row) order; a segment is a maximal
// run of consecutive rows of one strip.  `len` rows starting at y_b are produced from the input rows
// y_b + j, j = j_first..j_last (the rows above/below exist unless the segment touches the image edge).
struct Segment {
  int n, strip, y_b, len, j_first, j_last;
};

__device__ __forceinline__ Segment make_seg(int n, int strip, int y_b, int len, int H) {
  Segment s;
  s.n = n;
  s.strip = strip;
  s.y_b = y_b;
  s.len = len;
  s.j_first = (y_b > 0) ? -1 : 0;
  s.j_last = (y_b + len < H) ? len : len - 1;
  return s;
}

// ext = false: segments of OUTPUT rows (stage 2; its input rows are rows of t).
// ext = true : the rows of t stage 2 needs for that segment, i.e. the segment grown by its halo rows
//              (stage 1; its input rows are rows of x).
struct SegIter {
  int t, t_end, H, strips, n, strip, y_b;
  bool ext;
  __device__ SegIter(const BlockParams& p, int b, int e, bool ext_)
      : t(b), t_end(e), H(p.H), strips(p.strips), ext(ext_) {
    const int col = b / p.H;
    y_b = b - col * p.H;
    n = col / p.strips;
    strip = col - n * p.strips;
  }
  __device__ __forceinline__ bool valid() const { return t < t_end; }
  __device__ __forceinline__ Segment get() const {
    Segment s = make_seg(n, strip, y_b, min(H - y_b, t_end - t), H);
    if (ext) s = make_seg(n, strip, y_b + s.j_first, s.j_last - s.j_first + 1, H);
    return s;
  }
  __device__ __forceinline__ void next() {
    t += min(H - y_b, t_end - t);
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

// flat cursor over the rows a CTA produces
struct RowCur {
  SegIter it;
  Segment s;
  int o;
  uint32_t ord, sc0;
  bool valid;
  __device__ RowCur(const BlockParams& p, int b, int e, bool ext)
      : it(p, b, e, ext), s{0, 0, 0, 0, 0, -1}, o(0), ord(0), sc0(0), valid(false) {
    valid = it.valid();
    if (valid) s = it.get();
  }
  __device__ __forceinline__ void next() {
    ++ord;
    if (++o < s.len) return;
    sc0 += (uint32_t)(s.j_last - s.j_first + 1);
    it.next();
    valid = it.valid();
    o = 0;
    if (valid) s = it.get();
  }
  // step whose completion finishes the 3x3 accumulators of this row
  __device__ __forceinline__ uint32_t sc_last() const { return sc0 + (uint32_t)(min(o + 1, s.j_last) - s.j_first); }
};

// ---------------------------------------------------------------------------------------- MMA issuer
// Same issue discipline as pnp_conv_rows.cu (see the notes there): straight-line tcgen05.mma with
// descriptor = base word + immediate, barriers of the next step polled through the scout's counter in
// the middle of the current step, commits deferred behind the next step's first MMA.
// One INTERIOR step (source row j with 1 <= j <= len-2: three output rows, two already touched, one new;
// next step in the same segment) as straight-line code.  The MMA-issuing thread shares its sub-partition's
// scheduler with epilogue warps and is not favoured by it (tools/umma_interf.cu), so every instruction it
// does not execute is tensor-pipe time: all offsets are immediates, the only runtime inputs are the A-row
// descriptor word, the accumulator column of the window and the flags to poll.  N1 = rows of the 3-row
// window before the accumulator ring wraps (3 = no wrap).
template <bool kPar, int N1>
__device__ __forceinline__ void mma_fast_step(uint32_t tmem_base, uint32_t win_col, uint32_t a_row, uint32_t w_lo,
                                              uint32_t go_step, uint32_t next_target, uint32_t go_par,
                                              uint32_t par_target, uint32_t pend_bar, uint32_t par_done_bar) {
  constexpr uint32_t kI64 = umma_idesc_bf16(128, 64), kI128 = umma_idesc_bf16(128, 128),
                     kI192 = umma_idesc_bf16(128, 192);
  const uint32_t d0 = tmem_base + win_col;
  // first MMA of the step: rows 0,1 of the window accumulate, row 2 is overwritten
  if (N1 >= 2) {
    umma_bf16_lo(d0, a_row, kDescHiSw128, w_lo, kDescHiSw128, kI128, 1);
    umma_bf16_lo(N1 == 3 ? d0 + 128 : tmem_base, a_row, kDescHiSw128, w_lo + 2 * kSbb, kDescHiSw128, kI64, 0);
  } else {
    umma_bf16_lo(d0, a_row, kDescHiSw128, w_lo, kDescHiSw128, kI64, 1);
    umma_bf16_lo(tmem_base, a_row, kDescHiSw128, w_lo + kSbb, kDescHiSw128, kI64, 1);
    umma_bf16_lo(tmem_base + 64, a_row, kDescHiSw128, w_lo + 2 * kSbb, kDescHiSw128, kI64, 0);
  }
  if (pend_bar != 0) umma_commit(pend_bar);   // the previous step's commit rides behind this step's first MMA
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
    if (dx == 2) {
      spin_until_ge(go_step, next_target, 5);  // barriers of the NEXT step; normally long satisfied
      tc_fence_after();
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (dx == 0 && k == 0) continue;
      const uint32_t a_lo = a_row + dx * 8 + 2 * k;
      const uint32_t b_lo = w_lo + dx * kDxb + 2 * k;
      if (N1 == 3) {
        umma_bf16_lo(d0, a_lo, kDescHiSw128, b_lo, kDescHiSw128, kI192, 1);
      } else {
        umma_bf16_lo(d0, a_lo, kDescHiSw128, b_lo, kDescHiSw128, N1 == 2 ? kI128 : kI64, 1);
        umma_bf16_lo(tmem_base, a_lo, kDescHiSw128, b_lo + N1 * kSbb, kDescHiSw128, N1 == 2 ? kI64 : kI128, 1);
      }
    }
  }
  if (kPar) {
    spin_until_ge(go_par, par_target, 10);
    tc_fence_after();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_lo(tmem_base + kParCol, a_row + 8 + 2 * k, kDescHiSw128, w_lo + 3 * kDxb + 2 * k, kDescHiSw128, kI192,
                   k > 0);
    umma_commit(par_done_bar);
  }
}

template <bool kPar, int kAccRing, bool kTrace>
__device__ __forceinline__ void mma_issue_loop(const BlockParams& p, PairMisc* misc, uint32_t w_smem,
                                               uint32_t a_smem, uint32_t ring_slots, SegIter seg_it,
                                               uint32_t tmem_base, int role) {
  const uint32_t idesc0 = umma_idesc_bf16(128, 0);
  const uint32_t idesc_step = (64u >> 3) << 17;
  const uint32_t w_lo = umma_desc_lo(w_smem);
  const uint32_t par_w_lo = w_lo + 3 * kDxb;
  mbar_wait(smem_u32(&misc->w_full), 0, 4);
  struct StepCtx {
    bool valid;
    int j, len, j_first;
    uint32_t ord0, sc, a_slot;
  };
  Segment seg = seg_it.valid() ? seg_it.get() : Segment{0, 0, 0, 0, 0, -1};
  Ring ar(ring_slots);
  StepCtx cur{seg_it.valid(), seg.j_first, seg.len, seg.j_first, 0u, 0u, ar.slot};
  auto advance = [&](StepCtx& c) {
    ar.advance();
    c.sc += 1;
    c.a_slot = ar.slot;
    if (c.j < seg.j_last) {
      c.j += 1;
      return;
    }
    const uint32_t next_ord0 = c.ord0 + (uint32_t)seg.len;
    seg_it.next();
    c.valid = seg_it.valid();
    if (c.valid) {
      seg = seg_it.get();
      c.j = seg.j_first;
      c.len = seg.len;
      c.j_first = seg.j_first;
      c.ord0 = next_ord0;
    }
  };
  const uint32_t go_step = smem_u32(&misc->go_step), go_par = smem_u32(&misc->go_par);
  bool pend = false;
  uint32_t pend_bar = 0;
  auto mma_range = [&](uint32_t slot_lo, int cnt, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
    const int n1 = min(cnt, kAccRing - (int)slot_lo);
    umma_bf16_lo(tmem_base + slot_lo * 64, a_lo, kDescHiSw128, b_lo, kDescHiSw128, idesc0 + n1 * idesc_step, acc);
    if (cnt > n1)
      umma_bf16_lo(tmem_base, a_lo, kDescHiSw128, b_lo + n1 * kSbb, kDescHiSw128, idesc0 + (cnt - n1) * idesc_step,
                   acc);
  };
  if (cur.valid) spin_until_ge(go_step, 1, 5);
  tc_fence_after();
  const uint32_t par_done_bar = smem_u32(&misc->par_done);
  while (cur.valid) {
    const bool tr = kTrace && blockIdx.x < 2 && cur.sc < 128;
    long long* trp = p.trace + (role * 128 + (int)cur.sc) * 8;
    if (tr) trp[0] = clock64();
    if (cur.j >= 1 && cur.j + 2 <= cur.len) {
      // interior step: see mma_fast_step
      const uint32_t s0 = (cur.ord0 + (uint32_t)(cur.j - 1)) % kAccRing;
      const uint32_t a_row = umma_desc_lo(a_smem + cur.a_slot * kASlotBytes);
      const uint32_t pb = pend ? pend_bar : 0u;
      const uint32_t par_target = kEpiWarps * (cur.ord0 + (uint32_t)cur.j);
      if (s0 + 3 <= kAccRing)
        mma_fast_step<kPar, 3>(tmem_base, s0 * 64, a_row, w_lo, go_step, cur.sc + 2, go_par, par_target, pb, par_done_bar);
      else if (s0 + 2 == kAccRing)
        mma_fast_step<kPar, 2>(tmem_base, s0 * 64, a_row, w_lo, go_step, cur.sc + 2, go_par, par_target, pb, par_done_bar);
      else
        mma_fast_step<kPar, 1>(tmem_base, s0 * 64, a_row, w_lo, go_step, cur.sc + 2, go_par, par_target, pb, par_done_bar);
      pend = true;
      pend_bar = smem_u32(&misc->step_done[cur.sc & (kStepRing - 1)]);
      if (tr) trp[2] = clock64();
      ar.advance();
      cur.sc += 1;
      cur.a_slot = ar.slot;
      cur.j += 1;
      continue;
    }
    const int lo = max(cur.j - 1, 0);
    const int hi = min(cur.j + 1, cur.len - 1);
    const int cnt = hi - lo + 1;
    const int new_from = (cur.j == cur.j_first) ? lo : cur.j + 1;   // rows first touched in this step
    const int old_cnt = min(max(new_from - lo, 0), cnt);
    const int new_cnt = cnt - old_cnt;
    const bool centre = (cur.j >= 0 && cur.j < cur.len);
    const uint32_t slot_lo = (cur.ord0 + lo) % kAccRing;
    const uint32_t a_row = umma_desc_lo(a_smem + cur.a_slot * kASlotBytes);
    const uint32_t b_row = w_lo + (uint32_t)(lo - (cur.j - 1)) * kSbb;   // first dy sub-block in range
    const uint32_t cur_sc = cur.sc;
    const uint32_t cur_od = cur.ord0 + (uint32_t)max(cur.j, 0);
    StepCtx nxt = cur;
    advance(nxt);
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      if (dx == 2) {
        if (nxt.valid) spin_until_ge(go_step, nxt.sc + 1, 5);   // normally long satisfied
        tc_fence_after();
        if (tr) trp[1] = clock64();
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_lo = a_row + dx * 8 + 2 * k;
        const uint32_t b_lo = b_row + dx * kDxb + 2 * k;
        if (dx == 0 && k == 0) {
          if (old_cnt > 0) mma_range(slot_lo, old_cnt, a_lo, b_lo, 1);
          if (new_cnt > 0) mma_range((slot_lo + old_cnt) % kAccRing, new_cnt, a_lo, b_lo + old_cnt * kSbb, 0);
          if (pend) {
            umma_commit(pend_bar);
            pend = false;
          }
        } else {
          mma_range(slot_lo, cnt, a_lo, b_lo, 1);
        }
      }
    }
    if (kPar && centre) {
      // partition 1x1 convs of this row of t: centre pixel column, N = 192, own TMEM region
      spin_until_ge(go_par, kEpiWarps * cur_od, 10);   // every epilogue warp has read row cur_od-1's region
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_lo(tmem_base + kParCol, a_row + 8 + 2 * k, kDescHiSw128, par_w_lo + 2 * k, kDescHiSw128,
                     idesc0 + 3 * idesc_step, k > 0);
      // own commit, not deferred: the epilogue's read of this region is on the only recurrence of the
      // pipeline (the next row's 1x1 MMAs wait for it)
      umma_commit(par_done_bar);
    }
    pend = true;
    pend_bar = smem_u32(&misc->step_done[cur_sc & (kStepRing - 1)]);
    if (tr) trp[2] = clock64();
    cur = nxt;
  }
  if (pend) umma_commit(pend_bar);
}

// ---------------------------------------------------------------------------------------- barrier scout
// Does every mbarrier wait the MMAs depend on and publishes plain counters (see pnp_conv_rows.cu).
// Stage 2 additionally acknowledges each landed row of t to the producer (its staging tile is free).
template <bool kPar, int kAccRing, bool kTrace>
__device__ __forceinline__ void scout_loop(const BlockParams& p, PairMisc* misc, uint32_t ring_slots, SegIter it,
                                           uint32_t remote_stage_free) {
  const uint32_t go_step = smem_u32(&misc->go_step);
  Ring ar(ring_slots);
  uint32_t sc = 0, ord0 = 0;
  for (; it.valid(); it.next()) {
    const Segment s = it.get();
    for (int j = s.j_first; j <= s.j_last; ++j, ++sc, ar.advance()) {
      const bool tr = kTrace && blockIdx.x < 2 && sc < 128;
      long long* trp = p.trace + ((kPar ? 0 : 1) * 128 + (int)sc) * 8;
      if (tr) trp[3] = clock64();
      mbar_wait(smem_u32(&misc->a_full[ar.slot]), ar.phase, 5);
      if (tr) trp[4] = clock64();
      if (!kPar) mbar_arrive_remote(remote_stage_free + (sc & 1) * 8);
      const int lo = max(j - 1, 0), hi = min(j + 1, s.len - 1);
      const int new_from = (j == s.j_first) ? lo : j + 1;
      for (int o = max(new_from, lo); o <= hi; ++o) {
        const uint32_t od = ord0 + o;
        mbar_wait(smem_u32(&misc->acc_free[od % kAccRing]), ((od / kAccRing) & 1) ^ 1, 6);
      }
      if (tr) trp[5] = clock64();
      st_release_shared(go_step, sc + 1);
    }
    ord0 += s.len;
  }
}

}  // namespace

template <bool kTrace>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBlockThreads, 1)
resblock_pair_kernel(const __grid_constant__ BlockParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - raw);
  const int role = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const PairLayout L0 = pair_layout(0, p.s_a, p.n_t);
  const PairLayout L1 = pair_layout(1, p.s_a, p.n_t);
  const PairLayout L = role == 0 ? L0 : L1;
  PairMisc* misc = reinterpret_cast<PairMisc*>(sgen + L.misc);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_begin = pair * p.tiles_per_pair;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_pair);
  const uint32_t ring_slots = role == 0 ? p.s_a : p.n_t;
  const int w_bytes = role == 0 ? kBlockW0Bytes : kBlockW1Bytes;

  if (threadIdx.x < 64) {
    const float* b = role == 0 ? p.bias0 : p.bias1;
    misc->bias[threadIdx.x] = b ? b[threadIdx.x] : 0.0f;
  }
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&misc->w_full), 1);
      for (int i = 0; i < kMaxSlots; ++i) {
        mbar_init(smem_u32(&misc->a_full[i]), 1);
        mbar_init(smem_u32(&misc->acc_free[i]), kEpiWarps);
        mbar_init(smem_u32(&misc->t_free[i]), 1);
      }
      for (int i = 0; i < kStepRing; ++i) mbar_init(smem_u32(&misc->step_done[i]), 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(smem_u32(&misc->stage_free[i]), 1);
        mbar_init(smem_u32(&misc->staged[i]), kEpiWarps);
      }
      mbar_init(smem_u32(&misc->par_done), 1);
      misc->go_step = 0;
      misc->go_par = 0;
      mbar_fence_init();
      tma_prefetch_desc(role == 0 ? &p.tm_src : &p.tm_out);
    }
    __syncwarp();
    tmem_alloc(smem_u32(&misc->tmem_base), kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the partner's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = misc->tmem_base;
  griddep_launch_dependents();
  griddep_wait();
  if (kTrace && threadIdx.x == 0) {
    if (blockIdx.x == 0) p.trace[6 * 128 * 8] = clock64();
    p.trace[6 * 128 * 8 + 8 + blockIdx.x * 2] = (long long)globaltimer_ns();
  }
  const uint32_t w_smem = sbase + L.w;
  const uint32_t ring_smem = sbase + L.ring;

  if (warp < 4) {
    setmaxnreg_dec<kCtrlRegs>();      // control warpgroup: hands registers to the epilogue warpgroups
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t wbar = smem_u32(&misc->w_full);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(role == 0 ? p.w0 : p.w1);
      mbar_arrive_expect_tx(wbar, w_bytes);
      for (int off = 0; off < w_bytes; off += kWChunkBytes) bulk_load_1d(w_smem + off, wsrc + off, kWChunkBytes, wbar);
      if (role == 0) {
        // ====================================================== stage 1: TMA producer of source rows
        Ring ar(ring_slots);
        uint32_t sc = 0;
        for (SegIter it(p, t_begin, t_end, true); it.valid(); it.next()) {
          const Segment s = it.get();
          const int x0 = s.strip * kBlockOutPx;
          for (int j = s.j_first; j <= s.j_last; ++j, ++sc, ar.advance()) {
            if (sc >= ring_slots) {          // slot last used by step sc - ring_slots
              const uint32_t ps = sc - ring_slots;
              mbar_wait(smem_u32(&misc->step_done[ps & (kStepRing - 1)]), (ps >> 3) & 1, 1);
            }
            const uint32_t fb = smem_u32(&misc->a_full[ar.slot]);
            if ((p.debug_skip & 1) && sc >= ring_slots) {
              mbar_arrive(fb);
            } else {
              mbar_arrive_expect_tx(fb, kRowBytes);
              tma_load_4d(ring_smem + ar.slot * kASlotBytes, &p.tm_src, fb, 0, x0 - 2, s.y_b + j, s.n);
            }
          }
        }
      } else {
        // ====================================================== stage 2: ring-slot release relay
        // a row of t may be overwritten once the MMAs of its step have completed
        const uint32_t r_t_free = mapa_shared(smem_u32(&misc->t_free[0]), 0);
        Ring tr(ring_slots);
        uint32_t sc = 0;
        for (SegIter it(p, t_begin, t_end, false); it.valid(); it.next()) {
          const Segment s = it.get();
          for (int j = s.j_first; j <= s.j_last; ++j, ++sc, tr.advance()) {
            mbar_wait(smem_u32(&misc->step_done[sc & (kStepRing - 1)]), (sc >> 3) & 1, 2);
            mbar_arrive_remote(r_t_free + tr.slot * 8);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      if (role == 0)
        mma_issue_loop<true, kRing0, kTrace>(p, misc, w_smem, ring_smem, ring_slots, SegIter(p, t_begin, t_end, true),
                                     tmem_base, 0);
      else
        mma_issue_loop<false, kRing1, kTrace>(p, misc, w_smem, ring_smem, ring_slots, SegIter(p, t_begin, t_end, false),
                                      tmem_base, 1);
    }
  } else if (warp == 2) {
    if (elect_one()) {
      if (role == 0)
        scout_loop<true, kRing0, kTrace>(p, misc, ring_slots, SegIter(p, t_begin, t_end, true), 0u);
      else
        scout_loop<false, kRing1, kTrace>(p, misc, ring_slots, SegIter(p, t_begin, t_end, false),
                                  mapa_shared(smem_u32(&misc->stage_free[0]), 0));
    }
  } else if (warp == 3) {
    // ============================================================ stage 1: DSMEM copy issuer
    // Pushes every staged row of t into the partner's ring.  A thread of its own: the remote
    // arrive.expect_tx + bulk copy took ~2000 cycles when the epilogue's store lane issued them.
    if (role == 1 && elect_one()) {
      // stage 2: TMA store issuer.  Stores the staged output rows and releases each staging tile as soon as
      // the store has read it.
      uint32_t k = 0;
      for (RowCur cur(p, t_begin, t_end, false); cur.valid; cur.next(), ++k) {
        const uint32_t b = k & 1;
        mbar_wait(smem_u32(&misc->staged[b]), (k >> 1) & 1, 14);
        if (!(p.debug_skip & 2)) {
          tma_store_4d(&p.tm_out, sbase + L1.stage + b * kTileBytes, 0, cur.s.strip * kBlockOutPx, cur.s.y_b + cur.o,
                       cur.s.n);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        mbar_arrive(smem_u32(&misc->stage_free[b]));
      }
      tma_store_wait_all<0>();
    }
    if (role == 0 && elect_one()) {
      const uint32_t r_ring = mapa_shared(sbase + L1.ring, 1);
      const uint32_t r_full = mapa_shared(smem_u32(&misc->a_full[0]), 1);
      Ring tr(p.n_t);
      uint32_t ord = 0;
      for (RowCur cur(p, t_begin, t_end, true); cur.valid; cur.next(), tr.advance(), ++ord) {
        const uint32_t b = ord & 1;
        mbar_wait(smem_u32(&misc->staged[b]), (ord >> 1) & 1, 14);
        mbar_wait(smem_u32(&misc->t_free[tr.slot]), tr.phase ^ 1, 13);
        mbar_arrive_expect_tx_remote(r_full + tr.slot * 8, kTileBytes);
        bulk_copy_to_cluster(r_ring + tr.slot * kASlotBytes, sbase + L0.stage + b * kTileBytes, kTileBytes,
                             r_full + tr.slot * 8);
      }
    }
  }
  } else {
    setmaxnreg_inc<kEpiRegs>();
    // ============================================================ epilogue: sixteen warps, thread = one pixel
    // (TMEM lane) x 16 channels.  Small on purpose: ncu showed the earlier 8-warp / 4-warp epilogues (32 / 64
    // channels per thread, ~800 straight-line instructions per row) spending 40 % of their samples on
    // instruction-fetch stalls -- the body did not fit the 6 KB L0 instruction cache of a sub-partition.
    // This one is ~150 instructions, all four warps of a sub-partition run the same code in phase, and the
    // per-row latency is short enough for ONE team to keep up with the MMA steps.
    const int q = warp & 3;
    const int cq = (warp - 4) >> 2;            // channel group: channels kEpiCh*cq .. kEpiCh*cq + kEpiCh-1
    const int m = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sw = (uint32_t)(m & 7);
    constexpr int kChunk = 16;                 // channels per partition-accumulator round trip
    constexpr int kChunks = kEpiCh / kChunk;
    float bias_r[kEpiCh];
#pragma unroll
    for (int j = 0; j < kEpiCh; ++j) bias_r[j] = misc->bias[cq * kEpiCh + j];
    if (role == 0) {
      // ---------------------------------------------------------- stage 1: t = relu(3x3 + bias + blend)
      // Iteration k: the step that completes the 3x3 result of row k also carries the 1x1 accumulators of row
      // k+1 (issued last in it, own commit), so one barrier wait covers both.  Order inside the iteration is
      // dictated by the only recurrence of the pipeline -- the 1x1 MMAs of row k+2 wait until row k+1's
      // partition region has been read: TMEM reads and the region hand-back (plain counter) come first, the
      // slow shared-memory work (staging stores, proxy fence, mbarrier arrivals) after.
      const uint32_t go_par = smem_u32(&misc->go_par);
      auto par_load = [&](const RowCur& c, float& q0, float& q1, float& q2) {
        const int px = c.s.strip * kBlockOutPx - 1 + m;
        q0 = q1 = q2 = 0.f;
        if (c.valid && px >= 0 && px < p.W) {
          const float* pp = p.par + (long long)c.s.n * p.par_sn + (long long)(c.s.y_b + c.o) * p.par_sy + px;
          q0 = __ldg(pp);
          q1 = __ldg(pp + p.par_sc);
          q2 = __ldg(pp + 2 * p.par_sc);
        }
      };
      float dy[kEpiCh];
      const uint32_t par_col = lane_base + kParCol + cq * kEpiCh;
      auto par_ld = [&](int c, float (&a1)[kChunk], float (&a2)[kChunk], float (&a3)[kChunk]) {
        tmem_ld16(par_col + c * kChunk, a1);
        tmem_ld16(par_col + 64 + c * kChunk, a2);
        tmem_ld16(par_col + 128 + c * kChunk, a3);
      };
      auto blend = [&](const float (&a1)[kChunk], const float (&a2)[kChunk], const float (&a3)[kChunk], float q0,
                       float q1, float q2, int c) {
#pragma unroll
        for (int j = 0; j < kChunk; ++j)
          dy[c * kChunk + j] = fmaf(q2, a3[j], fmaf(q1, a2[j], fmaf(q0, a1[j], bias_r[c * kChunk + j])));
      };
      // whole partition part of one row (first row of a segment: nothing else to read with it)
      auto par_row = [&](float q0, float q1, float q2) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          float a1[kChunk], a2[kChunk], a3[kChunk];
          par_ld(c, a1, a2, a3);
          tmem_ld_wait();
          if (c == kChunks - 1) {
            tc_fence_before();
            warp_flag_add(go_par);
          }
          blend(a1, a2, a3, q0, q1, q2, c);
        }
      };
      RowCur cur(p, t_begin, t_end, true);     // row k
      RowCur look = cur;                       // runs ahead: fetches partition values one row early
      float n0, n1, n2;
      {
        float q0, q1, q2;
        par_load(look, q0, q1, q2);            // row 0
        if (look.valid) look.next();
        par_load(look, n0, n1, n2);            // row 1
        if (cur.valid) {
          mbar_wait_warp(smem_u32(&misc->par_done), 0, 11);
          tc_fence_after();
          par_row(q0, q1, q2);
        }
      }
      uint32_t k = 0;
      uint32_t slot = 0;
      while (cur.valid) {
        const bool etr = kTrace && blockIdx.x < 2 && (warp == 4 || warp == 5) && lane == 0 && k < 128;
        long long* etp = p.trace + ((warp - 2) * 128 + (int)k) * 8;
        if (etr) etp[0] = clock64();
        const int px = cur.s.strip * kBlockOutPx - 1 + m;
        const bool in_img = (px >= 0) && (px < p.W);
        const bool has_next = look.valid;      // look is at row k+1
        const bool same_seg = has_next && (cur.o + 1 < cur.s.len);
        if (same_seg) {
          mbar_wait_warp(smem_u32(&misc->par_done), (k + 1) & 1, 11);
        } else {                               // segment end: the row's own last step; the next row's 1x1 come later
          const uint32_t scl = cur.sc_last();
          mbar_wait_warp(smem_u32(&misc->step_done[scl & (kStepRing - 1)]), (scl >> 3) & 1, 9);
        }
        tc_fence_after();
        if (etr) etp[1] = clock64();
        uint32_t w[kEpiCh / 2];
        const uint32_t acc_col = lane_base + slot * 64 + cq * kEpiCh;
        {
          float v[kEpiCh];
#pragma unroll
          for (int g = 0; g < kEpiCh / 16; ++g) tmem_ld16(acc_col + g * 16, v + g * 16);
          if (same_seg) {
            // 3x3 result of row k and the first chunk of row k+1's partition accumulators in one round trip
            float a1[kChunk], a2[kChunk], a3[kChunk];
            par_ld(0, a1, a2, a3);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < kEpiCh / 2; ++j) w[j] = pack_bf16x2_relu(v[2 * j] + dy[2 * j], v[2 * j + 1] + dy[2 * j + 1]);
            if (kChunks == 1) {
              tc_fence_before();
              warp_flag_add(go_par);
              if (etr) etp[3] = clock64();
            }
            blend(a1, a2, a3, n0, n1, n2, 0);  // dy of row k is dead from here on
#pragma unroll
            for (int c = 1; c < kChunks; ++c) {
              par_ld(c, a1, a2, a3);
              tmem_ld_wait();
              if (c == kChunks - 1) {          // region handed back a few TMEM round trips after it became readable
                tc_fence_before();
                warp_flag_add(go_par);
                if (etr) etp[3] = clock64();
              }
              blend(a1, a2, a3, n0, n1, n2, c);
            }
          } else {
            tmem_ld_wait();
            tc_fence_before();
#pragma unroll
            for (int j = 0; j < kEpiCh / 2; ++j) w[j] = pack_bf16x2_relu(v[2 * j] + dy[2 * j], v[2 * j + 1] + dy[2 * j + 1]);
          }
        }
        warp_arrive_relaxed(smem_u32(&misc->acc_free[slot]));
        if (!in_img) {                         // t outside the image is conv1's zero padding
#pragma unroll
          for (int j = 0; j < kEpiCh / 2; ++j) w[j] = 0u;
        }
        // the copy out of this staging tile (row k-2) landed long ago; the wait is a formality
        mbar_wait_warp(smem_u32(&misc->stage_free[k & 1]), ((k >> 1) & 1) ^ 1, 12);
        uint8_t* rowp = sgen + L0.stage + (k & 1) * kTileBytes + m * 128;
#pragma unroll
        for (int c = 0; c < kEpiCh / 8; ++c)   // 16-byte chunks of the pixel's 128-byte row
          *reinterpret_cast<uint4*>(rowp + (((cq * (kEpiCh / 8) + c) ^ sw) << 4)) =
              make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        fence_proxy_async_smem();
        warp_arrive(smem_u32(&misc->staged[k & 1]));       // warp 3 pushes the tile to the partner
        if (etr) etp[2] = clock64();
        if (has_next && !same_seg) {           // first row of a new segment: its 1x1 MMAs come with a later step
          mbar_wait_warp(smem_u32(&misc->par_done), (k + 1) & 1, 11);
          tc_fence_after();
          par_row(n0, n1, n2);
        }
        if (has_next) {
          look.next();
          par_load(look, n0, n1, n2);          // row k+2
        }
        cur.next();
        ++k;
        if (++slot == kRing0) slot = 0;
      }
    } else {
      // ---------------------------------------------------------- stage 2: out = x + 3x3(t) + bias
      const bool m_ok = m < kBlockOutPx;
      uint32_t k = 0, slot = 0;
      for (RowCur cur(p, t_begin, t_end, false); cur.valid; cur.next(), ++k) {
        const bool etr = kTrace && blockIdx.x < 2 && warp == 4 && lane == 0 && k < 128;
        long long* etp = p.trace + (4 * 128 + (int)k) * 8;
        if (etr) etp[0] = clock64();
        const Segment& s = cur.s;
        const int px = s.strip * kBlockOutPx + m;
        const int y = s.y_b + cur.o;
        const bool valid = m_ok && px < p.W;
        uint4 idv[kEpiCh / 8];
#pragma unroll
        for (int c = 0; c < kEpiCh / 8; ++c) idv[c] = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          const uint4* ip = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) +
                                                           (((long long)s.n * p.H + y) * p.W + px) * 128 +
                                                           cq * (kEpiCh * 2));
#pragma unroll
          for (int c = 0; c < kEpiCh / 8; ++c) idv[c] = ldg_nc_v4(ip + c);
        }
        const uint32_t scl = cur.sc_last();
        mbar_wait_warp(smem_u32(&misc->step_done[scl & (kStepRing - 1)]), (scl >> 3) & 1, 9);
        tc_fence_after();
        if (etr) etp[1] = clock64();
        float v[kEpiCh];
#pragma unroll
        for (int g = 0; g < kEpiCh / 16; ++g) tmem_ld16(lane_base + slot * 64 + cq * kEpiCh + g * 16, v + g * 16);
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive_relaxed(smem_u32(&misc->acc_free[slot]));
        uint32_t w[kEpiCh / 2];
#pragma unroll
        for (int c = 0; c < kEpiCh / 8; ++c) {
          const uint32_t iw[4] = {idv[c].x, idv[c].y, idv[c].z, idv[c].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = 8 * c + 2 * j;
            w[4 * c + j] = pack_bf16x2(v[e] + bias_r[e] + bf16_lo(iw[j]), v[e + 1] + bias_r[e + 1] + bf16_hi(iw[j]));
          }
        }
        // the TMA store out of this staging tile (row k-2) has been read; formality
        mbar_wait_warp(smem_u32(&misc->stage_free[k & 1]), ((k >> 1) & 1) ^ 1, 12);
        uint8_t* rowp = sgen + L1.stage + (k & 1) * kTileBytes + m * 128;
#pragma unroll
        for (int c = 0; c < kEpiCh / 8; ++c)
          *reinterpret_cast<uint4*>(rowp + (((cq * (kEpiCh / 8) + c) ^ sw) << 4)) =
              make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
        fence_proxy_async_smem();
        warp_arrive(smem_u32(&misc->staged[k & 1]));       // warp 3 issues the TMA store
        if (etr) etp[2] = clock64();
        if (++slot == kRing1) slot = 0;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kTrace && threadIdx.x == 0) {
    if (blockIdx.x == 0) p.trace[6 * 128 * 8 + 1] = clock64();
    p.trace[6 * 128 * 8 + 8 + blockIdx.x * 2 + 1] = (long long)globaltimer_ns();
  }
  cluster_sync_all();            // no remote arrive / bulk copy may target a CTA that has exited
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

size_t block_smem_bytes(const BlockParams& p) {
  const uint32_t a = pair_layout(0, p.s_a, p.n_t).total, b = pair_layout(1, p.s_a, p.n_t).total;
  return (a > b ? a : b) + 1024;
}

namespace {
template <bool kTrace>
cudaError_t launch_block_variant(const BlockParams& p, int pairs, cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(resblock_pair_kernel<kTrace>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kBlockThreads);
  cfg.dynamicSmemBytes = block_smem_bytes(p);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, resblock_pair_kernel<kTrace>, p);
}
}  // namespace

// the tracing variant is a separate kernel so that its extra code never sits in the instruction cache
// of production launches
cudaError_t launch_block(const BlockParams& p, int pairs, cudaStream_t stream) {
  return p.trace != nullptr ? launch_block_variant<true>(p, pairs, stream)
                            : launch_block_variant<false>(p, pairs, stream);
}

}  // namespace pnp
