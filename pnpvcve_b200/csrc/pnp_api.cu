// extern "C" surface of libpnpvcve.so (declared in include/pnp_vcve.h).
#include <cmath>
#include <cstddef>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pnp_vcve.h"
#include "pnp_block.cuh"
#include "pnp_conv.cuh"
#include "pnp_ops.cuh"

namespace {

thread_local char g_err[512] = "";
int g_base_off_mode = 0;

int fail(int code, const char* fmt, const char* detail = "") {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return PNP_ERR_CUDA;
}

struct DeviceInfo {
  bool ok = false;
  int sms = 0;
};

// One entry per device ordinal; queried lazily.  Read-only after first use.
DeviceInfo g_dev[64];
bool g_dev_known[64] = {false};

int device_info(DeviceInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(PNP_ERR_ARG, "device ordinal out of range");
  if (!g_dev_known[dev]) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    g_dev[dev].ok = (prop.major == 10);
    g_dev[dev].sms = prop.multiProcessorCount;
    g_dev_known[dev] = true;
  }
  *out = &g_dev[dev];
  if (!g_dev[dev].ok) return fail(PNP_ERR_ARCH, "libpnpvcve needs an sm_100 device (B200)");
  return PNP_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (64, W, H, N) bf16 NHWC tensor, box (64, box_w, 1, 1), 128-byte swizzle, zero fill out of range.
// spx/sy/sn: element strides between pixels / rows / images (0 = contiguous NHWC).
int make_map(CUtensorMap* m, const void* base, int N, int H, int W, int box_w, long long spx = 0, long long sy = 0,
             long long sn = 0) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(PNP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  if (spx != 0) {
    strides[0] = (cuuint64_t)spx * 2;
    strides[1] = (cuuint64_t)sy * 2;
    strides[2] = (cuuint64_t)sn * 2;
  }
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PNP_ERR_CUDA;
  }
  return PNP_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

static_assert(sizeof(pnp_conv_desc) == 240 && offsetof(pnp_conv_desc, out_spx) == 200 &&
                  offsetof(pnp_conv_desc, wpack_stable) == 232,
              "pnp_conv_desc layout is part of the ABI (mirrored by pnpvcve_b200/_lib.py: ConvDesc)");
int pnp_abi_version(void) { return 6; }

const char* pnp_last_error(void) { return g_err; }

int pnp_device_check(void) {
  DeviceInfo* d;
  return device_info(&d);
}

int pnp_set_base_offset_mode(int mode) {
  if (mode != 0 && mode != 1) return fail(PNP_ERR_ARG, "base offset mode must be 0 or 1");
  g_base_off_mode = mode;
  return PNP_OK;
}

int pnp_mv_warp(const void* src, const float* flow_x, const float* flow_y, int64_t flow_row_stride,
                int64_t flow_image_stride, void* dst, int N, int H, int W, int32_t* dbg_x0, int32_t* dbg_y0,
                void* stream) {
  if (!src || !flow_x || !flow_y || !dst) return fail(PNP_ERR_ARG, "pnp_mv_warp: null pointer");
  if (N <= 0 || H <= 0 || W <= 0) return fail(PNP_ERR_ARG, "pnp_mv_warp: bad shape");
  if (!aligned16(src) || !aligned16(dst) || src == dst)
    return fail(PNP_ERR_ARG, "pnp_mv_warp: src/dst must be distinct 16-byte aligned buffers");
  if ((dbg_x0 == nullptr) != (dbg_y0 == nullptr) || (dbg_x0 && N != 1))
    return fail(PNP_ERR_ARG, "pnp_mv_warp: dbg outputs come in pairs and need N == 1");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_mv_warp(src, flow_x, flow_y, flow_row_stride, flow_image_stride, dst, N, H, W, dbg_x0,
                                      dbg_y0, d->sms, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_mv_warp");
}

int pnp_lr_im2col(const float* lr, int64_t sn, int64_t sc, int64_t sy, void* dst, int N, int H, int W,
                  void* stream) {
  if (!lr || !dst) return fail(PNP_ERR_ARG, "pnp_lr_im2col: null pointer");
  if (N <= 0 || H <= 0 || W <= 0 || !aligned16(dst)) return fail(PNP_ERR_ARG, "pnp_lr_im2col: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_lr_im2col(lr, sn, sc, sy, dst, N, H, W, d->sms, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_lr_im2col");
}

int pnp_pack_conv3x3(const float* w, int n_experts, const float* coef, const float* row_scale, int out_ch,
                     int in_total, int in_begin, int in_begin2, int in_count, void* dst, int center_chunks,
                     void* stream) {
  if (!w || !dst) return fail(PNP_ERR_ARG, "pnp_pack_conv3x3: null pointer");
  if (n_experts < 1 || (coef == nullptr && n_experts != 1) || out_ch < 1 || out_ch > 64 || in_count < 1 ||
      in_count > 64 || in_begin < 0 || in_begin + in_count > in_total ||
      (in_begin2 >= 0 && in_begin2 + in_count > in_total) || (center_chunks != 1 && center_chunks != 4) ||
      !aligned16(dst))
    return fail(PNP_ERR_ARG, "pnp_pack_conv3x3: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_pack_conv3x3(w, n_experts, coef, row_scale, out_ch, in_total, in_begin, in_begin2, in_count,
                                           dst, center_chunks, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_pack_conv3x3");
}

int pnp_pack_conv3x3_rowstack(const float* w, int n_experts, const float* coef, const float* row_scale,
                              int out_ch, int in_total, int in_begin, int in_begin2, int in_count, void* dst,
                              int tap_n, int flip_ky, void* stream) {
  if (!w || !dst) return fail(PNP_ERR_ARG, "pnp_pack_conv3x3_rowstack: null pointer");
  if (n_experts < 1 || (coef == nullptr && n_experts != 1) || out_ch < 1 || (tap_n != 64 && tap_n != 16) ||
      out_ch > tap_n || in_count < 1 || in_count > 64 || in_begin < 0 || in_begin + in_count > in_total ||
      (in_begin2 >= 0 && in_begin2 + in_count > in_total) || !aligned16(dst))
    return fail(PNP_ERR_ARG, "pnp_pack_conv3x3_rowstack: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_pack_conv3x3_rowstack(w, n_experts, coef, row_scale, out_ch, in_total, in_begin, in_begin2,
                                                    in_count, dst, tap_n, flip_ky != 0, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_pack_conv3x3_rowstack");
}

int pnp_pack_rows(const float* w, int rows, int cols, int64_t row_stride, int64_t col_stride, void* dst,
                  int row_offset, void* stream) {
  if (!w || !dst) return fail(PNP_ERR_ARG, "pnp_pack_rows: null pointer");
  if (rows < 1 || rows > 64 || cols < 1 || cols > 64 || row_offset < 0 || !aligned16(dst))
    return fail(PNP_ERR_ARG, "pnp_pack_rows: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_pack_rows(w, rows, cols, row_stride, col_stride, dst, row_offset,
                                        static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_pack_rows");
}

int pnp_pack_aux(const float* w, int out_ch, int in_total, void* dst, void* stream) {
  if (!w || !dst) return fail(PNP_ERR_ARG, "pnp_pack_aux: null pointer");
  if (out_ch < 1 || out_ch > 64 || in_total < 3 || !aligned16(dst))
    return fail(PNP_ERR_ARG, "pnp_pack_aux: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_pack_aux(w, out_ch, in_total, dst, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_pack_aux");
}

int pnp_caa_heads(const float* base_qp, const float* qp, int frames, const float* base0_w,
                  const float* base0_b, const float* base2_w, const float* base2_b, const float* se0_w,
                  const float* se2_w, int n_experts, int se_hidden, float* experts, float* gamma,
                  void* stream) {
  if (!base_qp || !qp || !base0_w || !base0_b || !base2_w || !base2_b || !se0_w || !se2_w || !experts || !gamma)
    return fail(PNP_ERR_ARG, "pnp_caa_heads: null pointer");
  if (frames < 1 || n_experts < 1 || n_experts > 16 || se_hidden < 1 || se_hidden > 64)
    return fail(PNP_ERR_ARG, "pnp_caa_heads: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_caa_heads(base_qp, qp, frames, base0_w, base0_b, base2_w, base2_b, se0_w, se2_w,
                                        n_experts, se_hidden, experts, gamma, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_caa_heads");
}

int pnp_mix_bias(const float* conv2_bias, int n_blocks, int n_experts, const float* experts,
                 const float* gamma, int frames, float* out, void* stream) {
  if (!conv2_bias || !experts || !gamma || !out) return fail(PNP_ERR_ARG, "pnp_mix_bias: null pointer");
  if (n_blocks < 1 || n_blocks > 65535 || n_experts < 1 || frames < 1)
    return fail(PNP_ERR_ARG, "pnp_mix_bias: bad argument");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_mix_bias(conv2_bias, (long long)n_experts * 64, n_blocks, n_experts, experts, gamma,
                                       frames, out, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_mix_bias");
}

int pnp_mv_rasterize(const float* records, const int32_t* frame_offsets, const int32_t* is_b,
                     const int32_t* p_target, int T, int R, int H, int W, uint32_t* owner_fwd,
                     uint32_t* owner_bwd, uint32_t* part_mask, float* mvs, float* partitions,
                     int32_t* status, void* stream) {
  if (!frame_offsets || !is_b || !p_target || !owner_fwd || !owner_bwd || !part_mask || !mvs || !partitions ||
      !status || (R > 0 && !records))
    return fail(PNP_ERR_ARG, "pnp_mv_rasterize: null pointer");
  if (T < 1 || T > 65535 || R < 0 || H < 1 || H > 65535 || W < 1 || (long long)T * H * W >= (1LL << 31))
    return fail(PNP_ERR_ARG, "pnp_mv_rasterize: bad shape");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  cudaError_t e = pnp::launch_mv_rasterize(records, frame_offsets, is_b, p_target, T, R, H, W, owner_fwd, owner_bwd,
                                           part_mask, mvs, partitions, status, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_mv_rasterize");
}

int pnp_frame_quality(const float* a, int64_t a_sf, int64_t a_sc, int64_t a_sy, const float* b, int64_t b_sf,
                      int64_t b_sc, int64_t b_sy, int F, int H, int W, int crop_border, unsigned long long* sse,
                      double* ssim_sum, void* stream) {
  if (!a || !b || !sse || !ssim_sum) return fail(PNP_ERR_ARG, "pnp_frame_quality: null pointer");
  if (F < 1 || crop_border < 0 || H - 2 * crop_border < 11 || W - 2 * crop_border < 11)
    return fail(PNP_ERR_ARG, "pnp_frame_quality: the cropped frame must be at least 11 x 11");
  if ((long long)F * 3 > 65535) return fail(PNP_ERR_ARG, "pnp_frame_quality: at most 21845 frames per call");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  // cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)), normalised to sum 1 (float64)
  double g[11], s = 0.0;
  for (int i = 0; i < 11; ++i) {
    g[i] = exp(-((double)(i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    s += g[i];
  }
  for (int i = 0; i < 11; ++i) g[i] /= s;
  const int ch_first = crop_border != 0 ? 2 : 0, n_ch = crop_border != 0 ? 1 : 3;
  cudaError_t e = pnp::launch_frame_quality(a, a_sf, a_sc, a_sy, b, b_sf, b_sc, b_sy, F, H, W, crop_border, ch_first,
                                            n_ch, g, sse, ssim_sum, d->sms, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_frame_quality");
}

int pnp_conv3x3(const pnp_conv_desc* c, void* stream) {
  if (!c) return fail(PNP_ERR_ARG, "pnp_conv3x3: null descriptor");
  if (!c->src || !c->wpack) return fail(PNP_ERR_ARG, "pnp_conv3x3: null src/wpack");
  if (c->N < 1 || c->H < 1 || c->W < 1) return fail(PNP_ERR_ARG, "pnp_conv3x3: bad shape");
  const bool last = (c->mode == PNP_CONV_LAST);
  if (c->mode != PNP_CONV_BF16 && !last) return fail(PNP_ERR_ARG, "pnp_conv3x3: bad mode");
  if (last) {
    if (!c->lq || !c->outf || c->center_n != 16 || c->tap_n != 16 || c->aux || c->idt || c->par)
      return fail(PNP_ERR_ARG, "pnp_conv3x3: PNP_CONV_LAST needs lq/outf, N=16, no aux/idt/par");
  } else {
    if (!c->out || c->tap_n != 64 || (c->center_n != 64 && c->center_n != 256))
      return fail(PNP_ERR_ARG, "pnp_conv3x3: PNP_CONV_BF16 needs out, tap_n=64, center_n in {64,256}");
    if ((c->par != nullptr) != (c->center_n == 256))
      return fail(PNP_ERR_ARG, "pnp_conv3x3: par requires center_n == 256 and vice versa");
    if (c->out == c->src) return fail(PNP_ERR_ARG, "pnp_conv3x3: out must not alias src (halo rows)");
  }
  if ((c->aux != nullptr) != (c->aux_k16 > 0) || c->aux_k16 < 0 || c->aux_k16 > 4)
    return fail(PNP_ERR_ARG, "pnp_conv3x3: aux / aux_k16 mismatch");
  if (c->aux && c->center_n == 256) return fail(PNP_ERR_ARG, "pnp_conv3x3: aux and par are exclusive");
  const bool rowstack = (c->wlayout == PNP_WLAYOUT_ROWSTACK);
  if (c->wlayout != PNP_WLAYOUT_TAPMAJOR && !rowstack) return fail(PNP_ERR_ARG, "pnp_conv3x3: bad wlayout");
  if (rowstack && c->center_n == 256 && (c->aux || c->idt))
    return fail(PNP_ERR_ARG, "pnp_conv3x3: row-stacked layout with partition convs takes no aux / idt");
  const int center_chunks = (c->center_n == 256) ? 4 : 1;
  const int need_chunks = center_chunks + 8 + (c->aux ? 1 : 0);
  if (!rowstack && c->n_wchunks != need_chunks)
    return fail(PNP_ERR_ARG, "pnp_conv3x3: n_wchunks does not match the layout");
  if (c->act < 0 || c->act > 2) return fail(PNP_ERR_ARG, "pnp_conv3x3: bad act");
  const bool strided_out = c->out_spx != 0 || c->out_sy != 0 || c->out_sn != 0;
  if (strided_out && (last || c->out_spx < 64 || c->out_sy < c->out_spx * c->W || (c->N > 1 && c->out_sn < c->out_sy * c->H) ||
                      ((c->out_spx | c->out_sy | c->out_sn) & 7)))
    return fail(PNP_ERR_ARG, "pnp_conv3x3: out strides must be non-overlapping multiples of 8 elements (PNP_CONV_BF16 only)");
  if (c->lq_up4 && (!last || !rowstack || (c->H & 3) || (c->W & 3)))
    return fail(PNP_ERR_ARG, "pnp_conv3x3: lq_up4 needs PNP_CONV_LAST, the row-stacked layout and H, W multiples of 4");
  if (!aligned16(c->src) || !aligned16(c->wpack) || (c->aux && !aligned16(c->aux)) ||
      (c->idt && !aligned16(c->idt)) || (c->out && !aligned16(c->out)))
    return fail(PNP_ERR_ARG, "pnp_conv3x3: pointers must be 16-byte aligned");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;

  pnp::ConvParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_map(&p.tm_src, c->src, c->N, c->H, c->W, pnp::kHaloPx))) return rc;
  if (c->aux && (rc = make_map(&p.tm_aux, c->aux, c->N, c->H, c->W, pnp::kTilePx))) return rc;
  if (c->idt && (rc = make_map(&p.tm_id, c->idt, c->N, c->H, c->W, pnp::kTilePx))) return rc;
  if (!last && (rc = make_map(&p.tm_out, c->out, c->N, c->H, c->W, pnp::kTilePx, c->out_spx, c->out_sy, c->out_sn)))
    return rc;
  if (last) p.tm_out = p.tm_src;  // never used; keeps the prefetch harmless
  p.wpack = c->wpack;
  p.scale = c->scale;
  p.bias = c->bias;
  p.par = c->par;
  p.par_sn = c->par_sn; p.par_sc = c->par_sc; p.par_sy = c->par_sy;
  p.lq = c->lq;
  p.lq_sn = c->lq_sn; p.lq_sc = c->lq_sc; p.lq_sy = c->lq_sy;
  p.outf = c->outf;
  p.of_sn = c->of_sn; p.of_sc = c->of_sc; p.of_sy = c->of_sy;
  p.H = c->H; p.W = c->W; p.N = c->N;
  p.strips = (c->W + pnp::kTilePx - 1) / pnp::kTilePx;
  const long long tiles = (long long)c->N * p.strips * c->H;
  if (tiles > 0x7fffffffLL) return fail(PNP_ERR_ARG, "pnp_conv3x3: too many tiles");
  p.tiles_total = (int)tiles;
  int grid = d->sms < p.tiles_total ? d->sms : p.tiles_total;
  p.tiles_per_cta = (p.tiles_total + grid - 1) / grid;
  grid = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.n_wchunks = c->n_wchunks;
  p.center_n = c->center_n;
  p.tap_n = c->tap_n;
  p.aux_k16 = c->aux_k16;
  p.has_id = c->idt != nullptr;
  p.act = c->act;
  p.mode = last ? pnp::kModeLast : pnp::kModeBf16;
  p.flip_y = (rowstack && c->flip_y) ? 1 : 0;
  p.w_stable = c->wpack_stable ? 1 : 0;
  p.lq_up4 = c->lq_up4 ? 1 : 0;
  p.par_sparse = (c->par && c->par_sparse) ? 1 : 0;
  {
    // diagnostic: block launch B (row-stacked, identity, bottom-up) reads t and x for the last time
    const char* hp = getenv("PNP_L2_HINTS");
    p.l2_dead_reads = (hp && atoi(hp) != 0 && rowstack && c->idt && c->flip_y) ? 1 : 0;
  }
  {
    const char* sp = getenv("PNP_PAR_SPLIT");     // diagnostic switch for the row-stacked partition variant (default on)
    p.par_split = (rowstack && c->par && !(sp && atoi(sp) == 0)) ? 1 : 0;
  }
  p.base_off_mode = g_base_off_mode;
  {
    const char* dbg = getenv("PNP_DEBUG_SKIP");   // what-if profiling only; results are wrong when set
    p.debug_skip = dbg ? atoi(dbg) : 0;
    const char* trc = getenv("PNP_TRACE_PTR");    // device pointer (decimal) of a >= 4 KB buffer
    p.trace = trc ? reinterpret_cast<long long*>(strtoull(trc, nullptr, 10)) : nullptr;
  }
  // shared-memory budget: weights + (aux ring) + staging ring + source-row ring from what is left.
  // With an identity operand the staging ring also prefetches identity tiles (n_io - 2 tiles ahead),
  // so it gets 4 slots as long as 5 source rows (3 in use + 2 in flight) still fit.
  const long long budget = 232448 - 2048;
  const long long w_bytes = rowstack ? (((long long)9 * c->tap_n * 128 + (c->aux ? pnp::kWChunkBytes : 0) +
                                         (c->center_n == 256 ? 3 * 64 * 128 : 0) + 1023) & ~1023LL)
                                     : (long long)p.n_wchunks * pnp::kWChunkBytes;
  auto fixed_bytes = [&](int n_io) {
    return w_bytes + (c->aux ? 2 * pnp::kTileBytes : 0) + (long long)n_io * pnp::kTileBytes;
  };
  p.n_io = 2;
  if (rowstack && c->par) p.n_io = 3;   // the next row's 1x1 blend is parked in its staging slot one row early
  if (c->idt) {
    p.n_io = 4;
    while (p.n_io > 2 && (budget - fixed_bytes(p.n_io)) / pnp::kASlotBytes < 5) --p.n_io;
  }
  long long slots = (budget - fixed_bytes(p.n_io)) / pnp::kASlotBytes;
  const long long max_slots = rowstack ? 6 : pnp::kMaxASlots;   // a row lives one step there: 6 = 5 in flight
  if (slots > max_slots) slots = max_slots;
  if (slots < 4) return fail(PNP_ERR_RESOURCE, "pnp_conv3x3: shared-memory budget cannot hold 4 source rows");
  // the row-stacked MMA thread checks the next step's barriers before the current step is committed;
  // with an identity operand that needs a staging ring of >= 3 slots to stay deadlock free
  if (rowstack && c->idt && p.n_io < 3)
    return fail(PNP_ERR_RESOURCE, "pnp_conv3x3: row-stacked layout with idt needs 3 staging slots (drop aux)");
  p.s_a = (int)slots;
  if (const char* ov = getenv("PNP_RINGS")) {     // diagnostic: "<n_io>,<s_a>" override of the shared-memory split
    int nio = 0, sa = 0;
    if (sscanf(ov, "%d,%d", &nio, &sa) == 2 && nio >= 2 && nio <= pnp::kMaxIoSlots && sa >= 4 && sa <= pnp::kMaxASlots &&
        fixed_bytes(nio) + (long long)sa * pnp::kASlotBytes <= budget && !(rowstack && c->idt && nio < 3) &&
        !(rowstack && c->par && nio != 3)) {
      p.n_io = nio;
      p.s_a = sa;
    }
  }
  cudaError_t e = rowstack ? pnp::launch_conv_rows(p, grid, static_cast<cudaStream_t>(stream))
                           : pnp::launch_conv(p, grid, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_conv3x3");
}

int pnp_resblock(const pnp_block_desc* c, void* stream) {
  if (!c) return fail(PNP_ERR_ARG, "pnp_resblock: null descriptor");
  if (!c->x || !c->out || !c->w_stage1 || !c->w_stage2 || !c->par)
    return fail(PNP_ERR_ARG, "pnp_resblock: null x/out/weights/par");
  if (c->N < 1 || c->H < 1 || c->W < 1) return fail(PNP_ERR_ARG, "pnp_resblock: bad shape");
  if (c->out == c->x) return fail(PNP_ERR_ARG, "pnp_resblock: out must not alias x (halo rows, identity)");
  if (!aligned16(c->x) || !aligned16(c->out) || !aligned16(c->w_stage1) || !aligned16(c->w_stage2))
    return fail(PNP_ERR_ARG, "pnp_resblock: pointers must be 16-byte aligned");
  DeviceInfo* d;
  int rc = device_info(&d);
  if (rc) return rc;
  if (d->sms < 2) return fail(PNP_ERR_RESOURCE, "pnp_resblock: needs at least one SM pair");

  pnp::BlockParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_map(&p.tm_src, c->x, c->N, c->H, c->W, pnp::kHaloPx))) return rc;
  if ((rc = make_map(&p.tm_out, c->out, c->N, c->H, c->W, pnp::kBlockOutPx))) return rc;
  p.w0 = c->w_stage1;
  p.w1 = c->w_stage2;
  p.bias0 = c->bias1;
  p.bias1 = c->bias2;
  p.par = c->par;
  p.par_sn = c->par_sn; p.par_sc = c->par_sc; p.par_sy = c->par_sy;
  p.x = c->x;
  p.H = c->H; p.W = c->W; p.N = c->N;
  p.strips = (c->W + pnp::kBlockOutPx - 1) / pnp::kBlockOutPx;
  const long long tiles = (long long)c->N * p.strips * c->H;
  if (tiles > 0x7fffffffLL) return fail(PNP_ERR_ARG, "pnp_resblock: too many tiles");
  p.tiles_total = (int)tiles;
  int pairs = d->sms / 2 < p.tiles_total ? d->sms / 2 : p.tiles_total;
  p.tiles_per_pair = (p.tiles_total + pairs - 1) / pairs;
  pairs = (p.tiles_total + p.tiles_per_pair - 1) / p.tiles_per_pair;
  p.s_a = 5;
  p.n_t = 5;
  {
    const char* dbg = getenv("PNP_DEBUG_SKIP");   // what-if profiling only; results are wrong when set
    p.debug_skip = dbg ? atoi(dbg) : 0;
    const char* trc = getenv("PNP_TRACE_PTR");    // device pointer (decimal) of a >= 8 KB buffer
    p.trace = trc ? reinterpret_cast<long long*>(strtoull(trc, nullptr, 10)) : nullptr;
  }
  cudaError_t e = pnp::launch_block(p, pairs, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? PNP_OK : cuda_fail(e, "pnp_resblock");
}

}  // extern "C"
