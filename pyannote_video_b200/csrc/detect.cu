// detect.cu — input and output stages of the CNN (MMOD) face detector around the conv stack:
//   resize_bilinear  dlib pyramid_up / pyramid_down<6> (resize_image + interpolate_bilinear) into the
//                    tiled pyramid plane (RGBA u8, A = 255 inside tiles)
//   det_candidates   score map -> compacted list of cells above adjust_threshold
//   det_nms          loss_mmod::to_label: cells -> boxes in image space, sort by score, greedy NMS
// Reference call site: face_detector_(rgb, 1), pyannote/video/face/face.py:66.
// All float arithmetic uses explicitly rounded, unfused operations in the same order as
// oracle/pyramid.py so results are bit-identical.
#include <atomic>
#include <cstdlib>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

// u8 -> f32 and floor without the conversion unit (I2F / F2I / FRND run at a quarter of the FP32 rate):
// (2^23 | b) - 2^23 is exact; floor(t) for 0 <= t < 2^22 is the mantissa of t (+) 2^23 rounded down
__device__ __forceinline__ float pv_u8f(uint32_t b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }
// floor of 0 <= t < 2^22, as float bits with the integer in the mantissa
__device__ __forceinline__ uint32_t pv_floor_bits(float t) { return __float_as_uint(__fadd_rd(t, 8388608.0f)); }

// one output pixel of dlib's resize_image/interpolate_bilinear (explicitly rounded, unfused float32)
template <int SRC_CH>
__device__ __forceinline__ uchar4 bilinear_px(const uint8_t* __restrict__ s, int src_pitch_px, int sx0, int sy0, int sw,
                                              int sh, int r, int c, float xs, float ys) {
  uchar4 o;
  o.w = 255;
  const float y = __fmul_rn((float)r, ys);
  const float x = __fmul_rn((float)c, xs);
  int top = (int)(pv_floor_bits(y) & 0x7FFFFFu), left = (int)(pv_floor_bits(x) & 0x7FFFFFu);   // y, x >= 0
  top = min(top, sh - 1);
  left = min(left, sw - 1);
  const int bot = min(top + 1, sh - 1), right = min(left + 1, sw - 1);
  const float tb = __fsub_rn(y, (float)top), lr = __fsub_rn(x, (float)left);
  const float omlr = __fsub_rn(1.0f, lr), omtb = __fsub_rn(1.0f, tb);
  uint32_t tl[3], tr[3], bl[3], br[3];
  if (SRC_CH == 4) {   // one 4-byte load per tap
    const uchar4* s4 = reinterpret_cast<const uchar4*>(s);
    const uchar4 a = s4[(long long)(sy0 + top) * src_pitch_px + sx0 + left];
    const uchar4 b = s4[(long long)(sy0 + top) * src_pitch_px + sx0 + right];
    const uchar4 cc = s4[(long long)(sy0 + bot) * src_pitch_px + sx0 + left];
    const uchar4 d = s4[(long long)(sy0 + bot) * src_pitch_px + sx0 + right];
    tl[0] = a.x; tl[1] = a.y; tl[2] = a.z;
    tr[0] = b.x; tr[1] = b.y; tr[2] = b.z;
    bl[0] = cc.x; bl[1] = cc.y; bl[2] = cc.z;
    br[0] = d.x; br[1] = d.y; br[2] = d.z;
  } else {
    const uint8_t* ptl = s + ((long long)(sy0 + top) * src_pitch_px + sx0 + left) * SRC_CH;
    const uint8_t* ptr_ = s + ((long long)(sy0 + top) * src_pitch_px + sx0 + right) * SRC_CH;
    const uint8_t* pbl = s + ((long long)(sy0 + bot) * src_pitch_px + sx0 + left) * SRC_CH;
    const uint8_t* pbr = s + ((long long)(sy0 + bot) * src_pitch_px + sx0 + right) * SRC_CH;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { tl[ch] = ptl[ch]; tr[ch] = ptr_[ch]; bl[ch] = pbl[ch]; br[ch] = pbr[ch]; }
  }
  uint8_t res[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float a = __fadd_rn(__fmul_rn(omlr, pv_u8f(tl[ch])), __fmul_rn(lr, pv_u8f(tr[ch])));
    const float b = __fadd_rn(__fmul_rn(omlr, pv_u8f(bl[ch])), __fmul_rn(lr, pv_u8f(br[ch])));
    const float v = __fadd_rn(__fmul_rn(omtb, a), __fmul_rn(tb, b));
    // floor(v + 0.5) clamped to [0, 255]: v >= 0, so only the upper clamp can bind
    res[ch] = (uint8_t)(pv_floor_bits(fminf(__fadd_rn(v, 0.5f), 255.5f)) & 0xFFu);
  }
  o.x = res[0];
  o.y = res[1];
  o.z = res[2];
  return o;
}

// ---- column-walking resize (the big pyramid levels) -------------------------------------------------
// Same arithmetic as bilinear_px, bit for bit, organised so that the conversion unit (I2F / F2I / FRND run at
// a quarter of the FP32 rate and bounded the per-pixel kernel: ncu, profiles/README.md) is never used and the
// horizontal interpolation of a source row is shared by the output rows that read it:
//   * u8 -> f32 as (2^23 | b) - 2^23, floor(t) for 0 <= t < 2^22 as the mantissa of t (+) 2^23 rounded down;
//   * a thread owns ONE output column (left/right/lr are constants) and walks kResizeRows output rows; the
//     horizontally interpolated source rows h[top], h[bot] (3 floats each) are kept between rows.
constexpr int kResizeRows = 16;
constexpr int kResizeCols = 2;      // columns per thread, 256 apart (coalesced rows, shared row arithmetic)

// byte `SEL` (0..2) of v as an exact float: prmt builds 0x4B0000vv = 2^23 + vv, minus 2^23
template <int SEL>
__device__ __forceinline__ float pv_u8f_sel(uint32_t v, uint32_t magic) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(magic), "n"(0x7540 | SEL));
  return __fsub_rn(__uint_as_float(r), 8388608.0f);
}

// horizontally interpolated source row: h[ch] = omlr * S[il][ch] + lr * S[ir][ch]  (il, ir: pixel indices in the image)
template <int SRC_CH>
__device__ __forceinline__ void hrow(const uint8_t* __restrict__ s, uint32_t il, uint32_t ir, float lr, float omlr,
                                     uint32_t magic, float (&h)[3]) {
  if (SRC_CH == 4) {
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s);
    const uint32_t a = __ldg(s4 + il);
    const uint32_t b = __ldg(s4 + ir);
    h[0] = __fadd_rn(__fmul_rn(omlr, pv_u8f_sel<0>(a, magic)), __fmul_rn(lr, pv_u8f_sel<0>(b, magic)));
    h[1] = __fadd_rn(__fmul_rn(omlr, pv_u8f_sel<1>(a, magic)), __fmul_rn(lr, pv_u8f_sel<1>(b, magic)));
    h[2] = __fadd_rn(__fmul_rn(omlr, pv_u8f_sel<2>(a, magic)), __fmul_rn(lr, pv_u8f_sel<2>(b, magic)));
  } else {
    const uint8_t* pl = s + il * 3u;
    const uint8_t* pr = s + ir * 3u;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      h[ch] = __fadd_rn(__fmul_rn(omlr, pv_u8f((uint32_t)__ldg(pl + ch))), __fmul_rn(lr, pv_u8f((uint32_t)__ldg(pr + ch))));
  }
}

template <int SRC_CH>
__global__ void __launch_bounds__(256) resize_cols_kernel(const uint8_t* __restrict__ src, long long src_img_stride,
                                                          int src_pitch_px, int sx0, int sy0, int sw, int sh,
                                                          uchar4* __restrict__ dst, long long dst_img_stride, int dst_pitch_px,
                                                          int dx0, int dy0, int dw, int dh, float xs, float ys, uint32_t magic) {
  // magic = 0x4B000000 arrives as a kernel argument so that prmt keeps it as a constant-bank operand and the
  // selector as its one immediate
  const int c0 = blockIdx.x * (256 * kResizeCols) + threadIdx.x;
  if (c0 >= dw) return;
  const int r0 = blockIdx.y * kResizeRows;
  const int r1 = min(r0 + kResizeRows, dh);
  const uint8_t* s = src + (long long)blockIdx.z * src_img_stride;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst) + (long long)blockIdx.z * dst_img_stride + (long long)(dy0 + r0) * dst_pitch_px + dx0 + c0;
  uint32_t il[kResizeCols], ir[kResizeCols];      // pixel index of the left/right tap in source row sy0
  float lr[kResizeCols], omlr[kResizeCols];
  bool on[kResizeCols];
#pragma unroll
  for (int j = 0; j < kResizeCols; ++j) {
    const int c = c0 + 256 * j;
    on[j] = c < dw;
    const float x = __fmul_rn((float)min(c, dw - 1), xs);
    const int left = min((int)(pv_floor_bits(x) & 0x7FFFFFu), sw - 1);
    const int right = min(left + 1, sw - 1);
    lr[j] = __fsub_rn(x, (float)left);
    omlr[j] = __fsub_rn(1.0f, lr[j]);
    il[j] = (uint32_t)(sy0 * src_pitch_px + sx0 + left);
    ir[j] = (uint32_t)(sy0 * src_pitch_px + sx0 + right);
  }
  int cur = -2;                                  // source row held in hc (hn holds min(cur + 1, sh - 1))
  float hc[kResizeCols][3], hn[kResizeCols][3];
#pragma unroll
  for (int j = 0; j < kResizeCols; ++j)
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) hc[j][ch] = hn[j][ch] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float y = __fmul_rn((float)r, ys);
    const int top = min((int)(pv_floor_bits(y) & 0x7FFFFFu), sh - 1);
    const int bot = min(top + 1, sh - 1);
    const float tb = __fsub_rn(y, (float)top), omtb = __fsub_rn(1.0f, tb);
    if (top != cur) {                            // warp-uniform: rows depend on r only
      const uint32_t o_top = (uint32_t)(top * src_pitch_px), o_bot = (uint32_t)(bot * src_pitch_px);
#pragma unroll
      for (int j = 0; j < kResizeCols; ++j) {
        if (top == cur + 1) { hc[j][0] = hn[j][0]; hc[j][1] = hn[j][1]; hc[j][2] = hn[j][2]; }
        else hrow<SRC_CH>(s, il[j] + o_top, ir[j] + o_top, lr[j], omlr[j], magic, hc[j]);
        if (bot == top) { hn[j][0] = hc[j][0]; hn[j][1] = hc[j][1]; hn[j][2] = hc[j][2]; }
        else hrow<SRC_CH>(s, il[j] + o_bot, ir[j] + o_bot, lr[j], omlr[j], magic, hn[j]);
      }
      cur = top;
    }
#pragma unroll
    for (int j = 0; j < kResizeCols; ++j) {
      uint32_t q[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float v = __fadd_rn(__fmul_rn(omtb, hc[j][ch]), __fmul_rn(tb, hn[j][ch]));
        // floor(v + 0.5) clamped to [0, 255]: v >= 0, so only the upper clamp can bind
        q[ch] = pv_floor_bits(fminf(__fadd_rn(v, 0.5f), 255.5f));
      }
      // q[ch] = 0x4B0000vv: byte0 = q0, byte1 = q1, byte2 = q2, byte3 = 255
      const uint32_t t01 = __byte_perm(q[0], q[1], 0x0040u);            // [q0.b0, q1.b0, -, -]
      const uint32_t t2a = __byte_perm(q[2], 0xFF000000u, 0x0070u);     // [q2.b0, 0xFF, -, -]
      if (on[j]) d[256 * j] = __byte_perm(t01, t2a, 0x5410u);
    }
    d += dst_pitch_px;
  }
}

// ---- the same kernel on sm_100a's packed fp32 pipe (FADD2 / FFMA2): the two columns of a thread travel as f32x2 -----------
// Bit-identical arithmetic: a product is fma(a, b, nz) with nz = -0 passed at run time (exactly the rounded product; ptxas
// 12.9 contracts a mul.rn.f32x2 / add.rn.f32x2 pair — and an fma with a literal -0 — into ONE FFMA2 with a single
// rounding, even with -fmad=false), sums are add.rn.f32x2, floor(t) is add.rm.f32x2 with 2^23.  Half the FP instructions per pixel pair.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// nz = (-0, -0) arrives as a KERNEL ARGUMENT: x * y + (-0) is exactly the rounded product (sign of zero included), and
// because ptxas cannot see the addend's value it cannot turn the fma back into a mul and contract it with the next add
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b, f2_t nz) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
  return r;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t f2_add_rd(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
template <int SEL>
__device__ __forceinline__ uint32_t pv_u8_sel_bits(uint32_t v, uint32_t magic) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(magic), "n"(0x7540 | SEL));
  return r;
}

// horizontally interpolated source row of BOTH columns: h[ch] = omlr * S[il][ch] + lr * S[ir][ch], packed (col 0, col 1)
template <int SRC_CH>
__device__ __forceinline__ void hrow2(const uint8_t* __restrict__ s, const uint32_t (&il)[2], const uint32_t (&ir)[2], uint32_t off,
                                      f2_t lr2, f2_t omlr2, uint32_t magic, f2_t nz, f2_t (&h)[3]) {
  const f2_t m23 = f2_pack(-8388608.0f, -8388608.0f);
  uint32_t a[2][3], b[2][3];
  if (SRC_CH == 4) {
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(s);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t va = __ldg(s4 + il[j] + off), vb = __ldg(s4 + ir[j] + off);
      a[j][0] = pv_u8_sel_bits<0>(va, magic); a[j][1] = pv_u8_sel_bits<1>(va, magic); a[j][2] = pv_u8_sel_bits<2>(va, magic);
      b[j][0] = pv_u8_sel_bits<0>(vb, magic); b[j][1] = pv_u8_sel_bits<1>(vb, magic); b[j][2] = pv_u8_sel_bits<2>(vb, magic);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint8_t* pl = s + (il[j] + off) * 3u;
      const uint8_t* pr = s + (ir[j] + off) * 3u;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        a[j][ch] = 0x4B000000u | (uint32_t)__ldg(pl + ch);
        b[j][ch] = 0x4B000000u | (uint32_t)__ldg(pr + ch);
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const f2_t L = f2_add(f2_pack(__uint_as_float(a[0][ch]), __uint_as_float(a[1][ch])), m23);   // exact: (2^23 + v) - 2^23
    const f2_t R = f2_add(f2_pack(__uint_as_float(b[0][ch]), __uint_as_float(b[1][ch])), m23);
    h[ch] = f2_add(f2_mul(omlr2, L, nz), f2_mul(lr2, R, nz));
  }
}

template <int SRC_CH>
__global__ void __launch_bounds__(256) resize_cols2_kernel(const uint8_t* __restrict__ src, long long src_img_stride,
                                                           int src_pitch_px, int sx0, int sy0, int sw, int sh,
                                                           uchar4* __restrict__ dst, long long dst_img_stride, int dst_pitch_px,
                                                           int dx0, int dy0, int dw, int dh, float xs, float ys, uint32_t magic, f2_t nz) {
  const int c0 = blockIdx.x * (256 * kResizeCols) + threadIdx.x;
  if (c0 >= dw) return;
  const int r0 = blockIdx.y * kResizeRows;
  const int r1 = min(r0 + kResizeRows, dh);
  const uint8_t* s = src + (long long)blockIdx.z * src_img_stride;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst) + (long long)blockIdx.z * dst_img_stride + (long long)(dy0 + r0) * dst_pitch_px + dx0 + c0;
  uint32_t il[2], ir[2];
  float lr[2], omlr[2];
  bool on[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = c0 + 256 * j;
    on[j] = c < dw;
    const float x = __fmul_rn((float)min(c, dw - 1), xs);
    const int left = min((int)(pv_floor_bits(x) & 0x7FFFFFu), sw - 1);
    const int right = min(left + 1, sw - 1);
    lr[j] = __fsub_rn(x, (float)left);
    omlr[j] = __fsub_rn(1.0f, lr[j]);
    il[j] = (uint32_t)(sy0 * src_pitch_px + sx0 + left);
    ir[j] = (uint32_t)(sy0 * src_pitch_px + sx0 + right);
  }
  const f2_t lr2 = f2_pack(lr[0], lr[1]), omlr2 = f2_pack(omlr[0], omlr[1]);
  const f2_t half2 = f2_pack(0.5f, 0.5f), big2 = f2_pack(8388608.0f, 8388608.0f);
  int cur = -2;
  f2_t hc[3], hn[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) hc[ch] = hn[ch] = 0ull;
  for (int r = r0; r < r1; ++r) {
    const float y = __fmul_rn((float)r, ys);
    const int top = min((int)(pv_floor_bits(y) & 0x7FFFFFu), sh - 1);
    const int bot = min(top + 1, sh - 1);
    const float tb = __fsub_rn(y, (float)top), omtb = __fsub_rn(1.0f, tb);
    if (top != cur) {                            // warp-uniform: rows depend on r only
      if (top == cur + 1) { hc[0] = hn[0]; hc[1] = hn[1]; hc[2] = hn[2]; }
      else hrow2<SRC_CH>(s, il, ir, (uint32_t)(top * src_pitch_px), lr2, omlr2, magic, nz, hc);
      if (bot == top) { hn[0] = hc[0]; hn[1] = hc[1]; hn[2] = hc[2]; }
      else hrow2<SRC_CH>(s, il, ir, (uint32_t)(bot * src_pitch_px), lr2, omlr2, magic, nz, hn);
      cur = top;
    }
    const f2_t tb2 = f2_pack(tb, tb), omtb2 = f2_pack(omtb, omtb);
    uint32_t q[2][3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const f2_t v = f2_add(f2_mul(omtb2, hc[ch], nz), f2_mul(tb2, hn[ch], nz));
      float v0, v1;
      f2_unpack(f2_add(v, half2), v0, v1);
      // floor(v + 0.5) clamped to [0, 255]: v >= 0, so only the upper clamp can bind
      float f0, f1;
      f2_unpack(f2_add_rd(f2_pack(fminf(v0, 255.5f), fminf(v1, 255.5f)), big2), f0, f1);
      q[0][ch] = __float_as_uint(f0);
      q[1][ch] = __float_as_uint(f1);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t t01 = __byte_perm(q[j][0], q[j][1], 0x0040u);
      const uint32_t t2a = __byte_perm(q[j][2], 0xFF000000u, 0x0070u);
      if (on[j]) d[256 * j] = __byte_perm(t01, t2a, 0x5410u);
    }
    d += dst_pitch_px;
  }
}

// 2-D launch: blockIdx.z = image; a CTA covers 8 rows x 128 columns, each thread 4 consecutive columns
// (independent bilinear samples in flight, 16 contiguous bytes written per thread)
template <int SRC_CH>
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const uint8_t* __restrict__ src, long long src_img_stride,
                                                              int src_pitch_px, int sx0, int sy0, int sw, int sh,
                                                              uchar4* __restrict__ dst, long long dst_img_stride,
                                                              int dst_pitch_px, int dx0, int dy0, int dw, int dh, float xs,
                                                              float ys, int B, int copy_only) {
  const int c0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const int r = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (c0 >= dw || r >= dh) return;
  const uint8_t* s = src + (long long)n * src_img_stride;
  uchar4 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = min(c0 + k, dw - 1);
    if (copy_only) {
      const uint8_t* p = s + ((long long)(sy0 + r) * src_pitch_px + sx0 + c) * SRC_CH;
      o[k] = make_uchar4(p[0], p[1], p[2], 255);
    } else {
      o[k] = bilinear_px<SRC_CH>(s, src_pitch_px, sx0, sy0, sw, sh, r, c, xs, ys);
    }
  }
  uchar4* d = dst + (long long)n * dst_img_stride + (long long)(dy0 + r) * dst_pitch_px + dx0 + c0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (c0 + k < dw) d[k] = o[k];
}

// the small tail of the pyramid (levels k0 .. L-1, each resized from its predecessor inside the plane):
// one CTA per image walks the levels in order, __syncthreads() between levels.
struct PyrLevel {
  int x0, y0, w, h;
  float xs, ys;   // scale from the previous level
};
constexpr int kMaxTail = 40;
struct PyrTail {
  int n;                     // number of levels to build
  PyrLevel prev;             // rectangle of level k0-1 (source of the first one)
  PyrLevel lv[kMaxTail];
};

constexpr int kTailCluster = 8;   // CTAs cooperating on one image (thread-block cluster, barrier.cluster between levels)

__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(1024)
    pyramid_tail_kernel(uchar4* __restrict__ plane, long long img_stride_px, int pitch_px, const __grid_constant__ PyrTail t) {
  uchar4* img = plane + (long long)(blockIdx.x / kTailCluster) * img_stride_px;
  const int rank = blockIdx.x % kTailCluster;
  PyrLevel src = t.prev;
  for (int k = 0; k < t.n; ++k) {
    const PyrLevel d = t.lv[k];
    const int total = d.w * d.h;
    for (int i = rank * 1024 + threadIdx.x; i < total; i += kTailCluster * 1024) {
      const int r = i / d.w, c = i - r * d.w;
      img[(long long)(d.y0 + r) * pitch_px + d.x0 + c] =
          bilinear_px<4>(reinterpret_cast<const uint8_t*>(img), pitch_px, src.x0, src.y0, src.w, src.h, r, c, d.xs, d.ys);
    }
    // level k complete and visible to the whole cluster before level k+1 reads it
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    src = d;
  }
}

// score[n,y,x] = bias + sum_kw D[(n*Hq + y)*Wq + x + kw][kw]   (D: fp32 rows of `cols` columns)
__global__ void det_shift_sum_kernel(const float* __restrict__ D, int B, int Hq, int Wq, int cols, int OH, int OW,
                                     int KW, float bias, float* __restrict__ scores) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * OH * OW) return;
  const int x = (int)(idx % OW);
  long long r = idx / OW;
  const int y = (int)(r % OH);
  const int n = (int)(r / OH);
  const float* base = D + (((long long)n * Hq + y) * Wq + x) * cols;
  float s = bias;
  for (int kw = 0; kw < KW; ++kw) s += base[(long long)kw * cols + kw];
  scores[idx] = s;
}

// un-padded NHWC partials (detconv): out-of-range columns are the conv's zero padding
__global__ void det_shift_sum_nhwc_kernel(const float* __restrict__ P, int B, int H, int W, int pitch, int cols, int KW,
                                          float bias, float* __restrict__ scores) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int x = (int)(idx % W);
  const long long r = idx / W;   // n*H + y
  const float* row = P + r * pitch * cols;
  float s = bias;
  const int half = KW / 2;
  for (int kw = 0; kw < KW; ++kw) {
    const int xx = x + kw - half;
    if (xx >= 0 && xx < W) s += row[(long long)xx * cols + kw];
  }
  scores[idx] = s;
}

// ---- candidates -----------------------------------------------------------------------------
__global__ void det_candidates_kernel(const float* __restrict__ scores, int B, int cells, float thr,
                                      int* __restrict__ counts, float* __restrict__ cand_score,
                                      int* __restrict__ cand_cell, int cap) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * cells) return;
  const float s = scores[idx];
  if (s > thr) {
    const int n = (int)(idx / cells);
    const int cell = (int)(idx - (long long)n * cells);
    const int slot = atomicAdd(&counts[n], 1);
    if (slot < cap) {
      cand_score[(long long)n * cap + slot] = s;
      cand_cell[(long long)n * cap + slot] = cell;
    }
  }
}

struct DetGeom {
  int n_levels;
  int window;
  int ow;         // score-map width
  int cell_mul;   // plane = cell * cell_mul + cell_add
  int cell_add;
  double iou_thresh, covered_thresh;
};

__device__ __forceinline__ long long rect_area(int l, int t, int r, int b) {
  if (r < l || b < t) return 0;
  return (long long)(r - l + 1) * (long long)(b - t + 1);
}

__device__ __forceinline__ bool boxes_overlap(const int4 a, const int4 b, double iou, double cov) {
  const long long inner = rect_area(max(a.x, b.x), max(a.y, b.y), min(a.z, b.z), min(a.w, b.w));
  if (inner == 0) return false;
  const long long aa = rect_area(a.x, a.y, a.z, a.w), ab = rect_area(b.x, b.y, b.z, b.w);
  const long long outer = aa + ab - inner;
  if ((double)inner / (double)outer > iou) return true;
  if ((double)inner / (double)aa > cov || (double)inner / (double)ab > cov) return true;
  return false;
}

constexpr int kNmsCap = 4096;

// one CTA per frame.  key = (score desc, cell asc) — the same total order as the oracle.
__global__ void __launch_bounds__(1024) det_nms_kernel(const int* __restrict__ counts, const float* __restrict__ cand_score,
                                                       const int* __restrict__ cand_cell, int cap,
                                                       const int* __restrict__ rects,   // [L,4] x0,y0,w,h
                                                       const float* __restrict__ fxy,   // [L,2]
                                                       DetGeom g, int max_det, int* __restrict__ out_boxes,
                                                       float* __restrict__ out_scores, int* __restrict__ out_counts) {
  extern __shared__ uint8_t sm[];
  float* s_score = reinterpret_cast<float*>(sm);
  int* s_cell = reinterpret_cast<int*>(s_score + kNmsCap);
  int4* s_box = reinterpret_cast<int4*>(s_cell + kNmsCap);
  uint8_t* s_dead = reinterpret_cast<uint8_t*>(s_box + kNmsCap);
  __shared__ int s_nkept;
  const int n = blockIdx.x;
  int cnt = counts[n];
  if (cnt > cap || cnt > kNmsCap) {  // overflow: report, do not guess
    if (threadIdx.x == 0) out_counts[n] = -cnt;
    return;
  }
  int npow = 1;
  while (npow < cnt) npow <<= 1;
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    if (i < cnt) {
      s_score[i] = cand_score[(long long)n * cap + i];
      s_cell[i] = cand_cell[(long long)n * cap + i];
    } else {
      s_score[i] = -INFINITY;
      s_cell[i] = 0x7fffffff;
    }
  }
  __syncthreads();
  // bitonic sort: "a before b" iff score_a > score_b or (== and cell_a < cell_b)
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float sa = s_score[i], sb = s_score[ixj];
          const int ca = s_cell[i], cb = s_cell[ixj];
          const bool a_first = (sa > sb) || (sa == sb && ca < cb);
          const bool up = ((i & k) == 0);
          if (up ? !a_first : a_first) {
            s_score[i] = sb; s_score[ixj] = sa;
            s_cell[i] = cb; s_cell[ixj] = ca;
          }
        }
      }
      __syncthreads();
    }
  }
  // boxes
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int cell = s_cell[i];
    const int r = cell / g.ow, c = cell - r * g.ow;
    const int px = c * g.cell_mul + g.cell_add, py = r * g.cell_mul + g.cell_add;
    int lv = -1;
    for (int L = 0; L < g.n_levels; ++L) {
      const int x0 = rects[4 * L], y0 = rects[4 * L + 1], w = rects[4 * L + 2], h = rects[4 * L + 3];
      if (px >= x0 && px < x0 + w && py >= y0 && py < y0 + h) { lv = L; break; }
    }
    if (lv < 0) {
      s_dead[i] = 2;  // centre in padding: no box
      s_box[i] = make_int4(0, 0, -1, -1);
    } else {
      const int l = px - g.window / 2 - rects[4 * lv], t = py - g.window / 2 - rects[4 * lv + 1];
      const int rr = l + g.window - 1, bb = t + g.window - 1;
      const float fx = fxy[2 * lv], fy = fxy[2 * lv + 1];
      int4 b;
      b.x = (int)floorf(__fadd_rn(__fmul_rn((float)l, fx), 0.5f));
      b.y = (int)floorf(__fadd_rn(__fmul_rn((float)t, fy), 0.5f));
      b.z = (int)floorf(__fadd_rn(__fmul_rn((float)rr, fx), 0.5f));
      b.w = (int)floorf(__fadd_rn(__fmul_rn((float)bb, fy), 0.5f));
      s_box[i] = b;
      s_dead[i] = 0;
    }
  }
  if (threadIdx.x == 0) s_nkept = 0;
  __syncthreads();
  // greedy NMS in sorted order
  for (int i = 0; i < cnt; ++i) {
    if (s_dead[i] == 0) {  // uniform across the CTA (read after the barrier below)
      const int4 bi = s_box[i];
      if (threadIdx.x == 0) {
        const int k = s_nkept;
        if (k < max_det) {
          int* ob = out_boxes + ((long long)n * max_det + k) * 4;
          ob[0] = bi.x; ob[1] = bi.y; ob[2] = bi.z; ob[3] = bi.w;
          out_scores[(long long)n * max_det + k] = s_score[i];
        }
        s_nkept = k + 1;
      }
      for (int j = i + 1 + threadIdx.x; j < cnt; j += blockDim.x)
        if (s_dead[j] == 0 && boxes_overlap(s_box[j], bi, g.iou_thresh, g.covered_thresh)) s_dead[j] = 1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[n] = s_nkept;
}

}  // namespace

extern "C" int pv_resize_bilinear(const void* src, int src_channels, int64_t src_img_stride, int src_pitch_px, int sx0,
                                  int sy0, int sw, int sh, void* dst_rgba, int64_t dst_img_stride_px, int dst_pitch_px,
                                  int dx0, int dy0, int dw, int dh, float xs, float ys, int B, int copy_only,
                                  void* stream) {
  PV_REQUIRE(src && dst_rgba, "pv_resize_bilinear: null argument");
  PV_REQUIRE(src_channels == 3 || src_channels == 4, "pv_resize_bilinear: src_channels=%d", src_channels);
  PV_REQUIRE(sw > 0 && sh > 0 && dw > 0 && dh > 0 && B > 0, "pv_resize_bilinear: empty rect");
  const int threads = 256;
  const dim3 blocks((unsigned)((dw + 127) / 128), (unsigned)((dh + 7) / 8), (unsigned)B);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!copy_only) {
    const dim3 cb((unsigned)((dw + 256 * kResizeCols - 1) / (256 * kResizeCols)), (unsigned)((dh + kResizeRows - 1) / kResizeRows), (unsigned)B);
    // PV_RESIZE_PACKED=0 selects the scalar-fp32 kernel (same results bit for bit; the packed one halves the FP instructions)
    static int packed = -1;
    if (packed < 0) {
      const char* e = getenv("PV_RESIZE_PACKED");
      packed = (e && e[0] == '0') ? 0 : 1;
    }
    if (src_channels == 3 && packed)
      resize_cols2_kernel<3><<<cb, 256, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px, sx0, sy0, sw,
                                                sh, static_cast<uchar4*>(dst_rgba), dst_img_stride_px, dst_pitch_px, dx0, dy0,
                                                dw, dh, xs, ys, 0x4B000000u, 0x8000000080000000ull);
    else if (src_channels == 3)
      resize_cols_kernel<3><<<cb, 256, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px, sx0, sy0, sw,
                                               sh, static_cast<uchar4*>(dst_rgba), dst_img_stride_px, dst_pitch_px, dx0, dy0,
                                               dw, dh, xs, ys, 0x4B000000u);
    else if (packed)
      resize_cols2_kernel<4><<<cb, 256, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px, sx0, sy0, sw,
                                                sh, static_cast<uchar4*>(dst_rgba), dst_img_stride_px, dst_pitch_px, dx0, dy0,
                                                dw, dh, xs, ys, 0x4B000000u, 0x8000000080000000ull);
    else
      resize_cols_kernel<4><<<cb, 256, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px, sx0, sy0, sw,
                                               sh, static_cast<uchar4*>(dst_rgba), dst_img_stride_px, dst_pitch_px, dx0, dy0,
                                               dw, dh, xs, ys, 0x4B000000u);
  } else if (src_channels == 3)
    resize_bilinear_kernel<3><<<blocks, threads, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px,
                                                         sx0, sy0, sw, sh, static_cast<uchar4*>(dst_rgba),
                                                         dst_img_stride_px, dst_pitch_px, dx0, dy0, dw, dh, xs, ys, B,
                                                         copy_only);
  else
    resize_bilinear_kernel<4><<<blocks, threads, 0, s>>>(static_cast<const uint8_t*>(src), src_img_stride, src_pitch_px,
                                                         sx0, sy0, sw, sh, static_cast<uchar4*>(dst_rgba),
                                                         dst_img_stride_px, dst_pitch_px, dx0, dy0, dw, dh, xs, ys, B,
                                                         copy_only);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_pyramid_tail(void* plane_rgba, int64_t img_stride_px, int pitch_px, int B, int n_levels,
                               const int* rects_host, const float* scales_host, void* stream) {
  /* rects_host: [(n_levels+1), 4] x0,y0,w,h with the source level first; scales_host: [n_levels, 2] xs,ys */
  PV_REQUIRE(plane_rgba && rects_host && scales_host, "pv_pyramid_tail: null argument");
  PV_REQUIRE(n_levels >= 1 && n_levels <= kMaxTail, "pv_pyramid_tail: n_levels=%d", n_levels);
  PyrTail t;
  t.n = n_levels;
  t.prev.x0 = rects_host[0]; t.prev.y0 = rects_host[1]; t.prev.w = rects_host[2]; t.prev.h = rects_host[3];
  t.prev.xs = t.prev.ys = 0.f;
  for (int k = 0; k < n_levels; ++k) {
    const int* r = rects_host + 4 * (k + 1);
    t.lv[k].x0 = r[0]; t.lv[k].y0 = r[1]; t.lv[k].w = r[2]; t.lv[k].h = r[3];
    t.lv[k].xs = scales_host[2 * k];
    t.lv[k].ys = scales_host[2 * k + 1];
  }
  pyramid_tail_kernel<<<B * kTailCluster, 1024, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uchar4*>(plane_rgba), img_stride_px,
                                                                      pitch_px, t);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_det_shift_sum(const float* D, int B, int Hq, int Wq, int cols, int OH, int OW, int KW, float bias,
                                float* scores, void* stream) {
  PV_REQUIRE(D && scores, "pv_det_shift_sum: null argument");
  PV_REQUIRE(KW <= cols && OW + KW - 1 <= Wq && OH <= Hq, "pv_det_shift_sum: geometry");
  const long long total = (long long)B * OH * OW;
  const int threads = 256;
  det_shift_sum_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      D, B, Hq, Wq, cols, OH, OW, KW, bias, scores);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_det_shift_sum_nhwc(const float* P, int B, int H, int W, int pitch, int cols, int KW, float bias,
                                     float* scores, void* stream) {
  PV_REQUIRE(P && scores, "pv_det_shift_sum_nhwc: null argument");
  PV_REQUIRE(KW <= cols && pitch >= W && B > 0 && H > 0 && W > 0, "pv_det_shift_sum_nhwc: geometry");
  const long long total = (long long)B * H * W;
  const int threads = 256;
  det_shift_sum_nhwc_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      P, B, H, W, pitch, cols, KW, bias, scores);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_det_candidates(const float* scores, int B, int cells, float thr, int* counts, float* cand_score,
                                 int* cand_cell, int cap, void* stream) {
  PV_REQUIRE(scores && counts && cand_score && cand_cell, "pv_det_candidates: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PV_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int) * B, s));
  const long long total = (long long)B * cells;
  const int threads = 256;
  det_candidates_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(scores, B, cells, thr, counts,
                                                                                       cand_score, cand_cell, cap);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_det_nms(const int* counts, const float* cand_score, const int* cand_cell, int cap, int B,
                          const int* level_rects, const float* level_fxy, int n_levels, int window, int ow,
                          int cell_mul, int cell_add, double iou_thresh, double covered_thresh, int max_det,
                          int* out_boxes, float* out_scores, int* out_counts, void* stream) {
  PV_REQUIRE(counts && cand_score && cand_cell && level_rects && level_fxy && out_boxes && out_scores && out_counts,
             "pv_det_nms: null argument");
  PV_REQUIRE(cap <= kNmsCap, "pv_det_nms: cap=%d exceeds %d", cap, kNmsCap);
  DetGeom g;
  g.n_levels = n_levels;
  g.window = window;
  g.ow = ow;
  g.cell_mul = cell_mul;
  g.cell_add = cell_add;
  g.iou_thresh = iou_thresh;
  g.covered_thresh = covered_thresh;
  const size_t smem = kNmsCap * (sizeof(float) + sizeof(int) + sizeof(int4) + 1);
  static unsigned long long attr_set = 0;
  if (pv_attr_needed(&attr_set)) {
    PV_CUDA_CHECK(cudaFuncSetAttribute(det_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  det_nms_kernel<<<B, 1024, smem, static_cast<cudaStream_t>(stream)>>>(counts, cand_score, cand_cell, cap, level_rects,
                                                                      level_fxy, g, max_det, out_boxes, out_scores,
                                                                      out_counts);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
