// tracker.cu — bank of DSST correlation trackers, one CTA per live track, all tracks of a frame in
// ONE launch.  Replaces dlib.correlation_tracker.start_track / update / get_position called per
// tracker per frame from Python (pyannote/video/tracking.py:203,231,250-251).
//
// Per track state in HBM (float32): A[32][33x64] complex numerators (31 FHOG planes + the brightness plane) (half plane: the spectra of real features are
// Hermitian), B[33x64] denominator, position (l,t,r,b).  update = chip (bilinear, rect*1.4 -> 64x64) -> FHOG-31 (cell 1)
// x cosine window -> 16 packed 2-D FFTs (shared memory, radix-8) -> response = ifft2(sum F_i conj(A_i)/(B+lambda))
// -> argmax / sub-pixel / PSR (warp-shuffle + shared reductions) -> position -> filter update from the spectra that
// pass 1 spilled to a scratch plane (0.5 MB per track, re-read by the thread that wrote it).
// A second kernel runs the 1-D scale filter (32 scales, FHOG cell 4).  Mirrors oracle/dsst.py step by step.
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int FS = 64;
constexpr int NPIX = FS * FS;
constexpr int NCH = 32;            // 31 FHOG planes + the chip intensity / 255 (dlib make_chip's 32nd feature)
constexpr int kThreads = 512;

struct TrackerTables {
  float hann[FS];
  float uu[9], vv[9];
  float tw_re[32], tw_im[32];  // exp(-2 pi i k / 64)
};

struct BankParams {
  float2* A;        // [cap][31][2112]  numerators, half plane ky <= 32 (Hermitian spectra of real features)
  float2* F;        // [cap][31][2112]  scratch: the spectra of the current update, written by pass 1, read by pass 2
  float* B;         // [cap][2112]
  float* pos;       // [cap][4]
  float* psr;       // [cap]
  const int* ids;   // [n] slots handled by this launch
  const float* rects;  // start only: [n][4]
  const uint8_t* frame;  // uint8 [F,H,W,3]
  const int* frame_idx;  // [n] frame of each track of this launch (nullptr: frame 0)
  int H, W;
  float padding, lambda, nu;
};

// u8 -> f32 without the conversion unit (I2F runs at a quarter of the FP32 rate): (2^23 | b) - 2^23 is exact
__device__ __forceinline__ float u8f(uint32_t b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }

__device__ __forceinline__ int rev6(int v) { return (int)(__brev((unsigned)v) >> 26); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void bfly(float2& a, float2& b, const float2 w) {
  const float2 t = cmul(w, b);
  b = make_float2(a.x - t.x, a.y - t.y);
  a = make_float2(a.x + t.x, a.y + t.y);
}
__device__ __forceinline__ void bfly1(float2& a, float2& b) {   // twiddle 1
  const float2 t = b;
  b = make_float2(a.x - t.x, a.y - t.y);
  a = make_float2(a.x + t.x, a.y + t.y);
}

constexpr int PS = FS + 1;        // padded row pitch of the complex work plane (float2 elements): a warp that walks 32
                                  // rows at one column then touches 16 distinct 8-byte bank pairs per half-warp

// in-place 2-D DIT FFT of the 64x64 complex tile `x` (row pitch PS) whose input was stored bit-reversed in both
// dimensions; output in natural order.  inverse: conjugate twiddles (no scaling).  Radix-8: two passes per dimension
// (stages 1-3 on 8 consecutive elements with constant twiddles, stages 4-6 on elements 8 apart), i.e. 4 block barriers
// and 4 shared-memory round trips per 2-D transform (the radix-4 version needed 6, the radix-2 one 12).  Thread t owns
// line t & 63 and butterfly t >> 6, so a warp always walks 32 adjacent LINES: consecutive float2 for the column
// transform, pitch-65 rows for the row transform — both free of bank conflicts — and the twiddles are warp-uniform.
__device__ void fft2_64(float2* x, const TrackerTables& tb, bool inverse) {
  const int line = threadIdx.x & 63, q = threadIdx.x >> 6;     // 512 threads = 64 lines x 8 butterflies
  const float sg = inverse ? -1.f : 1.f;
  const float r = 0.70710678118654752f;
  const float2 w8 = make_float2(r, -sg * r), w16 = make_float2(0.f, -sg), w24 = make_float2(-r, -sg * r);
  for (int dim = 0; dim < 2; ++dim) {
    const int es = dim == 0 ? 1 : PS;   // element stride along the transformed dimension
    const int ls = dim == 0 ? PS : 1;   // stride between lines
    {
      // ---- stages 1..3: elements 8q .. 8q+7 of the line ----
      float2* p = x + line * ls + (q << 3) * es;
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = p[k * es];
      bfly1(v[0], v[1]); bfly1(v[2], v[3]); bfly1(v[4], v[5]); bfly1(v[6], v[7]);
      bfly1(v[0], v[2]); bfly(v[1], v[3], w16); bfly1(v[4], v[6]); bfly(v[5], v[7], w16);
      bfly1(v[0], v[4]); bfly(v[1], v[5], w8); bfly(v[2], v[6], w16); bfly(v[3], v[7], w24);
#pragma unroll
      for (int k = 0; k < 8; ++k) p[k * es] = v[k];
    }
    __syncthreads();
    {
      // ---- stages 4..6: elements q + 8k of the line; twiddles exp(-2 pi i m / 64), m < 32, depend on q only ----
      float2* p = x + line * ls + q * es;
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = p[(k << 3) * es];
      const float2 t4 = make_float2(tb.tw_re[4 * q], sg * tb.tw_im[4 * q]);
      bfly(v[0], v[1], t4); bfly(v[2], v[3], t4); bfly(v[4], v[5], t4); bfly(v[6], v[7], t4);
      const float2 t2a = make_float2(tb.tw_re[2 * q], sg * tb.tw_im[2 * q]);
      const float2 t2b = make_float2(tb.tw_re[2 * q + 16], sg * tb.tw_im[2 * q + 16]);
      bfly(v[0], v[2], t2a); bfly(v[4], v[6], t2a); bfly(v[1], v[3], t2b); bfly(v[5], v[7], t2b);
      bfly(v[0], v[4], make_float2(tb.tw_re[q], sg * tb.tw_im[q]));
      bfly(v[1], v[5], make_float2(tb.tw_re[q + 8], sg * tb.tw_im[q + 8]));
      bfly(v[2], v[6], make_float2(tb.tw_re[q + 16], sg * tb.tw_im[q + 16]));
      bfly(v[3], v[7], make_float2(tb.tw_re[q + 24], sg * tb.tw_im[q + 24]));
#pragma unroll
      for (int k = 0; k < 8; ++k) p[(k << 3) * es] = v[k];
    }
    __syncthreads();
  }
}

constexpr int NI = FS + 1;        // inverse-norm plane: entry (y0 + 1, x0 + 1) = 1 / sqrt(block(y0, x0) + eps), y0, x0 in [-1, 63]

struct Smem {
  float2* plane;   // [64 * PS]   complex work plane (the chip lives here before the first transform)
  float* mag;      // [4096]      gradient magnitude
  float* osum;     // [4096]      0.5 * sum of the four clipped normalised magnitudes
  float* invn;     // [NI * NI]   inverse block norms
  float* bsum;     // [4096]      sum over channels of |F|^2
  float* red;      // [64]
  int* redi;       // [32]
  uint8_t* ori;    // [4096]      snapped orientation 0..17
  uint8_t* gray;   // [4096]      chip intensity (r + g + b) / 3
  uint8_t* chip;   // [4096 * 3]  aliases `plane`
};

__device__ Smem carve(uint8_t* base) {
  Smem s;
  s.plane = reinterpret_cast<float2*>(base);
  s.mag = reinterpret_cast<float*>(s.plane + FS * PS);
  s.osum = s.mag + NPIX;
  s.invn = s.osum + NPIX;
  s.bsum = s.invn + NI * NI + 3;
  s.red = s.bsum + NPIX;
  s.redi = reinterpret_cast<int*>(s.red + 64);
  s.ori = reinterpret_cast<uint8_t*>(s.redi + 32);
  s.gray = s.ori + NPIX;
  s.chip = reinterpret_cast<uint8_t*>(s.plane);
  return s;
}
constexpr size_t kSmemBytes = (size_t)FS * PS * 8 + 3 * NPIX * 4 + (NI * NI + 3) * 4 + 64 * 4 + 32 * 4 + 2 * NPIX;
static_assert((size_t)NPIX * 3 <= (size_t)FS * PS * 8, "chip must fit into the work plane");
static_assert(2 * (kSmemBytes + 2048) <= 227 * 1024, "two CTAs per SM");

// chip + gradient orientation/magnitude + inverse block norms + per-cell orientation feature value
__device__ void features_prepare(const BankParams& p, const TrackerTables& tb, const float* rect, Smem& s, float* tf) {
  const int tid = threadIdx.x;
  const float l = rect[0], t = rect[1], r = rect[2], b = rect[3];
  const float cx = __fmul_rn(__fadd_rn(l, r), 0.5f), cy = __fmul_rn(__fadd_rn(t, b), 0.5f);
  const float hw = __fmul_rn(__fmul_rn(__fsub_rn(r, l), 0.5f), p.padding);
  const float hh = __fmul_rn(__fmul_rn(__fsub_rn(b, t), 0.5f), p.padding);
  const float rl = __fsub_rn(cx, hw), rt = __fsub_rn(cy, hh);
  const float sx = __fdiv_rn(__fmul_rn(2.0f, hw), (float)(FS - 1)), sy = __fdiv_rn(__fmul_rn(2.0f, hh), (float)(FS - 1));
  if (tid == 0) { tf[0] = rl; tf[1] = rt; tf[2] = sx; tf[3] = sy; }
  for (int i = tid; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    const float fx = __fadd_rn(rl, __fmul_rn((float)x, sx)), fy = __fadd_rn(rt, __fmul_rn((float)y, sy));
    const int left = (int)floorf(fx), top = (int)floorf(fy);
    uint8_t o[3] = {0, 0, 0};
    // left >= 0 && left + 1 < W, tested on the floats: a runaway box (inf / NaN / beyond int range) must read nothing
    if (fx >= 0.f && fx < (float)(p.W - 1) && fy >= 0.f && fy < (float)(p.H - 1)) {
      const float lr = __fsub_rn(fx, (float)left), tbv = __fsub_rn(fy, (float)top);
      const float omlr = __fsub_rn(1.0f, lr), omtb = __fsub_rn(1.0f, tbv);
      const uint8_t* ptl = p.frame + ((long long)top * p.W + left) * 3;
      const uint8_t* pbl = ptl + (long long)p.W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = __fadd_rn(__fmul_rn(omlr, u8f(ptl[c])), __fmul_rn(lr, u8f(ptl[3 + c])));
        const float bb = __fadd_rn(__fmul_rn(omlr, u8f(pbl[c])), __fmul_rn(lr, u8f(pbl[3 + c])));
        float v = __fadd_rn(__fmul_rn(omtb, a), __fmul_rn(tbv, bb));
        v = fminf(fmaxf(floorf(__fadd_rn(v, 0.5f)), 0.f), 255.f);
        o[c] = (uint8_t)v;
      }
    }
    s.chip[3 * i] = o[0];
    s.chip[3 * i + 1] = o[1];
    s.chip[3 * i + 2] = o[2];
  }
  __syncthreads();
  for (int i = tid; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    float m = 0.f;
    int bo = 0;
    if (y > 0 && y < FS - 1 && x > 0 && x < FS - 1) {
      float gx = 0.f, gy = 0.f, best = -1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float dx = __fsub_rn(u8f(s.chip[3 * (i + 1) + c]), u8f(s.chip[3 * (i - 1) + c]));
        const float dy = __fsub_rn(u8f(s.chip[3 * (i + FS) + c]), u8f(s.chip[3 * (i - FS) + c]));
        const float v = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (v > best) { best = v; gx = dx; gy = dy; }
      }
      m = sqrtf(best);
      float best_dot = 0.f;
#pragma unroll
      for (int o = 0; o < 9; ++o) {
        const float dot = __fadd_rn(__fmul_rn(tb.uu[o], gx), __fmul_rn(tb.vv[o], gy));
        if (dot > best_dot) { best_dot = dot; bo = o; }
        else if (-dot > best_dot) { best_dot = -dot; bo = o + 9; }
      }
    }
    s.mag[i] = m;
    s.ori[i] = (uint8_t)bo;
    s.gray[i] = (uint8_t)(((uint32_t)s.chip[3 * i] + s.chip[3 * i + 1] + s.chip[3 * i + 2]) / 3u);   // dlib assign_pixel rgb -> gray
  }
  __syncthreads();
  // inverse norm of every 2x2 block of squared magnitudes, once per update (it was recomputed — 4 loads, a square root
  // and a division — 4 times per pixel for the orientation sum and once per pixel for each of the 4 texture channels)
  for (int j = tid; j < NI * NI; j += kThreads) {
    const int y0 = j / NI - 1, x0 = j - (y0 + 1) * NI - 1;
    float q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
      float m = 0.f;
      if (yy >= 0 && yy < FS && xx >= 0 && xx < FS) m = s.mag[yy * FS + xx];
      q[k] = __fmul_rn(m, m);
    }
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(q[0], q[1]), q[2]), q[3]);
    s.invn[j] = __fdiv_rn(1.0f, sqrtf(__fadd_rn(sum, 0.0001f)));
  }
  __syncthreads();
}

__device__ __forceinline__ float inv_block(const Smem& s, int y0, int x0) { return s.invn[(y0 + 1) * NI + x0 + 1]; }

__device__ void features_osum(Smem& s) {
  for (int i = threadIdx.x; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    const float m = s.mag[i];
    const float h0 = fminf(__fmul_rn(m, inv_block(s, y - 1, x - 1)), 0.2f);
    const float h1 = fminf(__fmul_rn(m, inv_block(s, y - 1, x)), 0.2f);
    const float h2 = fminf(__fmul_rn(m, inv_block(s, y, x - 1)), 0.2f);
    const float h3 = fminf(__fmul_rn(m, inv_block(s, y, x)), 0.2f);
    s.osum[i] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h0, h1), h2), h3));
  }
  __syncthreads();
}

__device__ __forceinline__ float feature_value(const Smem& s, const TrackerTables& tb, int ch, int i, int y, int x) {
  const int o = s.ori[i];
  float v;
  if (ch < 18) v = (o == ch) ? s.osum[i] : 0.f;
  else if (ch < 27) v = ((o >= 9 ? o - 9 : o) == ch - 18) ? s.osum[i] : 0.f;
  else if (ch < 31) {
    const int k = ch - 27;
    const float nk = inv_block(s, y - 1 + (k >> 1), x - 1 + (k & 1));
    v = __fmul_rn(0.2357f, fminf(__fmul_rn(s.mag[i], nk), 0.2f));
  } else {
    v = __fdiv_rn(u8f(s.gray[i]), 255.0f);          // the 32nd feature: overall brightness
  }
  return __fmul_rn(v, __fmul_rn(tb.hann[y], tb.hann[x]));
}

// two windowed REAL feature planes packed into one complex plane (cha -> re, chb -> im; chb < 0: zero),
// stored bit-reversed for the FFT.  One complex FFT then yields both spectra:
//   Fa[k] = (Z[k] + conj(Z[-k])) / 2,   Fb[k] = (Z[k] - conj(Z[-k])) / (2i)
__device__ void build_plane2(const Smem& s, const TrackerTables& tb, int cha, int chb) {
  for (int i = threadIdx.x; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    float va = 0.f, vb = 0.f;
    if (y > 0 && y < FS - 1 && x > 0 && x < FS - 1) {
      va = feature_value(s, tb, cha, i, y, x);
      if (chb >= 0) vb = feature_value(s, tb, chb, i, y, x);
    }
    s.plane[rev6(y) * PS + rev6(x)] = make_float2(va, vb);
  }
  __syncthreads();
}

__device__ __forceinline__ void split_spectra(const Smem& s, int y, int x, float2& fa, float2& fb) {
  const float2 z = s.plane[y * PS + x];
  const float2 zm = s.plane[((FS - y) & (FS - 1)) * PS + ((FS - x) & (FS - 1))];
  fa = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
  fb = make_float2(0.5f * (z.y + zm.y), 0.5f * (zm.x - z.x));
}

// FFT of the Gaussian target centred at (px,py); conj(G^) of the thread's 8 bins is returned in g[]
__device__ void target_hat(Smem& s, const TrackerTables& tb, float px, float py, float2 g[8]) {
  for (int i = threadIdx.x; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    const float dx = (float)x - px, dy = (float)y - py;
    s.plane[rev6(y) * PS + rev6(x)] = make_float2(expf(-(dx * dx + dy * dy) / 3.0f), 0.f);
  }
  __syncthreads();
  fft2_64(s.plane, tb, false);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = threadIdx.x + k * kThreads;
    const float2 v = s.plane[(i >> 6) * PS + (i & 63)];
    g[k] = make_float2(v.x, -v.y);
  }
  __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The features are REAL, so every spectrum is Hermitian, F[-k] = conj(F[k]): only the half plane ky <= 32 (33 x 64 = 2112
// bins) is stored and processed — numerators A, denominator B, and the per-update spectra F that pass 1 spills so that
// pass 2 (the filter update, which needs the peak found by pass 1) streams them back instead of redoing 16 transforms.
// Every thread owns the same bins (h = tid + 512 k, k < 4, plus h = 2048 + tid for tid < 64) for the whole kernel: the
// response accumulator and conj(G^) live in registers, the spilled spectra are re-read by the thread that wrote them (no
// synchronisation), the CTA needs 106 KB of shared memory and two CTAs (two tracks) share an SM.
constexpr int NHB = 33 * FS;          // stored half-plane bins per channel
constexpr int NB = 5;                 // bins per thread (the fifth only for tid < 64)

template <bool START>
__global__ void __launch_bounds__(kThreads, 2) tracker_kernel(BankParams p, const __grid_constant__ TrackerTables tbc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem s = carve(smem_raw);
  // the tables are indexed per lane (twiddles, window): from the constant bank every distinct index of a warp is a
  // replay, so they are copied to shared memory once
  __shared__ TrackerTables tb;
  for (int i = threadIdx.x; i < (int)(sizeof(TrackerTables) / 4); i += kThreads)
    reinterpret_cast<uint32_t*>(&tb)[i] = reinterpret_cast<const uint32_t*>(&tbc)[i];
  __syncthreads();
  __shared__ float tf[4];
  __shared__ float peak[4];  // ppx, ppy
  const int tid = threadIdx.x;
  const int slot = p.ids[blockIdx.x];
  if (p.frame_idx) p.frame += (size_t)p.frame_idx[blockIdx.x] * p.H * p.W * 3;
  float2* A = p.A + (size_t)slot * NCH * NHB;
  float2* F = p.F + (size_t)slot * NCH * NHB;
  float* B = p.B + (size_t)slot * NHB;
  float rect[4];
  if (START) {
#pragma unroll
    for (int k = 0; k < 4; ++k) rect[k] = p.rects[blockIdx.x * 4 + k];
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) rect[k] = p.pos[slot * 4 + k];
    if ((tid & 15) == 0) prefetch_l2(A + tid), prefetch_l2(A + NHB + tid);
  }
  features_prepare(p, tb, rect, s, tf);
  features_osum(s);
  const int ybase = tid >> 6, xcol = tid & 63;    // bin k of this thread: row ybase + 8 k (k < 4) or row 32 (k = 4), column xcol
  const int nb = tid < FS ? NB : NB - 1;

  if (START) {
    float2 g[8];
    target_hat(s, tb, 0.5f * (FS - 1), 0.5f * (FS - 1), g);
    // g[k] holds rows ybase + 8 k, k < 8; row 32 of the threads tid < 64 is g[4] (ybase = 0)
    for (int i = tid; i < NHB; i += kThreads) s.bsum[i] = 0.f;
    __syncthreads();
    for (int ch = 0; ch < NCH; ch += 2) {
      const bool two = ch + 1 < NCH;
      build_plane2(s, tb, ch, two ? ch + 1 : -1);
      fft2_64(s.plane, tb, false);
      for (int k = 0; k < nb; ++k) {
        const int h = tid + k * kThreads;
        float2 fa, fb;
        split_spectra(s, ybase + 8 * k, xcol, fa, fb);
        A[(size_t)ch * NHB + h] = cmul(g[k], fa);
        float bs = fa.x * fa.x + fa.y * fa.y;
        if (two) {
          A[(size_t)(ch + 1) * NHB + h] = cmul(g[k], fb);
          bs += fb.x * fb.x + fb.y * fb.y;
        }
        s.bsum[h] += bs;
      }
      __syncthreads();
    }
    for (int i = tid; i < NHB; i += kThreads) B[i] = s.bsum[i];
    if (tid < 4) p.pos[slot * 4 + tid] = rect[tid];
    if (tid == 0) p.psr[slot] = 0.f;
    return;
  }

  // ---- pass 1: response; the spectra go to the scratch plane F ----
  float2 acc[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) acc[k] = make_float2(0.f, 0.f);
  for (int i = tid; i < NHB; i += kThreads) s.bsum[i] = 0.f;
  __syncthreads();
  for (int ch = 0; ch < NCH; ch += 2) {
    const bool two = ch + 1 < NCH;
    if (ch + 2 < NCH && (tid & 15) == 0) {   // the next pair's numerators travel to L2 under this pair's transform
      prefetch_l2(A + (size_t)(ch + 2) * NHB + tid);
      if (ch + 3 < NCH) prefetch_l2(A + (size_t)(ch + 3) * NHB + tid);
    }
    build_plane2(s, tb, ch, two ? ch + 1 : -1);
    fft2_64(s.plane, tb, false);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (k < nb) {
        const int h = tid + k * kThreads;
        float2 fa, fb;
        split_spectra(s, ybase + 8 * k, xcol, fa, fb);
        const float2 a = A[(size_t)ch * NHB + h];
        F[(size_t)ch * NHB + h] = fa;
        acc[k].x += fa.x * a.x + fa.y * a.y;   // f * conj(a)
        acc[k].y += fa.y * a.x - fa.x * a.y;
        float bs = fa.x * fa.x + fa.y * fa.y;
        if (two) {
          const float2 b2 = A[(size_t)(ch + 1) * NHB + h];
          F[(size_t)(ch + 1) * NHB + h] = fb;
          acc[k].x += fb.x * b2.x + fb.y * b2.y;
          acc[k].y += fb.y * b2.x - fb.x * b2.y;
          bs += fb.x * fb.x + fb.y * fb.y;
        }
        s.bsum[h] += bs;
      }
    }
    __syncthreads();
  }
  {
    const float nu = p.nu, om = 1.0f - p.nu;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      if (k < nb) {
        const int h = tid + k * kThreads;
        const int y = ybase + 8 * k;
        const float bold = B[h];
        const float d = 1.0f / (bold + p.lambda);
        const float2 v = make_float2(acc[k].x * d, acc[k].y * d);
        s.plane[rev6(y) * PS + rev6(xcol)] = v;
        if (y >= 1 && y <= 31)               // the other half plane: R^[-k] = conj(R^[k]) (the response is real)
          s.plane[rev6(FS - y) * PS + rev6((FS - xcol) & (FS - 1))] = make_float2(v.x, -v.y);
        B[h] = om * bold + nu * s.bsum[h];     // the denominator's running update needs nothing from pass 2
      }
    }
  }
  __syncthreads();
  fft2_64(s.plane, tb, true);
  // real response (scaled by 1/4096) stays in the work plane (.x) until the target of pass 2 overwrites it
  float bestv = -INFINITY;
  int besti = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = tid + k * kThreads;
    float2* e = s.plane + (ybase + 8 * k) * PS + xcol;
    const float v = e->x * (1.0f / NPIX);
    e->x = v;
    if (v > bestv || (v == bestv && i < besti)) { bestv = v; besti = i; }
  }
  // block argmax (first occurrence)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
  }
  if ((tid & 31) == 0) { s.red[tid >> 5] = bestv; s.redi[tid >> 5] = besti; }
  __syncthreads();
  if (tid < 32) {
    float v = tid < kThreads / 32 ? s.red[tid] : -INFINITY;
    int ii = tid < kThreads / 32 ? s.redi[tid] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ii, o);
      if (ov > v || (ov == v && oi < ii)) { v = ov; ii = oi; }
    }
    if (tid == 0) { s.red[32] = v; s.redi[16] = ii; }
  }
  __syncthreads();
  const int pi = s.redi[16];
  const int py = pi >> 6, px = pi & 63;
  const float rmax = s.red[32];
  // PSR statistics outside the 8x8 window [px-4,px+3] x [py-4,py+3]
  double sum = 0.0, sumsq = 0.0;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int y = ybase + 8 * k, x = xcol;
    if (y >= py - 4 && y < py + 4 && x >= px - 4 && x < px + 4) continue;
    const double v = (double)s.plane[y * PS + x].x;
    sum += v;
    sumsq += v * v;
    ++cnt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ double dsum[kThreads / 32], dsq[kThreads / 32];
  __shared__ int dcnt[kThreads / 32];
  if ((tid & 31) == 0) { dsum[tid >> 5] = sum; dsq[tid >> 5] = sumsq; dcnt[tid >> 5] = cnt; }
  __syncthreads();
  if (tid == 0) {
    double S = 0, Q = 0;
    int n = 0;
    for (int w = 0; w < kThreads / 32; ++w) { S += dsum[w]; Q += dsq[w]; n += dcnt[w]; }
    const double mean = S / n;
    const double var = (Q - n * mean * mean) / (n - 1);
    const double psr = ((double)rmax - mean) / sqrt(var > 0 ? var : 1e-300);
    double ppx = px, ppy = py;
    if (px > 0 && px < FS - 1 && py > 0 && py < FS - 1) {
      const float2* e = s.plane + py * PS + px;
      // dlib max_point_interpolated: Newton step of the 3x3 finite-difference quadratic (cross term included), clamped to +-1
      const double c = e->x, xl = e[-1].x, xr = e[1].x, yu = e[-PS].x, yd = e[PS].x;
      const double dx = 0.5 * (xr - xl), dy = 0.5 * (yd - yu);
      const double dxx = xr - 2 * c + xl, dyy = yd - 2 * c + yu;
      const double dxy = 0.25 * (((double)e[PS + 1].x + (double)e[-PS - 1].x) - ((double)e[PS - 1].x + (double)e[-PS + 1].x));
      const double det = dxx * dyy - dxy * dxy;
      if (det != 0) {
        ppx += fmin(1.0, fmax(-1.0, -(dyy * dx - dxy * dy) / det));
        ppy += fmin(1.0, fmax(-1.0, -(dxx * dy - dxy * dx) / det));
      }
    }
    const double ix = (double)tf[0] + ppx * (double)tf[2], iy = (double)tf[1] + ppy * (double)tf[3];
    const double cx = 0.5 * ((double)rect[0] + (double)rect[2]), cy = 0.5 * ((double)rect[1] + (double)rect[3]);
    const double ddx = ix - cx, ddy = iy - cy;
    p.pos[slot * 4 + 0] = (float)((double)rect[0] + ddx);
    p.pos[slot * 4 + 1] = (float)((double)rect[1] + ddy);
    p.pos[slot * 4 + 2] = (float)((double)rect[2] + ddx);
    p.pos[slot * 4 + 3] = (float)((double)rect[3] + ddy);
    p.psr[slot] = (float)psr;
    peak[0] = (float)ppx;
    peak[1] = (float)ppy;
  }
  __syncthreads();

  // ---- pass 2: filter update A <- (1 - nu) A + nu conj(G^) F, streaming the spilled spectra back (no transforms) ----
  // conj(G^) of the half plane goes to shared memory (the work plane is free now) so that the stream does not have to
  // follow the bin ownership: the 31 x 2112 bins are one flat array walked with 128-bit loads, four of them in flight per
  // operand and thread (with one 8-byte load per bin and channel this loop was 9 % of the instructions but 43 % of the
  // stall samples: latency-bound).
  float2 g[8];
  target_hat(s, tb, peak[0], peak[1], g);
  float2* gsm = s.plane;                       // [NHB]
#pragma unroll
  for (int k = 0; k < NB; ++k)
    if (k < nb) gsm[tid + k * kThreads] = g[k];
  __syncthreads();                             // also orders pass 1's global writes of F before the reads below (CTA scope)
  const float nu = p.nu, om = 1.0f - p.nu;
  constexpr int NQ = NCH * NHB / 2;            // float4 elements (two bins each)
  float4* A4 = reinterpret_cast<float4*>(A);
  const float4* F4 = reinterpret_cast<const float4*>(F);
  for (int q0 = tid; q0 < NQ; q0 += 4 * kThreads) {
    float4 a[4], f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * kThreads;
      if (q < NQ) {
        a[u] = A4[q];
        f[u] = F4[q];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * kThreads;
      if (q < NQ) {
        const int h = (2 * q) % NHB;           // NHB is even: a pair never straddles two channels
        const float2 g0 = gsm[h], g1 = gsm[h + 1];
        const float2 gf0 = cmul(g0, make_float2(f[u].x, f[u].y)), gf1 = cmul(g1, make_float2(f[u].z, f[u].w));
        a[u].x = om * a[u].x + nu * gf0.x;
        a[u].y = om * a[u].y + nu * gf0.y;
        a[u].z = om * a[u].z + nu * gf1.x;
        a[u].w = om * a[u].w + nu * gf1.y;
        A4[q] = a[u];
      }
    }
  }
}


// =============================================================================================
// scale filter (second half of dlib correlation_tracker::update): 32 scales alpha^(k-16) of the
// position rectangle -> 23x23 chips -> FHOG cell 4 (4x4x31 = 496 features) x Hann over scales ->
// length-32 FFT per feature -> response over scales -> interpolated argmax -> position *= factor
// -> running update of As[496][32], Bs[32].  One CTA per track; mirrors oracle/dsst.py.
// =============================================================================================
constexpr int NS = 32;
constexpr int SW = 23;
constexpr int SCELLS = 6;
constexpr int SOUT = 4;
constexpr int SF = 31 * SOUT * SOUT;   // 496
constexpr int ZP = NS + 1;             // padded row length (bank conflicts)
constexpr int kGroups = kThreads / 32; // row groups for the reductions over features

struct ScaleTables {
  float hann[NS];
  float factor[NS];          // alpha^(k-16) in float32, computed on the host like the oracle
  float tw_re[NS], tw_im[NS];  // exp(-2 pi i m / 32)
  float uu[9], vv[9];
  float lambda, nu;
  double alpha;
};

struct ScaleParams {
  float2* As;      // [cap][496][32]
  float* Bs;       // [cap][32]
  float* pos;      // [cap][4]
  const int* ids;
  const uint8_t* frame;
  const int* frame_idx;
  int H, W;
};

__device__ __forceinline__ int rev5(int v) { return (int)(__brev((unsigned)v) >> 27); }

constexpr int SG = 4;                  // scales built per round (32 scales = 8 rounds)
constexpr int SPX = SW * SW;           // 529 pixels per scale chip
constexpr int SROWS = SF / 2;          // two real features share one complex row: 248 rows
constexpr int NBIN = SCELLS * SCELLS * 18;

// Two changes against the first version (ncu: 1 CTA / SM at 131 KB, 224 block barriers around 529-pixel loops):
//  * the 32 scale chips are built SG = 4 at a time, so the seven barriers of a round are amortised over 2116 pixels;
//  * features 2r and 2r+1 (both real) are the real and imaginary part of ONE complex row, the 248 row transforms give
//    both spectra, Fa[k] = (Z[k] + conj(Z[-k])) / 2, Fb[k] = (Z[k] - conj(Z[-k])) / (2i): half the transforms and 66 KB
//    instead of 131 KB of shared memory, i.e. two tracks per SM.
template <bool START>
__global__ void __launch_bounds__(kThreads, 2) tracker_scale_kernel(ScaleParams p, const __grid_constant__ ScaleTables tbc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ ScaleTables tb;        // per-lane indexed tables: shared memory, not the constant bank (see tracker_kernel)
  for (int i = threadIdx.x; i < (int)(sizeof(ScaleTables) / 4); i += kThreads)
    reinterpret_cast<uint32_t*>(&tb)[i] = reinterpret_cast<const uint32_t*>(&tbc)[i];
  __syncthreads();
  float2* Z = reinterpret_cast<float2*>(smem_raw);                  // [SROWS][ZP]
  float* s_mag = reinterpret_cast<float*>(Z + SROWS * ZP);           // [SG][SPX]
  float* s_hist = s_mag + SG * SPX;                                  // [SG][36*18]
  unsigned* s_histi = reinterpret_cast<unsigned*>(s_hist + SG * NBIN);   // [SG][36*18] fixed-point scatter target
  float* s_nrm = reinterpret_cast<float*>(s_histi + SG * NBIN);      // [SG][36]
  float* s_red = s_nrm + SG * SCELLS * SCELLS;                       // [kGroups*64 + 64]
  uint8_t* s_chip = reinterpret_cast<uint8_t*>(s_red + kGroups * 64 + 64);  // [SG][SPX*3]
  uint8_t* s_ori = s_chip + SG * SPX * 3 + 4;                        // [SG][SPX]
  __shared__ double s_gre[NS], s_gim[NS];   // conj(FFT(target))
  __shared__ double s_rre[NS], s_rim[NS];
  __shared__ float s_resp[NS];
  __shared__ float s_peak;
  __shared__ int s_cwi[SW];                 // FHOG cell-4 bilinear weights of pixel coordinate c:
  __shared__ float s_cw0[SW], s_cw1[SW];    //   cp = (c + 0.5)/4 - 0.5, s_cwi = floor(cp), s_cw0 = cp - floor(cp), s_cw1 = 1 - s_cw0
  const int tid = threadIdx.x;
  if (tid < SW) {
    const float cp = __fsub_rn(__fdiv_rn(__fadd_rn((float)tid, 0.5f), 4.0f), 0.5f);
    const int icp = (int)floorf(cp);
    const float v0 = __fsub_rn(cp, (float)icp);
    s_cwi[tid] = icp;
    s_cw0[tid] = v0;
    s_cw1[tid] = __fsub_rn(1.0f, v0);
  }
  __syncthreads();
  const int slot = p.ids[blockIdx.x];
  if (p.frame_idx) p.frame += (size_t)p.frame_idx[blockIdx.x] * p.H * p.W * 3;
  float2* As = p.As + (size_t)slot * SF * NS;
  float* Bs = p.Bs + (size_t)slot * NS;
  const float l = p.pos[slot * 4 + 0], t = p.pos[slot * 4 + 1], r = p.pos[slot * 4 + 2], b = p.pos[slot * 4 + 3];
  const float cx = __fmul_rn(__fadd_rn(l, r), 0.5f), cy = __fmul_rn(__fadd_rn(t, b), 0.5f);
  const float hw0 = __fmul_rn(__fsub_rn(r, l), 0.5f), hh0 = __fmul_rn(__fsub_rn(b, t), 0.5f);

  for (int k0 = 0; k0 < NS; k0 += SG) {
    // ---- chips of scales k0 .. k0+SG-1 ----
    for (int i = tid; i < SG * SPX; i += kThreads) {
      const int g = i / SPX, pix = i - g * SPX;
      const int k = k0 + g;
      const float hw = __fmul_rn(hw0, tb.factor[k]), hh = __fmul_rn(hh0, tb.factor[k]);
      const float lk = __fsub_rn(cx, hw), rk = __fadd_rn(cx, hw), tk = __fsub_rn(cy, hh), bk = __fadd_rn(cy, hh);
      const float sx = __fdiv_rn(__fsub_rn(rk, lk), (float)(SW - 1)), sy = __fdiv_rn(__fsub_rn(bk, tk), (float)(SW - 1));
      const int y = pix / SW, x = pix - y * SW;
      const float fx = __fadd_rn(lk, __fmul_rn((float)x, sx)), fy = __fadd_rn(tk, __fmul_rn((float)y, sy));
      const int left = (int)floorf(fx), top = (int)floorf(fy);
      uint8_t o[3] = {0, 0, 0};
      if (fx >= 0.f && fx < (float)(p.W - 1) && fy >= 0.f && fy < (float)(p.H - 1)) {   // overflow-safe form of left >= 0 && left + 1 < W
        const float lr = __fsub_rn(fx, (float)left), tbv = __fsub_rn(fy, (float)top);
        const float omlr = __fsub_rn(1.0f, lr), omtb = __fsub_rn(1.0f, tbv);
        const uint8_t* ptl = p.frame + ((long long)top * p.W + left) * 3;
        const uint8_t* pbl = ptl + (long long)p.W * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float a = __fadd_rn(__fmul_rn(omlr, u8f(ptl[c])), __fmul_rn(lr, u8f(ptl[3 + c])));
          const float bb = __fadd_rn(__fmul_rn(omlr, u8f(pbl[c])), __fmul_rn(lr, u8f(pbl[3 + c])));
          float v = __fadd_rn(__fmul_rn(omtb, a), __fmul_rn(tbv, bb));
          o[c] = (uint8_t)fminf(fmaxf(floorf(__fadd_rn(v, 0.5f)), 0.f), 255.f);
        }
      }
      s_chip[3 * i] = o[0]; s_chip[3 * i + 1] = o[1]; s_chip[3 * i + 2] = o[2];
    }
    for (int bi = tid; bi < SG * NBIN; bi += kThreads) s_histi[bi] = 0u;
    __syncthreads();
    // ---- gradient magnitude / snapped orientation, scattered straight into the cell histograms ----
    // every pixel SCATTERS its magnitude into the (up to) 2 x 2 cells it overlaps, as 17-bit fixed point with
    // shared-memory integer atomics: integer addition is associative, so the result does not depend on the order of
    // arrival (deterministic), at a resolution of 2^-17 on sums < 2^15 — finer than float32 at these magnitudes.
    for (int i = tid; i < SG * SPX; i += kThreads) {
      const int g = i / SPX, pix = i - g * SPX;
      const int y = pix / SW, x = pix - y * SW;
      if (y < 1 || y > SW - 2 || x < 1 || x > SW - 2) continue;
      const uint8_t* ch = s_chip + 3 * (size_t)(g * SPX);
      float gx = 0.f, gy = 0.f, best = -1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float dx = __fsub_rn(u8f(ch[3 * (pix + 1) + c]), u8f(ch[3 * (pix - 1) + c]));
        const float dy = __fsub_rn(u8f(ch[3 * (pix + SW) + c]), u8f(ch[3 * (pix - SW) + c]));
        const float v = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (v > best) { best = v; gx = dx; gy = dy; }
      }
      const float m = sqrtf(best);
      float best_dot = 0.f;
      int o = 0;
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const float dot = __fadd_rn(__fmul_rn(tb.uu[q], gx), __fmul_rn(tb.vv[q], gy));
        if (dot > best_dot) { best_dot = dot; o = q; }
        else if (-dot > best_dot) { best_dot = -dot; o = q + 9; }
      }
      const int iyp = s_cwi[y], ixp = s_cwi[x];
      const float wy[2] = {s_cw1[y], s_cw0[y]}, wx[2] = {s_cw1[x], s_cw0[x]};   // cell iyp gets 1 - frac, cell iyp + 1 gets frac
      unsigned* hist = s_histi + g * NBIN;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int cyi = iyp + dy;
        if (cyi < 0 || cyi >= SCELLS) continue;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int cxi = ixp + dx;
          if (cxi < 0 || cxi >= SCELLS) continue;
          const float v = __fmul_rn(__fmul_rn(wx[dx], wy[dy]), m);
          atomicAdd(&hist[(cyi * SCELLS + cxi) * 18 + o], (unsigned)__float2uint_rn(__fmul_rn(v, 131072.0f)));
        }
      }
    }
    __syncthreads();
    for (int bi = tid; bi < SG * NBIN; bi += kThreads) s_hist[bi] = __fmul_rn((float)s_histi[bi], 1.0f / 131072.0f);   // [g][cell][o]
    __syncthreads();
    if (tid < SG * SCELLS * SCELLS) {
      const float* h = s_hist + tid * 18;      // (g, cell) flattened: g * 36 + cell
      float n = 0.f;
      for (int o = 0; o < 9; ++o) {
        const float sv = __fadd_rn(h[o], h[o + 9]);
        n = __fadd_rn(n, __fmul_rn(sv, sv));
      }
      s_nrm[tid] = n;
    }
    __syncthreads();
    // ---- 496 features of each of the SG scales ----
    for (int jj = tid; jj < SG * SF; jj += kThreads) {
      const int g = jj / SF, j = jj - g * SF;
      const int k = k0 + g;
      const int plane = j / (SOUT * SOUT), rem = j - plane * SOUT * SOUT;
      const int y = rem / SOUT, x = rem - y * SOUT;
      const int Y = y + 1, X = x + 1;
      const float* nrm = s_nrm + g * SCELLS * SCELLS;
      float ns[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int yy = Y - 1 + (q >> 1), xx = X - 1 + (q & 1);
        const float blk = __fadd_rn(__fadd_rn(__fadd_rn(nrm[yy * SCELLS + xx], nrm[yy * SCELLS + xx + 1]),
                                              nrm[(yy + 1) * SCELLS + xx]), nrm[(yy + 1) * SCELLS + xx + 1]);
        ns[q] = __fdiv_rn(1.0f, sqrtf(__fadd_rn(blk, 0.0001f)));
      }
      const float* h = s_hist + g * NBIN + (Y * SCELLS + X) * 18;
      float v;
      if (plane < 27) {
        const float hv = plane < 18 ? h[plane] : __fadd_rn(h[plane - 18], h[plane - 18 + 9]);
        const float h0 = fminf(__fmul_rn(hv, ns[0]), 0.2f), h1 = fminf(__fmul_rn(hv, ns[1]), 0.2f);
        const float h2 = fminf(__fmul_rn(hv, ns[2]), 0.2f), h3 = fminf(__fmul_rn(hv, ns[3]), 0.2f);
        v = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(h0, h1), h2), h3));
      } else {
        const int q = plane - 27;
        float tsum = 0.f;
        for (int o = 0; o < 18; ++o) tsum = __fadd_rn(tsum, fminf(__fmul_rn(h[o], ns[q]), 0.2f));
        v = __fmul_rn(0.2357f, tsum);
      }
      // feature j -> row j >> 1, real part for even j, imaginary part for odd j; stored bit-reversed over scales
      float* zz = reinterpret_cast<float*>(Z + (j >> 1) * ZP + rev5(k));
      zz[j & 1] = __fmul_rn(v, tb.hann[k]);
    }
    __syncthreads();
  }

  // ---- length-32 FFT over scales for every row (input stored bit-reversed) ----
  for (int j = tid; j < SROWS; j += kThreads) {
    float2* z = Z + j * ZP;
    for (int s = 1; s <= 5; ++s) {
      const int half = 1 << (s - 1);
      for (int bfly = 0; bfly < 16; ++bfly) {
        const int grp = bfly / half, q = bfly - grp * half;
        const int i0 = grp * (half << 1) + q, i1 = i0 + half;
        const int m = q * (16 / half);
        const float2 w = make_float2(tb.tw_re[m], tb.tw_im[m]);
        const float2 u = z[i0], tv = cmul(w, z[i1]);
        z[i0] = make_float2(u.x + tv.x, u.y + tv.y);
        z[i1] = make_float2(u.x - tv.x, u.y - tv.y);
      }
    }
  }
  __syncthreads();

  const int k = tid & 31, grp = tid >> 5;   // kGroups row groups
  const int km = (NS - k) & (NS - 1);
  float peak = 0.5f * NS;
  if (!START) {
    // response R^[k] = sum_j F_j[k] conj(As[j][k]) / (Bs[k] + lambda)
    float re = 0.f, im = 0.f;
    for (int rr = grp; rr < SROWS; rr += kGroups) {
      const float2 z = Z[rr * ZP + k], zm = Z[rr * ZP + km];
      const float2 fa = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
      const float2 fb = make_float2(0.5f * (z.y + zm.y), 0.5f * (zm.x - z.x));
      const float2 a = As[(size_t)(2 * rr) * NS + k], a2 = As[(size_t)(2 * rr + 1) * NS + k];
      re += fa.x * a.x + fa.y * a.y;
      im += fa.y * a.x - fa.x * a.y;
      re += fb.x * a2.x + fb.y * a2.y;
      im += fb.y * a2.x - fb.x * a2.y;
    }
    s_red[grp * 64 + k] = re;
    s_red[grp * 64 + 32 + k] = im;
    __syncthreads();
    if (tid < NS) {
      double sr = 0, si = 0;
      for (int g = 0; g < kGroups; ++g) { sr += s_red[g * 64 + tid]; si += s_red[g * 64 + 32 + tid]; }
      const double d = 1.0 / ((double)Bs[tid] + (double)tb.lambda);
      s_rre[tid] = sr * d;
      s_rim[tid] = si * d;
    }
    __syncthreads();
    if (tid < NS) {   // inverse DFT, real part
      double acc = 0;
      for (int q = 0; q < NS; ++q) {
        const int m = (q * tid) & 31;
        acc += s_rre[q] * (double)tb.tw_re[m] + s_rim[q] * (double)tb.tw_im[m];   // Re(R * conj(w)) = Re(R e^{+i...})
      }
      s_resp[tid] = (float)(acc / NS);
    }
    __syncthreads();
    if (tid == 0) {
      int pk = 0;
      float best = s_resp[0];
      for (int q = 1; q < NS; ++q) if (s_resp[q] > best) { best = s_resp[q]; pk = q; }
      double pp = pk;
      if (pk > 0 && pk < NS - 1) {
        const double c = s_resp[pk], a = s_resp[pk - 1], d2 = s_resp[pk + 1];
        const double den = a - 2 * c + d2;
        // dlib's 1-D max_point_interpolated (lagrange_poly_min_extrap) stays inside [pk - 1, pk + 1]
        if (den != 0) pp += fmin(1.0, fmax(-1.0, 0.5 * (a - d2) / den));
      }
      const double f = pow(tb.alpha, pp - NS / 2);
      const double dl = l, dt = t, dr = r, db = b;
      const double ccx = 0.5 * (dl + dr), ccy = 0.5 * (dt + db);
      const double nhw = 0.5 * (dr - dl) * f, nhh = 0.5 * (db - dt) * f;
      p.pos[slot * 4 + 0] = (float)(ccx - nhw);
      p.pos[slot * 4 + 1] = (float)(ccy - nhh);
      p.pos[slot * 4 + 2] = (float)(ccx + nhw);
      p.pos[slot * 4 + 3] = (float)(ccy + nhh);
      s_peak = (float)pp;
    }
    __syncthreads();
    peak = s_peak;
  }
  // conj(DFT(target)) at `peak`
  if (tid < NS) {
    double re = 0, im = 0;
    for (int q = 0; q < NS; ++q) {
      const float dq = (float)q - peak;
      const double g = (double)expf(-(dq * dq) / 1.0f);
      const int m = (q * tid) & 31;
      re += g * (double)tb.tw_re[m];
      im += g * (double)tb.tw_im[m];
    }
    s_gre[tid] = re;
    s_gim[tid] = -im;
  }
  __syncthreads();
  // filter update / initialisation
  {
    const float gre = (float)s_gre[k], gim = (float)s_gim[k];
    float bsum = 0.f;
    for (int rr = grp; rr < SROWS; rr += kGroups) {
      const float2 z = Z[rr * ZP + k], zm = Z[rr * ZP + km];
      float2 f[2];
      f[0] = make_float2(0.5f * (z.x + zm.x), 0.5f * (z.y - zm.y));
      f[1] = make_float2(0.5f * (z.y + zm.y), 0.5f * (zm.x - z.x));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 gf = make_float2(gre * f[h].x - gim * f[h].y, gre * f[h].y + gim * f[h].x);
        bsum += f[h].x * f[h].x + f[h].y * f[h].y;
        float2* ap = As + (size_t)(2 * rr + h) * NS + k;
        if (START) {
          *ap = gf;
        } else {
          float2 a = *ap;
          a.x = (1.0f - tb.nu) * a.x + tb.nu * gf.x;
          a.y = (1.0f - tb.nu) * a.y + tb.nu * gf.y;
          *ap = a;
        }
      }
    }
    __syncthreads();
    s_red[grp * 64 + k] = bsum;
    __syncthreads();
    if (tid < NS) {
      float sb = 0.f;
      for (int g = 0; g < kGroups; ++g) sb += s_red[g * 64 + tid];
      Bs[tid] = START ? sb : (1.0f - tb.nu) * Bs[tid] + tb.nu * sb;
    }
  }
}
constexpr size_t kScaleSmem = (size_t)SROWS * ZP * 8 + (size_t)SG * SPX * 4 + 2 * (size_t)SG * NBIN * 4 + SG * SCELLS * SCELLS * 4 +
                              (kGroups * 64 + 64) * 4 + SG * SPX * 3 + 4 + SG * SPX + 64;
static_assert(2 * (kScaleSmem + 4096) <= 227 * 1024, "two CTAs per SM");

struct Bank {
  int capacity;
  float2* As = nullptr;
  float* Bs = nullptr;
  ScaleTables stb;
  bool has_scale = false;
  float2* A;
  float2* F = nullptr;
  float* B;
  float* pos;
  float* psr;
  TrackerTables tb;
  float padding, lambda, nu;
};

}  // namespace

extern "C" int pv_tracker_create(int capacity, const float* hann64_host, const float* uu9_host, const float* vv9_host,
                                 const float* tw_re32_host, const float* tw_im32_host, float padding, float lambda,
                                 float nu, void** out_handle) {
  PV_REQUIRE(capacity > 0 && hann64_host && uu9_host && vv9_host && tw_re32_host && tw_im32_host && out_handle,
             "pv_tracker_create: bad argument");
  Bank* b = new Bank();
  b->capacity = capacity;
  memcpy(b->tb.hann, hann64_host, sizeof(float) * FS);
  memcpy(b->tb.uu, uu9_host, sizeof(float) * 9);
  memcpy(b->tb.vv, vv9_host, sizeof(float) * 9);
  memcpy(b->tb.tw_re, tw_re32_host, sizeof(float) * 32);
  memcpy(b->tb.tw_im, tw_im32_host, sizeof(float) * 32);
  b->padding = padding;
  b->lambda = lambda;
  b->nu = nu;
  cudaError_t e = cudaMalloc(&b->A, sizeof(float2) * (size_t)capacity * NCH * NHB);
  if (e == cudaSuccess) e = cudaMalloc(&b->F, sizeof(float2) * (size_t)capacity * NCH * NHB);
  if (e == cudaSuccess) e = cudaMalloc(&b->B, sizeof(float) * (size_t)capacity * NHB);
  if (e == cudaSuccess) e = cudaMalloc(&b->pos, sizeof(float) * (size_t)capacity * 4);
  if (e == cudaSuccess) e = cudaMalloc(&b->psr, sizeof(float) * (size_t)capacity);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tracker_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tracker_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
  if (e != cudaSuccess) {
    pv_set_error("pv_tracker_create: %s", cudaGetErrorString(e));
    delete b;
    return PV_ERR_CUDA;
  }
  *out_handle = b;
  return PV_OK;
}

extern "C" int pv_tracker_destroy(void* handle) {
  if (!handle) return PV_OK;
  Bank* b = static_cast<Bank*>(handle);
  cudaFree(b->A);
  cudaFree(b->F);
  cudaFree(b->B);
  cudaFree(b->pos);
  cudaFree(b->psr);
  cudaFree(b->As);
  cudaFree(b->Bs);
  delete b;
  return PV_OK;
}

static int launch(Bank* b, bool start, const void* frame, const int* frame_idx, int H, int W, const int* ids,
                  const float* rects, int n, void* stream) {
  if (n == 0) return PV_OK;
  BankParams p;
  p.A = b->A;
  p.F = b->F;
  p.B = b->B;
  p.pos = b->pos;
  p.psr = b->psr;
  p.ids = ids;
  p.rects = rects;
  p.frame = static_cast<const uint8_t*>(frame);
  p.frame_idx = frame_idx;
  p.H = H;
  p.W = W;
  p.padding = b->padding;
  p.lambda = b->lambda;
  p.nu = b->nu;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (start)
    tracker_kernel<true><<<n, kThreads, kSmemBytes, s>>>(p, b->tb);
  else
    tracker_kernel<false><<<n, kThreads, kSmemBytes, s>>>(p, b->tb);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

/* enable the scale filter: tables are HOST float arrays of 32 entries (Hann over scales, alpha^(k-16), DFT twiddles) */
extern "C" int pv_tracker_enable_scale(void* handle, const float* hann32_host, const float* factor32_host,
                                       const float* tw_re32_host, const float* tw_im32_host, double alpha, float lambda,
                                       float nu) {
  PV_REQUIRE(handle && hann32_host && factor32_host && tw_re32_host && tw_im32_host, "pv_tracker_enable_scale: null argument");
  Bank* b = static_cast<Bank*>(handle);
  memcpy(b->stb.hann, hann32_host, sizeof(float) * NS);
  memcpy(b->stb.factor, factor32_host, sizeof(float) * NS);
  memcpy(b->stb.tw_re, tw_re32_host, sizeof(float) * NS);
  memcpy(b->stb.tw_im, tw_im32_host, sizeof(float) * NS);
  memcpy(b->stb.uu, b->tb.uu, sizeof(float) * 9);
  memcpy(b->stb.vv, b->tb.vv, sizeof(float) * 9);
  b->stb.alpha = alpha;
  b->stb.lambda = lambda;
  b->stb.nu = nu;
  if (!b->As) {
    cudaError_t e = cudaMalloc(&b->As, sizeof(float2) * (size_t)b->capacity * SF * NS);
    if (e == cudaSuccess) e = cudaMalloc(&b->Bs, sizeof(float) * (size_t)b->capacity * NS);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tracker_scale_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScaleSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tracker_scale_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScaleSmem);
    if (e != cudaSuccess) {
      pv_set_error("pv_tracker_enable_scale: %s", cudaGetErrorString(e));
      return PV_ERR_CUDA;
    }
  }
  b->has_scale = true;
  return PV_OK;
}

static int launch_scale(Bank* b, bool start, const void* frame, const int* frame_idx, int H, int W, const int* ids, int n,
                        void* stream) {
  if (n == 0 || !b->has_scale) return PV_OK;
  ScaleParams p;
  p.As = b->As;
  p.Bs = b->Bs;
  p.pos = b->pos;
  p.ids = ids;
  p.frame = static_cast<const uint8_t*>(frame);
  p.frame_idx = frame_idx;
  p.H = H;
  p.W = W;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (start)
    tracker_scale_kernel<true><<<n, kThreads, kScaleSmem, s>>>(p, b->stb);
  else
    tracker_scale_kernel<false><<<n, kThreads, kScaleSmem, s>>>(p, b->stb);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

/* start trackers in slots ids[0..n) on `frame` (uint8 [H,W,3]) at rects [n,4] (l,t,r,b floats) */
extern "C" int pv_tracker_start(void* handle, const void* frame, int H, int W, const int* ids, const float* rects, int n,
                                void* stream) {
  return pv_tracker_start_frames(handle, frame, H, W, nullptr, ids, rects, n, stream);
}

/* advance the trackers in slots ids[0..n) to `frame`; PSR and positions are left in the bank */
extern "C" int pv_tracker_update(void* handle, const void* frame, int H, int W, const int* ids, int n, void* stream) {
  return pv_tracker_update_frames(handle, frame, H, W, nullptr, ids, n, stream);
}

/* batched forms: frames uint8 [F,H,W,3], frame_idx[i] = frame of track ids[i] (tracks of many shots / videos advance
 * in one launch; tracking state is independent per shot, pyannote/video/tracking.py:410-417) */
extern "C" int pv_tracker_start_frames(void* handle, const void* frames, int H, int W, const int* frame_idx, const int* ids,
                                       const float* rects, int n, void* stream) {
  PV_REQUIRE(handle && frames && ids && rects, "pv_tracker_start: null argument");
  int rc = launch(static_cast<Bank*>(handle), true, frames, frame_idx, H, W, ids, rects, n, stream);
  if (rc != PV_OK) return rc;
  return launch_scale(static_cast<Bank*>(handle), true, frames, frame_idx, H, W, ids, n, stream);
}

extern "C" int pv_tracker_update_frames(void* handle, const void* frames, int H, int W, const int* frame_idx, const int* ids,
                                        int n, void* stream) {
  PV_REQUIRE(handle && frames && ids, "pv_tracker_update: null argument");
  int rc = launch(static_cast<Bank*>(handle), false, frames, frame_idx, H, W, ids, nullptr, n, stream);
  if (rc != PV_OK) return rc;
  return launch_scale(static_cast<Bank*>(handle), false, frames, frame_idx, H, W, ids, n, stream);
}

/* device pointers to the bank's state: positions float [capacity,4], psr float [capacity] */
extern "C" int pv_tracker_state(void* handle, float** pos, float** psr) {
  PV_REQUIRE(handle && pos && psr, "pv_tracker_state: null argument");
  Bank* b = static_cast<Bank*>(handle);
  *pos = b->pos;
  *psr = b->psr;
  return PV_OK;
}
