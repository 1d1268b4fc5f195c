// tracker.cu — bank of DSST correlation trackers, one CTA per live track, all tracks of a frame in
// ONE launch.  Replaces dlib.correlation_tracker.start_track / update / get_position called per
// tracker per frame from Python (pyannote/video/tracking.py:203,231,250-251).
//
// Per track state in HBM (float32): A[31][64x64] complex numerators, B[64x64] denominator,
// position (l,t,r,b).  update = chip (bilinear, rect*1.4 -> 64x64) -> FHOG-31 (cell 1) x cosine
// window -> 31 2-D FFTs (shared memory, radix-2) -> response = ifft2(sum F_i conj(A_i)/(B+lambda))
// -> argmax / sub-pixel / PSR (warp-shuffle + shared reductions) -> position -> filter update
// (features are recomputed rather than spilled: 1 MB of state is read twice and written once).
// Mirrors oracle/dsst.py step by step; scale filter not implemented (stated gap).
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int FS = 64;
constexpr int NPIX = FS * FS;
constexpr int NCH = 31;
constexpr int kThreads = 256;

struct TrackerTables {
  float hann[FS];
  float uu[9], vv[9];
  float tw_re[32], tw_im[32];  // exp(-2 pi i k / 64)
};

struct BankParams {
  float2* A;        // [cap][31][4096]
  float* B;         // [cap][4096]
  float* pos;       // [cap][4]
  float* psr;       // [cap]
  const int* ids;   // [n] slots handled by this launch
  const float* rects;  // start only: [n][4]
  const uint8_t* frame;  // uint8 [H,W,3]
  int H, W;
  float padding, lambda, nu;
};

__device__ __forceinline__ int rev6(int v) { return (int)(__brev((unsigned)v) >> 26); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// in-place 2-D radix-2 DIT FFT of a 64x64 complex tile whose input was stored bit-reversed in both
// dimensions; output in natural order.  inverse: conjugate twiddles (no scaling).
__device__ void fft2_64(float2* x, const TrackerTables& tb, bool inverse) {
  const int tid = threadIdx.x;
  for (int dim = 0; dim < 2; ++dim) {
    const int es = dim == 0 ? 1 : FS;   // element stride along the transformed dimension
    const int ls = dim == 0 ? FS : 1;   // stride between lines
    for (int s = 1; s <= 6; ++s) {
      const int half = 1 << (s - 1);
      for (int b = tid; b < FS * 32; b += kThreads) {
        const int line = b >> 5, jj = b & 31;
        const int grp = jj / half, j = jj - grp * half;
        const int i0 = grp * (half << 1) + j, i1 = i0 + half;
        const int k = j * (32 / half);
        float2 w = make_float2(tb.tw_re[k], inverse ? -tb.tw_im[k] : tb.tw_im[k]);
        float2* p0 = x + line * ls + i0 * es;
        float2* p1 = x + line * ls + i1 * es;
        const float2 u = *p0, t = cmul(w, *p1);
        *p0 = make_float2(u.x + t.x, u.y + t.y);
        *p1 = make_float2(u.x - t.x, u.y - t.y);
      }
      __syncthreads();
    }
  }
}

struct Smem {
  uint8_t* chip;   // [4096*3]
  uint8_t* ori;    // [4096]
  float* mag;      // [4096]
  float* osum;     // [4096]
  float* bsum;     // [4096]
  float2* plane;   // [4096]
  float2* acc;     // [4096]
  float* red;      // [64]
  int* redi;       // [32]
};

__device__ Smem carve(uint8_t* base) {
  Smem s;
  s.plane = reinterpret_cast<float2*>(base);
  s.acc = s.plane + NPIX;
  s.mag = reinterpret_cast<float*>(s.acc + NPIX);
  s.osum = s.mag + NPIX;
  s.bsum = s.osum + NPIX;
  s.red = s.bsum + NPIX;
  s.redi = reinterpret_cast<int*>(s.red + 64);
  s.chip = reinterpret_cast<uint8_t*>(s.redi + 32);
  s.ori = s.chip + NPIX * 3;
  return s;
}
constexpr size_t kSmemBytes = 2 * NPIX * 8 + 3 * NPIX * 4 + 64 * 4 + 32 * 4 + NPIX * 3 + NPIX;

// chip + gradient orientation/magnitude + per-cell orientation feature value
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
    if (left >= 0 && left + 1 < p.W && top >= 0 && top + 1 < p.H) {
      const float lr = __fsub_rn(fx, (float)left), tbv = __fsub_rn(fy, (float)top);
      const float omlr = __fsub_rn(1.0f, lr), omtb = __fsub_rn(1.0f, tbv);
      const uint8_t* ptl = p.frame + ((long long)top * p.W + left) * 3;
      const uint8_t* pbl = ptl + (long long)p.W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = __fadd_rn(__fmul_rn(omlr, (float)ptl[c]), __fmul_rn(lr, (float)ptl[3 + c]));
        const float bb = __fadd_rn(__fmul_rn(omlr, (float)pbl[c]), __fmul_rn(lr, (float)pbl[3 + c]));
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
        const float dx = __fsub_rn((float)s.chip[3 * (i + 1) + c], (float)s.chip[3 * (i - 1) + c]);
        const float dy = __fsub_rn((float)s.chip[3 * (i + FS) + c], (float)s.chip[3 * (i - FS) + c]);
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
  }
  __syncthreads();
}

__device__ __forceinline__ float nrm_at(const Smem& s, int y, int x) {
  if (y < 0 || y >= FS || x < 0 || x >= FS) return 0.f;
  const float m = s.mag[y * FS + x];
  return __fmul_rn(m, m);
}
__device__ __forceinline__ float inv_block(const Smem& s, int y0, int x0) {
  const float a = nrm_at(s, y0, x0), b = nrm_at(s, y0, x0 + 1), c = nrm_at(s, y0 + 1, x0), d = nrm_at(s, y0 + 1, x0 + 1);
  const float sum = __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d);
  return __fdiv_rn(1.0f, sqrtf(__fadd_rn(sum, 0.0001f)));
}
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

// windowed feature plane `ch`, stored bit-reversed for the FFT
__device__ void build_plane(const Smem& s, const TrackerTables& tb, int ch) {
  for (int i = threadIdx.x; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    float v = 0.f;
    if (y > 0 && y < FS - 1 && x > 0 && x < FS - 1) {
      const int o = s.ori[i];
      if (ch < 18) v = (o == ch) ? s.osum[i] : 0.f;
      else if (ch < 27) v = ((o % 9) == ch - 18) ? s.osum[i] : 0.f;
      else {
        const int k = ch - 27;
        const float nk = inv_block(s, y - 1 + (k >> 1), x - 1 + (k & 1));
        v = __fmul_rn(0.2357f, fminf(__fmul_rn(s.mag[i], nk), 0.2f));
      }
      v = __fmul_rn(v, __fmul_rn(tb.hann[y], tb.hann[x]));
    }
    s.plane[rev6(y) * FS + rev6(x)] = make_float2(v, 0.f);
  }
  __syncthreads();
}

// FFT of the Gaussian target centred at (px,py); leaves conj(G^) in s.acc
__device__ void target_hat(Smem& s, const TrackerTables& tb, float px, float py) {
  for (int i = threadIdx.x; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    const float dx = (float)x - px, dy = (float)y - py;
    s.plane[rev6(y) * FS + rev6(x)] = make_float2(expf(-(dx * dx + dy * dy) / 3.0f), 0.f);
  }
  __syncthreads();
  fft2_64(s.plane, tb, false);
  for (int i = threadIdx.x; i < NPIX; i += kThreads) s.acc[i] = make_float2(s.plane[i].x, -s.plane[i].y);
  __syncthreads();
}

template <bool START>
__global__ void __launch_bounds__(kThreads) tracker_kernel(BankParams p, const __grid_constant__ TrackerTables tb) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem s = carve(smem_raw);
  __shared__ float tf[4];
  __shared__ float peak[4];  // ppx, ppy
  const int tid = threadIdx.x;
  const int slot = p.ids[blockIdx.x];
  float2* A = p.A + (size_t)slot * NCH * NPIX;
  float* B = p.B + (size_t)slot * NPIX;
  float rect[4];
  if (START) {
#pragma unroll
    for (int k = 0; k < 4; ++k) rect[k] = p.rects[blockIdx.x * 4 + k];
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) rect[k] = p.pos[slot * 4 + k];
  }
  features_prepare(p, tb, rect, s, tf);
  features_osum(s);

  if (START) {
    target_hat(s, tb, 0.5f * (FS - 1), 0.5f * (FS - 1));
    for (int i = tid; i < NPIX; i += kThreads) s.bsum[i] = 0.f;
    __syncthreads();
    for (int ch = 0; ch < NCH; ++ch) {
      build_plane(s, tb, ch);
      fft2_64(s.plane, tb, false);
      for (int i = tid; i < NPIX; i += kThreads) {
        const float2 f = s.plane[i];
        A[(size_t)ch * NPIX + i] = cmul(s.acc[i], f);
        s.bsum[i] += f.x * f.x + f.y * f.y;
      }
      __syncthreads();
    }
    for (int i = tid; i < NPIX; i += kThreads) B[i] = s.bsum[i];
    if (tid < 4) p.pos[slot * 4 + tid] = rect[tid];
    if (tid == 0) p.psr[slot] = 0.f;
    return;
  }

  // ---- pass 1: response ----
  for (int i = tid; i < NPIX; i += kThreads) {
    s.acc[i] = make_float2(0.f, 0.f);
    s.bsum[i] = 0.f;
  }
  __syncthreads();
  for (int ch = 0; ch < NCH; ++ch) {
    build_plane(s, tb, ch);
    fft2_64(s.plane, tb, false);
    for (int i = tid; i < NPIX; i += kThreads) {
      const float2 f = s.plane[i];
      const float2 a = A[(size_t)ch * NPIX + i];
      float2 acc = s.acc[i];
      acc.x += f.x * a.x + f.y * a.y;   // f * conj(a)
      acc.y += f.y * a.x - f.x * a.y;
      s.acc[i] = acc;
      s.bsum[i] += f.x * f.x + f.y * f.y;
    }
    __syncthreads();
  }
  for (int i = tid; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    const float d = 1.0f / (B[i] + p.lambda);
    s.plane[rev6(y) * FS + rev6(x)] = make_float2(s.acc[i].x * d, s.acc[i].y * d);
  }
  __syncthreads();
  fft2_64(s.plane, tb, true);
  // real response (scaled by 1/4096) kept in s.osum (features are rebuilt in pass 2)
  float bestv = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = tid; i < NPIX; i += kThreads) {
    const float v = s.plane[i].x * (1.0f / NPIX);
    s.acc[i].x = v;   // stash R in acc.x until the target is built
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
  for (int i = tid; i < NPIX; i += kThreads) {
    const int y = i >> 6, x = i & 63;
    if (y >= py - 4 && y < py + 4 && x >= px - 4 && x < px + 4) continue;
    const double v = (double)s.acc[i].x;
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
  __shared__ double dsum[8], dsq[8];
  __shared__ int dcnt[8];
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
      const double c = s.acc[pi].x, xl = s.acc[pi - 1].x, xr = s.acc[pi + 1].x, yu = s.acc[pi - FS].x, yd = s.acc[pi + FS].x;
      const double dxx = xl - 2 * c + xr, dyy = yu - 2 * c + yd;
      if (dxx != 0) ppx += 0.5 * (xl - xr) / dxx;
      if (dyy != 0) ppy += 0.5 * (yu - yd) / dyy;
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

  // ---- pass 2: filter update ----
  target_hat(s, tb, peak[0], peak[1]);
  const float nu = p.nu, om = 1.0f - p.nu;
  for (int ch = 0; ch < NCH; ++ch) {
    build_plane(s, tb, ch);
    fft2_64(s.plane, tb, false);
    for (int i = tid; i < NPIX; i += kThreads) {
      const float2 gf = cmul(s.acc[i], s.plane[i]);
      float2 a = A[(size_t)ch * NPIX + i];
      a.x = om * a.x + nu * gf.x;
      a.y = om * a.y + nu * gf.y;
      A[(size_t)ch * NPIX + i] = a;
    }
    __syncthreads();
  }
  for (int i = tid; i < NPIX; i += kThreads) B[i] = om * B[i] + nu * s.bsum[i];
}

struct Bank {
  int capacity;
  float2* A;
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
  cudaError_t e = cudaMalloc(&b->A, sizeof(float2) * (size_t)capacity * NCH * NPIX);
  if (e == cudaSuccess) e = cudaMalloc(&b->B, sizeof(float) * (size_t)capacity * NPIX);
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
  cudaFree(b->B);
  cudaFree(b->pos);
  cudaFree(b->psr);
  delete b;
  return PV_OK;
}

static int launch(Bank* b, bool start, const void* frame, int H, int W, const int* ids, const float* rects, int n,
                  void* stream) {
  if (n == 0) return PV_OK;
  BankParams p;
  p.A = b->A;
  p.B = b->B;
  p.pos = b->pos;
  p.psr = b->psr;
  p.ids = ids;
  p.rects = rects;
  p.frame = static_cast<const uint8_t*>(frame);
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

/* start trackers in slots ids[0..n) on `frame` (uint8 [H,W,3]) at rects [n,4] (l,t,r,b floats) */
extern "C" int pv_tracker_start(void* handle, const void* frame, int H, int W, const int* ids, const float* rects, int n,
                                void* stream) {
  PV_REQUIRE(handle && frame && ids && rects, "pv_tracker_start: null argument");
  return launch(static_cast<Bank*>(handle), true, frame, H, W, ids, rects, n, stream);
}

/* advance the trackers in slots ids[0..n) to `frame`; PSR and positions are left in the bank */
extern "C" int pv_tracker_update(void* handle, const void* frame, int H, int W, const int* ids, int n, void* stream) {
  PV_REQUIRE(handle && frame && ids, "pv_tracker_update: null argument");
  return launch(static_cast<Bank*>(handle), false, frame, H, W, ids, nullptr, n, stream);
}

/* device pointers to the bank's state: positions float [capacity,4], psr float [capacity] */
extern "C" int pv_tracker_state(void* handle, float** pos, float** psr) {
  PV_REQUIRE(handle && pos && psr, "pv_tracker_state: null argument");
  Bank* b = static_cast<Bank*>(handle);
  *pos = b->pos;
  *psr = b->psr;
  return PV_OK;
}
