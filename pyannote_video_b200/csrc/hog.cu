// hog.cu — dlib's HOG frontal face detector, the detector the reference really calls:
// dlib.get_frontal_face_detector()(rgb, 1), pyannote/video/face/face.py:54,66
// (object_detector<scan_fhog_pyramid<pyramid_down<6>>>).  Mirrors oracle/hog.py step by step:
//
//   hog_grad   per level pixel: colour channel with the largest gradient, orientation snapped to 18 bins, magnitude
//   hog_hist   one WARP per 8x8 cell, lane = orientation bin: the (at most 16 x 16) pixels that vote for the cell are
//              walked in raster order, so every (cell, bin) sum is accumulated in exactly the oracle's order
//              (deterministic and bit-exact: no atomics); lane 0 also leaves the cell's block-norm energy
//   hog_feat   31 Felzenszwalb features of every interior cell -> bf16, 32 channels, written into a FEATURE PLANE that
//              tiles the levels' feature maps with zero gaps (the sliding window of one level never sees another's cells)
//   (scores)   the 10 x 10 x 31 linear filters are one tcgen05 convolution over that plane: csrc/rsconv.cu,
//              instance <32, 16, 10, 10, 1> (up to 16 filters side by side as output channels, fp32 out)
//   hog_cand   cells whose score reaches the filter's threshold -> compacted (score, code) list per frame
//   hog_nms    one CTA per frame: total order (score desc, level, filter, row, col), boxes through dlib's
//              fhog_to_image / pyramid_down<6>::rect_up / pyramid_down<2>::rect_down, greedy NMS (test_box_overlap)
//
// The image pyramid itself is the tiled plane the CNN detector already builds (pv_resize_bilinear / pyramid_tail).
// All float arithmetic uses explicitly rounded, unfused operations in the oracle's order.
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kCell = 8;
constexpr int kFilt = 10;
constexpr int kPadOff = (kFilt - 1) / 2;     // 4
constexpr int kDetBox = kFilt - 2;           // 8 cells (padding 1 on each side left out)
constexpr int kNmsCap = 4096;

__device__ __forceinline__ float u8f(uint32_t b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }

__device__ __forceinline__ int level_of(const PvHogGeo& g, int idx, int which) {
  // which: 0 = pixel prefix, 1 = cell prefix, 2 = interior-cell prefix
  int lv = 0;
  for (int l = 1; l < g.n_levels; ++l) {
    const int off = which == 0 ? g.lv[l].px_off : (which == 1 ? g.lv[l].cell_off : g.lv[l].feat_off);
    if (idx >= off) lv = l;
  }
  return lv;
}

__global__ void __launch_bounds__(256) hog_grad_kernel(const uint32_t* __restrict__ plane, const __grid_constant__ PvHogGeo g,
                                                       uint8_t* __restrict__ ori, float* __restrict__ mag, const float* __restrict__ uv) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= g.total_px) return;
  const int b = blockIdx.y;
  const int l = level_of(g, idx, 0);
  const PvHogLevel L = g.lv[l];
  const int r = idx - L.px_off;
  const int y = r / L.w, x = r - y * L.w;
  const long long base = (long long)b * g.Hp * g.Wp + (long long)(L.y0 + y) * g.Wp + L.x0 + x;
  float m = 0.f;
  int bo = 0;
  if (y > 0 && y < L.h - 1 && x > 0 && x < L.w - 1) {
    const uint32_t pl = plane[base - 1], pr = plane[base + 1], pu = plane[base - g.Wp], pd = plane[base + g.Wp];
    float gx = 0.f, gy = 0.f, best = -1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float dx = __fsub_rn(u8f((pr >> (8 * c)) & 255u), u8f((pl >> (8 * c)) & 255u));
      const float dy = __fsub_rn(u8f((pd >> (8 * c)) & 255u), u8f((pu >> (8 * c)) & 255u));
      const float v = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      if (v > best) { best = v; gx = dx; gy = dy; }
    }
    m = __fsqrt_rn(best);
    float best_dot = 0.f;
#pragma unroll
    for (int o = 0; o < 9; ++o) {
      const float dot = __fadd_rn(__fmul_rn(uv[o], gx), __fmul_rn(uv[9 + o], gy));
      if (dot > best_dot) { best_dot = dot; bo = o; }
      else if (-dot > best_dot) { best_dot = -dot; bo = o + 9; }
    }
  }
  ori[base] = (uint8_t)bo;
  mag[base] = m;
}

// bilinear vote weight of pixel coordinate v for cell index c along one axis (0 when the pixel does not vote for it)
__device__ __forceinline__ float vote(int v, int c) {
  const float p = __fsub_rn(__fmul_rn(__fadd_rn((float)v, 0.5f), 0.125f), 0.5f);   // (v + 0.5) / 8 - 0.5
  const float fl = floorf(p);
  const int ip = (int)fl;
  const float v0 = __fsub_rn(p, fl), v1 = __fsub_rn(1.0f, v0);
  return ip == c ? v1 : (ip == c - 1 ? v0 : 0.f);
}

__global__ void __launch_bounds__(256) hog_hist_kernel(const uint8_t* __restrict__ ori, const float* __restrict__ mag,
                                                       const __grid_constant__ PvHogGeo g, float* __restrict__ hist,
                                                       float* __restrict__ nrm) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= g.total_cells) return;
  const int b = blockIdx.y;
  const int l = level_of(g, warp, 1);
  const PvHogLevel L = g.lv[l];
  const int r = warp - L.cell_off;
  const int cy = r / L.cx, cx = r - cy * L.cx;
  const int ya = max(kCell * cy - 4, 1), yb = min(kCell * cy + 11, L.h - 2);
  const int xa = max(kCell * cx - 4, 1), xb = min(kCell * cx + 11, L.w - 2);
  const long long base = (long long)b * g.Hp * g.Wp + (long long)L.y0 * g.Wp + L.x0;
  float acc = 0.f;
  // lane j < 16 owns column xa + j of the footprint: its horizontal weight is fixed, the row's data arrive with one
  // coalesced load and travel to every lane by shuffle
  const int xl = xa + (lane & 15);
  const float wx_l = xl <= xb ? vote(xl, cx) : 0.f;
  for (int y = ya; y <= yb; ++y) {
    const float wy = vote(y, cy);
    const long long row = base + (long long)y * g.Wp;
    int o_l = 0;
    float m_l = 0.f;
    if ((lane & 16) == 0 && xl <= xb) {
      o_l = ori[row + xl];
      m_l = mag[row + xl];
    }
    const int nx = xb - xa + 1;
    for (int j = 0; j < nx; ++j) {
      const int o = __shfl_sync(0xffffffffu, o_l, j);
      const float m = __shfl_sync(0xffffffffu, m_l, j);
      const float wx = __shfl_sync(0xffffffffu, wx_l, j);
      const float w = __fmul_rn(__fmul_rn(wx, wy), m);
      if (o == lane) acc = __fadd_rn(acc, w);
    }
  }
  const long long cell = (long long)b * g.total_cells + warp;
  if (lane < 18) hist[cell * 18 + lane] = acc;
  // block-norm energy: sum over the 9 unsigned orientations of (h[o] + h[o+9])^2, in order
  const float hi = __shfl_down_sync(0xffffffffu, acc, 9);
  const float s = __fadd_rn(acc, hi);
  const float sq = __fmul_rn(s, s);
  float e = 0.f;
#pragma unroll
  for (int o = 0; o < 9; ++o) e = __fadd_rn(e, __shfl_sync(0xffffffffu, sq, o));
  if (lane == 0) nrm[cell] = e;
}

__global__ void __launch_bounds__(256) hog_feat_kernel(const float* __restrict__ hist, const float* __restrict__ nrm,
                                                       const __grid_constant__ PvHogGeo g, __nv_bfloat16* __restrict__ feat) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= g.total_feat) return;
  const int b = blockIdx.y;
  const int l = level_of(g, idx, 2);
  const PvHogLevel L = g.lv[l];
  const int hx = L.cx - 2;
  const int r = idx - L.feat_off;
  const int y = r / hx, x = r - y * hx;
  const int Y = y + 1, X = x + 1;
  const float* nb = nrm + (long long)b * g.total_cells + L.cell_off;
  float ns[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int y0 = Y - 1 + (k >> 1), x0 = X - 1 + (k & 1);
    const float blk = __fadd_rn(__fadd_rn(__fadd_rn(nb[y0 * L.cx + x0], nb[y0 * L.cx + x0 + 1]), nb[(y0 + 1) * L.cx + x0]),
                                nb[(y0 + 1) * L.cx + x0 + 1]);
    ns[k] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(blk, 0.0001f)));
  }
  const float* h = hist + ((long long)b * g.total_cells + L.cell_off + Y * L.cx + X) * 18;
  float hv[18];
#pragma unroll
  for (int o = 0; o < 18; ++o) hv[o] = h[o];
  float out[32];
  float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 18; ++o) {
    float hk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) hk[k] = fminf(__fmul_rn(hv[o], ns[k]), 0.2f);
    out[o] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(hk[0], hk[1]), hk[2]), hk[3]));
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = __fadd_rn(t[k], hk[k]);
  }
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    const float s = __fadd_rn(hv[o], hv[o + 9]);
    float hk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) hk[k] = fminf(__fmul_rn(s, ns[k]), 0.2f);
    out[18 + o] = __fmul_rn(0.5f, __fadd_rn(__fadd_rn(__fadd_rn(hk[0], hk[1]), hk[2]), hk[3]));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[27 + k] = __fmul_rn(0.2357f, t[k]);
  out[31] = 0.f;
  __nv_bfloat16* dp = feat + (((long long)b * g.FH + L.fy0 + y) * g.fpitch + L.fx0 + x) * 32;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = pv_pack_bf16x2(out[16 * q + 2 * k], out[16 * q + 2 * k + 1]);
    pv_stg256(dp + 16 * q, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
  }
}

__global__ void __launch_bounds__(256) hog_cand_kernel(const float* __restrict__ scores, int OHs, int opitch,
                                                       const __grid_constant__ PvHogGeo g, const float* __restrict__ thr, int D,
                                                       int cap, int* __restrict__ counts, float* __restrict__ cand_score,
                                                       int* __restrict__ cand_code) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= g.total_feat) return;
  const int b = blockIdx.y;
  const int l = level_of(g, idx, 2);
  const PvHogLevel L = g.lv[l];
  const int hx = L.cx - 2;
  const int r = idx - L.feat_off;
  const int y = r / hx, x = r - y * hx;
  // score-map entry (y, x) of the level = conv output at tile position + 1 (the conv pads 5, dlib's map 4 cells)
  const float* sp = scores + (((long long)b * OHs + L.fy0 + y + 1) * opitch + L.fx0 + x + 1) * 16;
  for (int d = 0; d < D; ++d) {
    const float s = sp[d];
    if (s >= thr[d]) {
      const int slot = atomicAdd(&counts[b], 1);
      if (slot < cap) {
        cand_score[(long long)b * cap + slot] = s;
        cand_code[(long long)b * cap + slot] = (l << 23) | (d << 20) | (y << 10) | x;
      }
    }
  }
}

__device__ __forceinline__ long long rect_area(int l, int t, int r, int b) {
  const long long w = (long long)r - l + 1, h = (long long)b - t + 1;
  return (w > 0 && h > 0) ? w * h : 0;
}
// dlib test_box_overlap: inner / area(bounding box of both) > iou, or inner / area(either) > covered
__device__ __forceinline__ bool hog_overlap(const int4 a, const int4 b, double iou, double cov) {
  const long long inner = rect_area(max(a.x, b.x), max(a.y, b.y), min(a.z, b.z), min(a.w, b.w));
  if (inner == 0) return false;
  const long long outer = rect_area(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
  if ((double)inner / (double)outer > iou) return true;
  if ((double)inner / (double)rect_area(a.x, a.y, a.z, a.w) > cov || (double)inner / (double)rect_area(b.x, b.y, b.z, b.w) > cov) return true;
  return false;
}
__device__ __forceinline__ int fhog_to_image(int p) {
  const int v = (p + 1 - kPadOff) * kCell + 1;
  return v >= 0 ? v + kCell / 2 : v - kCell / 2;
}
__device__ __forceinline__ int rect_up(int v, int levels, int upsampled) {
  double p = (double)v;
  for (int i = 0; i < levels; ++i) p = __dadd_rn(__dmul_rn(p, 6.0 / 5.0), 0.3);      // pyramid_down<6>::point_up
  if (upsampled) p = p / 2.0;                                                       // pyramid_down<2>::point_down
  return (int)floor(__dadd_rn(p, 0.5));
}

__global__ void __launch_bounds__(1024) hog_nms_kernel(const int* __restrict__ counts, const float* __restrict__ cand_score,
                                                       const int* __restrict__ cand_code, int cap, int upsampled, double iou,
                                                       double cov, int max_det, int* __restrict__ out_boxes,
                                                       float* __restrict__ out_scores, int* __restrict__ out_which,
                                                       int* __restrict__ out_counts) {
  extern __shared__ uint8_t sm[];
  float* s_score = reinterpret_cast<float*>(sm);
  int* s_code = reinterpret_cast<int*>(s_score + kNmsCap);
  int4* s_box = reinterpret_cast<int4*>(s_code + kNmsCap);
  uint8_t* s_dead = reinterpret_cast<uint8_t*>(s_box + kNmsCap);
  __shared__ int s_nkept;
  const int n = blockIdx.x;
  const int cnt = counts[n];
  if (cnt > cap || cnt > kNmsCap) {  // overflow: report, do not guess
    if (threadIdx.x == 0) out_counts[n] = -cnt;
    return;
  }
  int npow = 1;
  while (npow < cnt) npow <<= 1;
  for (int i = threadIdx.x; i < npow; i += blockDim.x) {
    if (i < cnt) {
      s_score[i] = cand_score[(long long)n * cap + i];
      s_code[i] = cand_code[(long long)n * cap + i];
    } else {
      s_score[i] = -INFINITY;
      s_code[i] = 0x7fffffff;
    }
  }
  __syncthreads();
  // bitonic sort: "a before b" iff score_a > score_b or (== and code_a < code_b): code = (level, filter, row, col)
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const float sa = s_score[i], sb = s_score[ixj];
          const int ca = s_code[i], cb = s_code[ixj];
          const bool a_first = (sa > sb) || (sa == sb && ca < cb);
          const bool up = ((i & k) == 0);
          if (up ? !a_first : a_first) {
            s_score[i] = sb; s_score[ixj] = sa;
            s_code[i] = cb; s_code[ixj] = ca;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int code = s_code[i];
    const int lv = code >> 23, y = (code >> 10) & 1023, x = code & 1023;
    // dlib saliency position = the filter's centre in padded coordinates (y + 5, x + 5); box = centered_rect of kDetBox cells
    const int c = x + kFilt / 2, r = y + kFilt / 2;
    const int l0 = c - kDetBox / 2, t0 = r - kDetBox / 2;
    int4 b;
    b.x = rect_up(fhog_to_image(l0), lv, upsampled);
    b.y = rect_up(fhog_to_image(t0), lv, upsampled);
    b.z = rect_up(fhog_to_image(l0 + kDetBox - 1), lv, upsampled);
    b.w = rect_up(fhog_to_image(t0 + kDetBox - 1), lv, upsampled);
    s_box[i] = b;
    s_dead[i] = 0;
  }
  if (threadIdx.x == 0) s_nkept = 0;
  __syncthreads();
  for (int i = 0; i < cnt; ++i) {
    if (s_dead[i] == 0) {
      const int4 bi = s_box[i];
      if (threadIdx.x == 0) {
        const int k = s_nkept;
        if (k < max_det) {
          int* ob = out_boxes + ((long long)n * max_det + k) * 4;
          ob[0] = bi.x; ob[1] = bi.y; ob[2] = bi.z; ob[3] = bi.w;
          out_scores[(long long)n * max_det + k] = s_score[i];
          out_which[(long long)n * max_det + k] = (s_code[i] >> 20) & 7;
        }
        s_nkept = k + 1;
      }
      for (int j = i + 1 + threadIdx.x; j < cnt; j += blockDim.x)
        if (s_dead[j] == 0 && hog_overlap(s_box[j], bi, iou, cov)) s_dead[j] = 1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[n] = s_nkept;
}

}  // namespace

extern "C" int pv_hog_features(const void* plane_rgba, int B, const PvHogGeo* geo, const float* uv18, void* ori_u8, void* mag_f32,
                               float* hist, float* nrm, void* feat_bf16, void* stream) {
  PV_REQUIRE(plane_rgba && geo && uv18 && ori_u8 && mag_f32 && hist && nrm && feat_bf16, "pv_hog_features: null argument");
  PV_REQUIRE(B > 0 && geo->n_levels > 0 && geo->n_levels <= PV_HOG_MAX_LEVELS, "pv_hog_features: B=%d levels=%d", B, geo->n_levels);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(feat_bf16) & 31) == 0, "pv_hog_features: feature plane must be 32-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PvHogGeo g = *geo;
  hog_grad_kernel<<<dim3((unsigned)((g.total_px + 255) / 256), (unsigned)B), 256, 0, st>>>(
      static_cast<const uint32_t*>(plane_rgba), g, static_cast<uint8_t*>(ori_u8), static_cast<float*>(mag_f32), uv18);
  hog_hist_kernel<<<dim3((unsigned)(((long long)g.total_cells * 32 + 255) / 256), (unsigned)B), 256, 0, st>>>(
      static_cast<const uint8_t*>(ori_u8), static_cast<const float*>(mag_f32), g, hist, nrm);
  if (g.total_feat > 0)
    hog_feat_kernel<<<dim3((unsigned)((g.total_feat + 255) / 256), (unsigned)B), 256, 0, st>>>(
        hist, nrm, g, static_cast<__nv_bfloat16*>(feat_bf16));
  g_pv_launches.fetch_add(3);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_hog_decode(const float* scores, int B, int OHs, int opitch, const PvHogGeo* geo, const float* thr_dev, int D,
                             int upsampled, double iou_thresh, double covered_thresh, int cap, int max_det, int* counts,
                             float* cand_score, int* cand_code, int* out_boxes, float* out_scores, int* out_which, int* out_counts,
                             void* stream) {
  PV_REQUIRE(scores && geo && thr_dev && counts && cand_score && cand_code && out_boxes && out_scores && out_which && out_counts,
             "pv_hog_decode: null argument");
  PV_REQUIRE(B > 0 && D > 0 && D <= 8 && cap > 0 && cap <= kNmsCap && max_det > 0, "pv_hog_decode: D=%d cap=%d", D, cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PvHogGeo g = *geo;
  PV_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int) * B, st));
  if (g.total_feat > 0)
    hog_cand_kernel<<<dim3((unsigned)((g.total_feat + 255) / 256), (unsigned)B), 256, 0, st>>>(scores, OHs, opitch, g, thr_dev, D, cap,
                                                                                               counts, cand_score, cand_code);
  const size_t smem = (size_t)kNmsCap * (4 + 4 + 16 + 1);
  static unsigned long long attr = 0;
  if (pv_attr_needed(&attr)) PV_CUDA_CHECK(cudaFuncSetAttribute(hog_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hog_nms_kernel<<<B, 1024, smem, st>>>(counts, cand_score, cand_code, cap, upsampled, iou_thresh, covered_thresh, max_det, out_boxes,
                                        out_scores, out_which, out_counts);
  g_pv_launches.fetch_add(2);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
