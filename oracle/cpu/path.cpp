// ORACLE (test infrastructure) — byte / integer stages of the C++ CPU restatement (see pvcpu.h): image pyramid plane,
// MMOD decode + NMS, ERT landmarks, face chips, DSST correlation tracker.  Every function mirrors the numpy oracle it
// cites operation by operation (float32, unfused: this file is compiled with -ffp-contract=off), so that the two
// restatements can be compared bit for bit (tests/test_cpu_ref_cpu.py).
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

#include "pvcpu.h"

extern "C" int pvc_set_threads(int n) {
  if (n <= 0) n = omp_get_num_procs();
  omp_set_num_threads(n);
  return n;
}
extern "C" int pvc_get_threads(void) { return omp_get_max_threads(); }

// ------------------------------------------------------------------------------------------------------------------
// oracle/pyramid.py: resize_bilinear_u8 + build_plane
// ------------------------------------------------------------------------------------------------------------------
static void resize_bilinear(const uint8_t* src, int sc, long s_row, int H, int W, uint8_t* dst, long d_row, int oh, int ow) {
  // src pixel (y,x) channel c at src[y*s_row + x*sc + c]; dst is RGBA (4 bytes / pixel), A = 255
  const float ys = (float)(H - 1) / (float)std::max(oh - 1, 1);
  const float xs = (float)(W - 1) / (float)std::max(ow - 1, 1);
  std::vector<int> left(ow), right(ow);
  std::vector<float> lr(ow);
  for (int x = 0; x < ow; ++x) {
    const float fx = (float)x * xs;
    int l = (int)std::floor(fx);
    l = std::min(l, W - 1);
    left[x] = l;
    right[x] = std::min(l + 1, W - 1);
    lr[x] = fx - (float)l;
  }
#pragma omp parallel for schedule(static)
  for (int y = 0; y < oh; ++y) {
    const float fy = (float)y * ys;
    int top = (int)std::floor(fy);
    top = std::min(top, H - 1);
    const int bot = std::min(top + 1, H - 1);
    const float tb = fy - (float)top;
    const uint8_t* rt = src + (long)top * s_row;
    const uint8_t* rb = src + (long)bot * s_row;
    uint8_t* d = dst + (long)y * d_row;
    for (int x = 0; x < ow; ++x) {
      const float l = lr[x], oml = 1.0f - l, omt = 1.0f - tb;
      for (int c = 0; c < 3; ++c) {
        const float tl = rt[(long)left[x] * sc + c], tr = rt[(long)right[x] * sc + c];
        const float bl = rb[(long)left[x] * sc + c], br = rb[(long)right[x] * sc + c];
        const float a = (oml * tl) + (l * tr);
        const float b = (oml * bl) + (l * br);
        float v = (omt * a) + (tb * b);
        v = std::floor(v + 0.5f);
        v = std::min(std::max(v, 0.0f), 255.0f);
        d[4 * x + c] = (uint8_t)v;
      }
      d[4 * x + 3] = 255;
    }
  }
}

extern "C" int pvc_build_plane(const uint8_t* rgb, int H, int W, int upsample, const int* rects, int n_levels, int Hp, int Wp,
                               uint8_t* plane) {
  memset(plane, 0, (size_t)Hp * Wp * 4);
  const long prow = (long)Wp * 4;
  for (int lv = 0; lv < n_levels; ++lv) {
    const int x0 = rects[4 * lv], y0 = rects[4 * lv + 1], w = rects[4 * lv + 2], h = rects[4 * lv + 3];
    uint8_t* dst = plane + (long)y0 * prow + (long)x0 * 4;
    if (lv == 0) {
      if (upsample) {
        resize_bilinear(rgb, 3, (long)W * 3, H, W, dst, prow, h, w);
      } else {
        if (h != H || w != W) return -1;
        for (int y = 0; y < H; ++y)
          for (int x = 0; x < W; ++x) {
            for (int c = 0; c < 3; ++c) dst[(long)y * prow + 4 * x + c] = rgb[((long)y * W + x) * 3 + c];
            dst[(long)y * prow + 4 * x + 3] = 255;
          }
      }
    } else {
      const int px0 = rects[4 * (lv - 1)], py0 = rects[4 * (lv - 1) + 1], pw = rects[4 * (lv - 1) + 2], ph = rects[4 * (lv - 1) + 3];
      const uint8_t* src = plane + (long)py0 * prow + (long)px0 * 4;
      resize_bilinear(src, 4, prow, ph, pw, dst, prow, h, w);
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// oracle/pyramid.py: decode (loss_mmod::to_label) with oracle/geometry.py box_from_plane / level_at
// ------------------------------------------------------------------------------------------------------------------
static long rect_area(long l, long t, long r, long b) {
  if (r < l || b < t) return 0;
  return (r - l + 1) * (b - t + 1);
}
static bool boxes_overlap(const int* a, const int* b, double iou_thresh, double covered_thresh) {
  const long inner = rect_area(std::max(a[0], b[0]), std::max(a[1], b[1]), std::min(a[2], b[2]), std::min(a[3], b[3]));
  if (inner == 0) return false;
  const long aa = rect_area(a[0], a[1], a[2], a[3]), ab = rect_area(b[0], b[1], b[2], b[3]);
  const long outer = aa + ab - inner;
  if ((double)inner / (double)outer > iou_thresh) return true;
  if ((double)inner / (double)aa > covered_thresh || (double)inner / (double)ab > covered_thresh) return true;
  return false;
}

extern "C" int pvc_decode(const float* scores, int OH, int OW, const int* rects, const float* fxy, int n_levels, int window,
                          int cell_mul, int cell_add, float threshold, double iou_thresh, double covered_thresh,
                          int max_candidates, int max_out, int* boxes, float* out_scores) {
  struct Cand { float s; long cell; int box[4]; };
  std::vector<Cand> cands;
  for (int r = 0; r < OH; ++r)
    for (int c = 0; c < OW; ++c) {
      const float s = scores[(long)r * OW + c];
      if (!(s > threshold)) continue;
      const int px = cell_mul * c + cell_add, py = cell_mul * r + cell_add;
      int lv = -1;
      for (int k = 0; k < n_levels; ++k) {
        const int x0 = rects[4 * k], y0 = rects[4 * k + 1], w = rects[4 * k + 2], h = rects[4 * k + 3];
        if (px >= x0 && px < x0 + w && py >= y0 && py < y0 + h) { lv = k; break; }
      }
      if (lv < 0) continue;
      const int l = (px - rects[4 * lv]) - window / 2, t = (py - rects[4 * lv + 1]) - window / 2;
      const int rr = l + window - 1, bb = t + window - 1;
      const float fx = fxy[2 * lv], fy = fxy[2 * lv + 1];
      Cand cd;
      cd.s = s;
      cd.cell = (long)r * OW + c;
      cd.box[0] = (int)std::floor(((float)l * fx) + 0.5f);
      cd.box[1] = (int)std::floor(((float)t * fy) + 0.5f);
      cd.box[2] = (int)std::floor(((float)rr * fx) + 0.5f);
      cd.box[3] = (int)std::floor(((float)bb * fy) + 0.5f);
      cands.push_back(cd);
    }
  std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.s > b.s || (a.s == b.s && a.cell < b.cell); });
  if (max_candidates > 0 && (int)cands.size() > max_candidates) cands.resize(max_candidates);
  int n = 0;
  for (const Cand& cd : cands) {
    bool sup = false;
    for (int k = 0; k < n && !sup; ++k) sup = boxes_overlap(cd.box, boxes + 4 * k, iou_thresh, covered_thresh);
    if (sup) continue;
    if (n >= max_out) return -1;
    memcpy(boxes + 4 * n, cd.box, sizeof(int) * 4);
    out_scores[n] = cd.s;
    ++n;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------------------------
// oracle/landmarks.py: similarity_fit, ert_predict, chip_transform, extract_chips
// ------------------------------------------------------------------------------------------------------------------
struct Sim { float m00, m01, m10, m11, tx, ty; };

// least-squares similarity b ~ M a + t over the points idx[0..n): sequential float32 sums in index order
static Sim similarity_fit(const float* a, const float* b, const int* idx, int n) {
  const float nf = (float)n;
  float sax = 0, say = 0, sbx = 0, sby = 0;
  for (int k = 0; k < n; ++k) {
    const int i = idx ? idx[k] : k;
    sax = sax + a[2 * i];
    say = say + a[2 * i + 1];
    sbx = sbx + b[2 * i];
    sby = sby + b[2 * i + 1];
  }
  const float max_ = sax / nf, may = say / nf, mbx = sbx / nf, mby = sby / nf;
  float A = 0, Bc = 0, den = 0;
  for (int k = 0; k < n; ++k) {
    const int i = idx ? idx[k] : k;
    const float acx = a[2 * i] - max_, acy = a[2 * i + 1] - may;
    const float bcx = b[2 * i] - mbx, bcy = b[2 * i + 1] - mby;
    A = A + ((acx * bcx) + (acy * bcy));
    Bc = Bc + ((acx * bcy) - (acy * bcx));
    den = den + ((acx * acx) + (acy * acy));
  }
  Sim s;
  s.m00 = A / den;
  s.m10 = Bc / den;
  s.m01 = -s.m10;
  s.m11 = s.m00;
  s.tx = mbx - ((s.m00 * max_) + (s.m01 * may));
  s.ty = mby - ((s.m10 * max_) + (s.m11 * may));
  return s;
}

extern "C" int pvc_ert_predict(const uint8_t* rgb, int H, int W, const float* initial_shape, const int* anchor_idx,
                               const float* deltas, const int* split_idx1, const int* split_idx2, const float* split_thresh,
                               const float* leaf_values, int stages, int trees, int pool, const int* rects, int M,
                               int64_t* out_parts) {
  constexpr int P = 68, NSPLIT = 15, NLEAF = 16;
#pragma omp parallel for schedule(dynamic, 1)
  for (int m = 0; m < M; ++m) {
    const float l = (float)rects[4 * m], t = (float)rects[4 * m + 1];
    const float wr = (float)(rects[4 * m + 2] - rects[4 * m]), hr = (float)(rects[4 * m + 3] - rects[4 * m + 1]);
    float cur[2 * P];
    memcpy(cur, initial_shape, sizeof(cur));
    std::vector<float> feat(pool);
    for (int s = 0; s < stages; ++s) {
      const Sim sm = similarity_fit(initial_shape, cur, nullptr, P);
      for (int q = 0; q < pool; ++q) {
        const int an = anchor_idx[(long)s * pool + q];
        const float d0 = deltas[((long)s * pool + q) * 2], d1 = deltas[((long)s * pool + q) * 2 + 1];
        const float dx = ((sm.m00 * d0) + (sm.m01 * d1)) + cur[2 * an];
        const float dy = ((sm.m10 * d0) + (sm.m11 * d1)) + cur[2 * an + 1];
        const float px = l + (dx * wr), py = t + (dy * hr);
        const long ix = (long)std::floor(px + 0.5f), iy = (long)std::floor(py + 0.5f);
        float f = 0.f;
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) {
          const uint8_t* p = rgb + ((long)iy * W + ix) * 3;
          f = (float)(((unsigned)p[0] + p[1] + p[2]) / 3u);
        }
        feat[q] = f;
      }
      for (int tr = 0; tr < trees; ++tr) {
        const long tb = ((long)s * trees + tr) * NSPLIT;
        int node = 0;
        while (node < NSPLIT) {
          const float diff = feat[split_idx1[tb + node]] - feat[split_idx2[tb + node]];
          node = diff > split_thresh[tb + node] ? 2 * node + 1 : 2 * node + 2;
        }
        const float* leaf = leaf_values + (((long)s * trees + tr) * NLEAF + (node - NSPLIT)) * (2 * P);
        for (int i = 0; i < 2 * P; ++i) cur[i] = cur[i] + leaf[i];
      }
    }
    for (int i = 0; i < P; ++i) {
      const float x = l + (cur[2 * i] * wr), y = t + (cur[2 * i + 1] * hr);
      out_parts[((long)m * P + i) * 2] = (int64_t)std::floor(x + 0.5f);
      out_parts[((long)m * P + i) * 2 + 1] = (int64_t)std::floor(y + 0.5f);
    }
  }
  return 0;
}

extern "C" int pvc_extract_chips(const uint8_t* rgb, int H, int W, const int64_t* parts, int M, const float* from_pts,
                                 const int* pt_idx, int n_pts, int size, uint8_t* chips) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int m = 0; m < M; ++m) {
    float to[2 * 68];
    for (int i = 0; i < 2 * 68; ++i) to[i] = (float)parts[(long)m * 136 + i];
    const Sim sm = similarity_fit(from_pts, to, pt_idx, n_pts);
    uint8_t* out = chips + (long)m * size * size * 3;
    for (int r = 0; r < size; ++r)
      for (int c = 0; c < size; ++c) {
        const float x = ((sm.m00 * (float)c) + (sm.m01 * (float)r)) + sm.tx;
        const float y = ((sm.m10 * (float)c) + (sm.m11 * (float)r)) + sm.ty;
        const long left = (long)std::floor(x), top = (long)std::floor(y);
        const long right = left + 1, bot = top + 1;
        uint8_t* o = out + ((long)r * size + c) * 3;
        if (!(left >= 0 && top >= 0 && right < W && bot < H)) {
          o[0] = o[1] = o[2] = 0;
          continue;
        }
        const float lr = x - (float)left, tb = y - (float)top;
        const float oml = 1.0f - lr, omt = 1.0f - tb;
        const uint8_t* ptl = rgb + ((long)top * W + left) * 3;
        const uint8_t* pbl = ptl + (long)W * 3;
        for (int ch = 0; ch < 3; ++ch) {
          const float a = (oml * (float)ptl[ch]) + (lr * (float)ptl[3 + ch]);
          const float b = (oml * (float)pbl[ch]) + (lr * (float)pbl[3 + ch]);
          float v = (omt * a) + (tb * b);
          v = std::min(std::max(std::floor(v + 0.5f), 0.0f), 255.0f);
          o[ch] = (uint8_t)v;
        }
      }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// oracle/dsst.py: correlation tracker (translation filter + scale filter), double-precision spectra like dlib
// ------------------------------------------------------------------------------------------------------------------
namespace {
typedef std::complex<double> cd;
constexpr int FS = 64, NFH = 31, NCH = 32, NPIX = FS * FS;   // 31 FHOG planes + the intensity plane
constexpr int NSC = 32, SW = 23, SCELLS = 6, SOUT = 4, SF = 31 * SOUT * SOUT;
constexpr double kPi = 3.14159265358979323846;

struct Tables {
  float hann[FS], hann_s[NSC], uu[9], vv[9], factor[NSC];
  cd tw64[FS], tw32[NSC];
  Tables() {
    for (int i = 0; i < FS; ++i) hann[i] = (float)(0.5 - 0.5 * std::cos(2 * kPi * i / (FS - 1)));
    for (int i = 0; i < NSC; ++i) hann_s[i] = (float)(0.5 - 0.5 * std::cos(2 * kPi * i / (NSC - 1)));
    for (int o = 0; o < 9; ++o) { uu[o] = (float)std::cos(o * kPi / 9); vv[o] = (float)std::sin(o * kPi / 9); }
    for (int k = 0; k < NSC; ++k) factor[k] = std::pow((float)1.020, (float)(k - NSC / 2));
    for (int k = 0; k < FS; ++k) tw64[k] = cd(std::cos(-2 * kPi * k / FS), std::sin(-2 * kPi * k / FS));
    for (int k = 0; k < NSC; ++k) tw32[k] = cd(std::cos(-2 * kPi * k / NSC), std::sin(-2 * kPi * k / NSC));
  }
};
static const Tables& tables() {
  static Tables t;
  return t;
}

// n x n bilinear RGB chip of the float rect (rl, rt) + pixel * (sx, sy); outside -> 0 (extract_chip / extract_chip_n)
static void chip_sample(const uint8_t* rgb, int H, int W, float rl, float rt, float sx, float sy, int n, uint8_t* chip) {
  for (int y = 0; y < n; ++y) {
    const float fy = rt + ((float)y * sy);
    const long top = (long)std::floor(fy);
    const float tb = fy - (float)top;
    for (int x = 0; x < n; ++x) {
      const float fx = rl + ((float)x * sx);
      const long left = (long)std::floor(fx);
      const float lr = fx - (float)left;
      uint8_t* o = chip + ((long)y * n + x) * 3;
      if (!(left >= 0 && left + 1 < W && top >= 0 && top + 1 < H)) { o[0] = o[1] = o[2] = 0; continue; }
      const uint8_t* ptl = rgb + ((long)top * W + left) * 3;
      const uint8_t* pbl = ptl + (long)W * 3;
      const float oml = 1.0f - lr, omt = 1.0f - tb;
      for (int c = 0; c < 3; ++c) {
        const float a = (oml * (float)ptl[c]) + (lr * (float)ptl[3 + c]);
        const float b = (oml * (float)pbl[c]) + (lr * (float)pbl[3 + c]);
        float v = (omt * a) + (tb * b);
        v = std::min(std::max(std::floor(v + 0.5f), 0.0f), 255.0f);
        o[c] = (uint8_t)v;
      }
    }
  }
}

static inline void grad_at(const uint8_t* chip, int n, int y, int x, float& gx, float& gy, float& best) {
  best = -1.f; gx = gy = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float dx = (float)chip[((long)y * n + x + 1) * 3 + c] - (float)chip[((long)y * n + x - 1) * 3 + c];
    const float dy = (float)chip[((long)(y + 1) * n + x) * 3 + c] - (float)chip[((long)(y - 1) * n + x) * 3 + c];
    const float v = (dx * dx) + (dy * dy);
    if (v > best) { best = v; gx = dx; gy = dy; }
  }
}
static inline int snap_orientation(const Tables& tb, float gx, float gy) {
  int bo = 0;
  float best_dot = 0.f;
  for (int o = 0; o < 9; ++o) {
    const float dot = (tb.uu[o] * gx) + (tb.vv[o] * gy);
    if (dot > best_dot) { best_dot = dot; bo = o; }
    else if (-dot > best_dot) { best_dot = -dot; bo = o + 9; }
  }
  return bo;
}

// fhog_cell1: float [31][64][64]
static void fhog_cell1(const uint8_t* chip, float* out) {
  const Tables& tb = tables();
  const int n = FS;
  std::vector<float> mag(NPIX, 0.f), P((n + 2) * (n + 2), 0.f);
  std::vector<int> ori(NPIX, 0);
  for (int y = 1; y < n - 1; ++y)
    for (int x = 1; x < n - 1; ++x) {
      float gx, gy, best;
      grad_at(chip, n, y, x, gx, gy, best);
      mag[y * n + x] = std::sqrt(best);
      ori[y * n + x] = snap_orientation(tb, gx, gy);
    }
  const int pn = n + 2;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) P[(y + 1) * pn + x + 1] = mag[y * n + x] * mag[y * n + x];
  auto blk = [&](int y, int x, int dy, int dx) {
    const int y0 = 1 + dy + y, x0 = 1 + dx + x;
    return ((P[y0 * pn + x0] + P[y0 * pn + x0 + 1]) + P[(y0 + 1) * pn + x0]) + P[(y0 + 1) * pn + x0 + 1];
  };
  memset(out, 0, sizeof(float) * NFH * NPIX);
  const int dys[4] = {-1, -1, 0, 0}, dxs[4] = {-1, 0, -1, 0};
  for (int y = 1; y < n - 1; ++y)
    for (int x = 1; x < n - 1; ++x) {
      const int i = y * n + x;
      float h[4];
      for (int k = 0; k < 4; ++k) {
        const float nk = 1.0f / std::sqrt(blk(y, x, dys[k], dxs[k]) + 0.0001f);
        h[k] = std::min(mag[i] * nk, 0.2f);
      }
      const float osum = 0.5f * (((h[0] + h[1]) + h[2]) + h[3]);
      const int o = ori[i];
      out[(long)o * NPIX + i] = osum;
      out[(long)(18 + o % 9) * NPIX + i] = osum;
      for (int k = 0; k < 4; ++k) out[(long)(27 + k) * NPIX + i] = 0.2357f * h[k];
    }
}

// fhog_cell4 of a 23x23 chip -> float [31][4][4]
static void fhog_cell4(const uint8_t* chip, float* out) {
  const Tables& tb = tables();
  const int n = SW, cells = SCELLS;
  float hist[SCELLS][SCELLS][18];
  memset(hist, 0, sizeof(hist));
  for (int y = 1; y < n - 1; ++y)
    for (int x = 1; x < n - 1; ++x) {
      float gx, gy, best;
      grad_at(chip, n, y, x, gx, gy, best);
      const float mag = std::sqrt(best);
      const int bo = snap_orientation(tb, gx, gy);
      const float xp = ((float)x + 0.5f) / 4.0f - 0.5f, yp = ((float)y + 0.5f) / 4.0f - 0.5f;
      const int ixp = (int)std::floor(xp), iyp = (int)std::floor(yp);
      const float vx0 = xp - (float)ixp, vy0 = yp - (float)iyp;
      const float vx1 = 1.0f - vx0, vy1 = 1.0f - vy0;
      if (ixp >= 0 && iyp >= 0) hist[iyp][ixp][bo] += (vx1 * vy1) * mag;
      if (ixp + 1 < cells && iyp >= 0) hist[iyp][ixp + 1][bo] += (vx0 * vy1) * mag;
      if (ixp >= 0 && iyp + 1 < cells) hist[iyp + 1][ixp][bo] += (vx1 * vy0) * mag;
      if (ixp + 1 < cells && iyp + 1 < cells) hist[iyp + 1][ixp + 1][bo] += (vx0 * vy0) * mag;
    }
  float nrm[SCELLS][SCELLS];
  for (int y = 0; y < cells; ++y)
    for (int x = 0; x < cells; ++x) {
      float s = 0.f;
      for (int o = 0; o < 9; ++o) {
        const float v = hist[y][x][o] + hist[y][x][o + 9];
        s = s + (v * v);
      }
      nrm[y][x] = s;
    }
  const int dys[4] = {-1, -1, 0, 0}, dxs[4] = {-1, 0, -1, 0};
  for (int y = 0; y < SOUT; ++y)
    for (int x = 0; x < SOUT; ++x) {
      const int Y = y + 1, X = x + 1;
      const float* h = hist[Y][X];
      float ns[4];
      for (int k = 0; k < 4; ++k) {
        const int y0 = Y + dys[k], x0 = X + dxs[k];
        const float b = ((nrm[y0][x0] + nrm[y0][x0 + 1]) + nrm[y0 + 1][x0]) + nrm[y0 + 1][x0 + 1];
        ns[k] = 1.0f / std::sqrt(b + 0.0001f);
      }
      float t[4] = {0, 0, 0, 0};
      for (int o = 0; o < 18; ++o) {
        float hk[4];
        for (int k = 0; k < 4; ++k) hk[k] = std::min(h[o] * ns[k], 0.2f);
        out[(o * SOUT + y) * SOUT + x] = 0.5f * (((hk[0] + hk[1]) + hk[2]) + hk[3]);
        for (int k = 0; k < 4; ++k) t[k] = t[k] + hk[k];
      }
      for (int o = 0; o < 9; ++o) {
        const float s = h[o] + h[o + 9];
        float hk[4];
        for (int k = 0; k < 4; ++k) hk[k] = std::min(s * ns[k], 0.2f);
        out[((18 + o) * SOUT + y) * SOUT + x] = 0.5f * (((hk[0] + hk[1]) + hk[2]) + hk[3]);
      }
      for (int k = 0; k < 4; ++k) out[((27 + k) * SOUT + y) * SOUT + x] = 0.2357f * t[k];
    }
}

static void fft1d(cd* x, int n, int stride, const cd* tw, bool inverse) {
  // iterative radix-2 DIT; tw[k] = exp(-2 pi i k / n)
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  for (int i = 0; i < n; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    if (r > i) std::swap(x[(long)i * stride], x[(long)r * stride]);
  }
  for (int half = 1; half < n; half <<= 1) {
    const int step = n / (2 * half);
    for (int g = 0; g < n; g += 2 * half)
      for (int j = 0; j < half; ++j) {
        cd w = tw[j * step];
        if (inverse) w = std::conj(w);
        cd& a = x[(long)(g + j) * stride];
        cd& b = x[(long)(g + j + half) * stride];
        const cd t = w * b;
        b = a - t;
        a = a + t;
      }
  }
}
static void fft2(cd* x, bool inverse) {
  const Tables& tb = tables();
  for (int y = 0; y < FS; ++y) fft1d(x + (long)y * FS, FS, 1, tb.tw64, inverse);
  for (int c = 0; c < FS; ++c) fft1d(x + c, FS, FS, tb.tw64, inverse);
  if (inverse)
    for (int i = 0; i < NPIX; ++i) x[i] /= (double)NPIX;
}

struct Track {
  double pos[4];
  std::vector<cd> A;      // [32][4096]
  std::vector<double> B;  // [4096]
  std::vector<cd> As;     // [496][32]
  std::vector<double> Bs; // [32]
};
struct Bank {
  std::vector<Track> tr;
  bool use_scale;
};

static void features(const uint8_t* rgb, int H, int W, const double* rect, std::vector<cd>& F, float tf[4]) {
  const Tables& tb = tables();
  const float l = (float)rect[0], t = (float)rect[1], r = (float)rect[2], b = (float)rect[3];
  const float cx = (l + r) * 0.5f, cy = (t + b) * 0.5f;
  const float hw = ((r - l) * 0.5f) * 1.4f, hh = ((b - t) * 0.5f) * 1.4f;
  const float rl = cx - hw, rt = cy - hh;
  const float sx = (2.0f * hw) / (float)(FS - 1), sy = (2.0f * hh) / (float)(FS - 1);
  tf[0] = rl; tf[1] = rt; tf[2] = sx; tf[3] = sy;
  std::vector<uint8_t> chip(NPIX * 3);
  chip_sample(rgb, H, W, rl, rt, sx, sy, FS, chip.data());
  std::vector<float> fh((size_t)NCH * NPIX);
  fhog_cell1(chip.data(), fh.data());
  // 32nd plane: chip intensity (r + g + b) / 3 (unsigned integer division, dlib assign_pixel) / 255
  for (int i = 0; i < NPIX; ++i) {
    const unsigned g = ((unsigned)chip[3 * i] + chip[3 * i + 1] + chip[3 * i + 2]) / 3u;
    fh[(size_t)NFH * NPIX + i] = (float)g / 255.0f;
  }
  F.resize((size_t)NCH * NPIX);
  for (int ch = 0; ch < NCH; ++ch) {
    for (int y = 0; y < FS; ++y)
      for (int x = 0; x < FS; ++x) {
        const float w = tb.hann[y] * tb.hann[x];
        F[(size_t)ch * NPIX + y * FS + x] = cd((double)(fh[(size_t)ch * NPIX + y * FS + x] * w), 0.0);
      }
    fft2(&F[(size_t)ch * NPIX], false);
  }
}

static void target_hat(double px, double py, std::vector<cd>& G) {
  G.resize(NPIX);
  for (int y = 0; y < FS; ++y)
    for (int x = 0; x < FS; ++x) {
      const float dx = (float)x - (float)px, dy = (float)y - (float)py;
      G[y * FS + x] = cd((double)std::exp(-((dx * dx) + (dy * dy)) / 3.0f), 0.0);
    }
  fft2(G.data(), false);
  for (auto& g : G) g = std::conj(g);
}

static void scale_space(const uint8_t* rgb, int H, int W, const double* pos, std::vector<cd>& Fs) {
  const Tables& tb = tables();
  Fs.assign((size_t)SF * NSC, cd(0, 0));
  const float l = (float)pos[0], t = (float)pos[1], r = (float)pos[2], b = (float)pos[3];
  const float cx = (l + r) * 0.5f, cy = (t + b) * 0.5f;
  std::vector<uint8_t> chip(SW * SW * 3);
  float feat[SF];
  std::vector<double> col((size_t)SF * NSC);
  for (int k = 0; k < NSC; ++k) {
    const float hw = ((r - l) * 0.5f) * tb.factor[k], hh = ((b - t) * 0.5f) * tb.factor[k];
    const float lk = cx - hw, tk = cy - hh, rk = cx + hw, bk = cy + hh;
    const float sx = (rk - lk) / (float)(SW - 1), sy = (bk - tk) / (float)(SW - 1);
    chip_sample(rgb, H, W, lk, tk, sx, sy, SW, chip.data());
    fhog_cell4(chip.data(), feat);
    for (int j = 0; j < SF; ++j) col[(size_t)j * NSC + k] = (double)(feat[j] * tb.hann_s[k]);
  }
  for (int j = 0; j < SF; ++j)
    for (int m = 0; m < NSC; ++m) {
      cd acc(0, 0);
      for (int k = 0; k < NSC; ++k) acc += col[(size_t)j * NSC + k] * tb.tw32[(k * m) & (NSC - 1)];
      Fs[(size_t)j * NSC + m] = acc;
    }
}

static void scale_target_hat(double p, cd* Gs) {
  const Tables& tb = tables();
  double g[NSC];
  for (int k = 0; k < NSC; ++k) {
    const float d = (float)k - (float)p;
    g[k] = (double)std::exp(-(d * d) / 1.0f);
  }
  for (int m = 0; m < NSC; ++m) {
    cd acc(0, 0);
    for (int k = 0; k < NSC; ++k) acc += g[k] * tb.tw32[(k * m) & (NSC - 1)];
    Gs[m] = std::conj(acc);
  }
}

static void start_one(Bank& bk, Track& tk, const uint8_t* rgb, int H, int W, const double* rect) {
  for (int k = 0; k < 4; ++k) tk.pos[k] = rect[k];
  std::vector<cd> F, G;
  float tf[4];
  features(rgb, H, W, tk.pos, F, tf);
  const double c = (FS - 1) / 2.0;
  target_hat(c, c, G);
  tk.A.resize((size_t)NCH * NPIX);
  tk.B.assign(NPIX, 0.0);
  for (int ch = 0; ch < NCH; ++ch)
    for (int i = 0; i < NPIX; ++i) {
      tk.A[(size_t)ch * NPIX + i] = G[i] * F[(size_t)ch * NPIX + i];
      tk.B[i] += std::norm(F[(size_t)ch * NPIX + i]);
    }
  if (bk.use_scale) {
    std::vector<cd> Fs;
    scale_space(rgb, H, W, tk.pos, Fs);
    cd Gs[NSC];
    scale_target_hat(NSC / 2, Gs);
    tk.As.resize((size_t)SF * NSC);
    tk.Bs.assign(NSC, 0.0);
    for (int j = 0; j < SF; ++j)
      for (int m = 0; m < NSC; ++m) {
        tk.As[(size_t)j * NSC + m] = Gs[m] * Fs[(size_t)j * NSC + m];
        tk.Bs[m] += std::norm(Fs[(size_t)j * NSC + m]);
      }
  }
}

static double update_one(Bank& bk, Track& tk, const uint8_t* rgb, int H, int W) {
  const double NU = 0.025, LAMBDA = 0.001;
  double guess[4];
  memcpy(guess, tk.pos, sizeof(guess));
  std::vector<cd> F, G;
  float tf[4];
  features(rgb, H, W, guess, F, tf);
  std::vector<cd> R(NPIX, cd(0, 0));
  for (int ch = 0; ch < NCH; ++ch)
    for (int i = 0; i < NPIX; ++i) R[i] += F[(size_t)ch * NPIX + i] * std::conj(tk.A[(size_t)ch * NPIX + i]);
  for (int i = 0; i < NPIX; ++i) R[i] /= (tk.B[i] + LAMBDA);
  fft2(R.data(), true);
  int pi = 0;
  for (int i = 1; i < NPIX; ++i)
    if (R[i].real() > R[pi].real()) pi = i;
  const int py = pi / FS, px = pi % FS;
  double ppx = px, ppy = py;
  auto Rr = [&](int y, int x) { return R[y * FS + x].real(); };
  if (px > 0 && px < FS - 1 && py > 0 && py < FS - 1) {
    // dlib max_point_interpolated: Newton step of the 3x3 finite-difference quadratic (cross term included), clamped to +-1
    const double dx = 0.5 * (Rr(py, px + 1) - Rr(py, px - 1)), dy = 0.5 * (Rr(py + 1, px) - Rr(py - 1, px));
    const double dxx = Rr(py, px + 1) - 2 * Rr(py, px) + Rr(py, px - 1);
    const double dyy = Rr(py + 1, px) - 2 * Rr(py, px) + Rr(py - 1, px);
    const double dxy = 0.25 * ((Rr(py + 1, px + 1) + Rr(py - 1, px - 1)) - (Rr(py + 1, px - 1) + Rr(py - 1, px + 1)));
    const double det = dxx * dyy - dxy * dxy;
    if (det != 0) {
      ppx += std::min(1.0, std::max(-1.0, -(dyy * dx - dxy * dy) / det));
      ppy += std::min(1.0, std::max(-1.0, -(dxx * dy - dxy * dx) / det));
    }
  }
  double sum = 0, n = 0;
  for (int y = 0; y < FS; ++y)
    for (int x = 0; x < FS; ++x) {
      if (y >= std::max(py - 4, 0) && y < py + 4 && x >= std::max(px - 4, 0) && x < px + 4) continue;
      sum += Rr(y, x);
      n += 1;
    }
  const double mean = sum / n;
  double var = 0;
  for (int y = 0; y < FS; ++y)
    for (int x = 0; x < FS; ++x) {
      if (y >= std::max(py - 4, 0) && y < py + 4 && x >= std::max(px - 4, 0) && x < px + 4) continue;
      var += (Rr(y, x) - mean) * (Rr(y, x) - mean);
    }
  const double psr = (Rr(py, px) - mean) / std::sqrt(var / (n - 1));
  const double ix = (double)tf[0] + ppx * (double)tf[2], iy = (double)tf[1] + ppy * (double)tf[3];
  const double cx = 0.5 * (guess[0] + guess[2]), cy = 0.5 * (guess[1] + guess[3]);
  const double ddx = ix - cx, ddy = iy - cy;
  tk.pos[0] = guess[0] + ddx; tk.pos[1] = guess[1] + ddy; tk.pos[2] = guess[2] + ddx; tk.pos[3] = guess[3] + ddy;
  target_hat(ppx, ppy, G);
  std::vector<double> bs(NPIX, 0.0);
  for (int ch = 0; ch < NCH; ++ch)
    for (int i = 0; i < NPIX; ++i) {
      const cd f = F[(size_t)ch * NPIX + i];
      tk.A[(size_t)ch * NPIX + i] = (1 - NU) * tk.A[(size_t)ch * NPIX + i] + NU * (G[i] * f);
      bs[i] += std::norm(f);
    }
  for (int i = 0; i < NPIX; ++i) tk.B[i] = (1 - NU) * tk.B[i] + NU * bs[i];
  if (bk.use_scale) {
    std::vector<cd> Fs;
    scale_space(rgb, H, W, tk.pos, Fs);
    const Tables& tb = tables();
    cd Rs[NSC];
    for (int m = 0; m < NSC; ++m) {
      cd acc(0, 0);
      for (int j = 0; j < SF; ++j) acc += Fs[(size_t)j * NSC + m] * std::conj(tk.As[(size_t)j * NSC + m]);
      Rs[m] = acc / (tk.Bs[m] + 0.001);
    }
    double r[NSC];
    for (int k = 0; k < NSC; ++k) {
      cd acc(0, 0);
      for (int m = 0; m < NSC; ++m) acc += Rs[m] * std::conj(tb.tw32[(k * m) & (NSC - 1)]);
      r[k] = acc.real() / NSC;
    }
    int pk = 0;
    for (int k = 1; k < NSC; ++k)
      if (r[k] > r[pk]) pk = k;
    double p = pk;
    if (pk > 0 && pk < NSC - 1) {
      const double d = r[pk - 1] - 2 * r[pk] + r[pk + 1];
      if (d != 0) p += std::min(1.0, std::max(-1.0, 0.5 * (r[pk - 1] - r[pk + 1]) / d));   // stays inside [pk - 1, pk + 1] (dlib lagrange_poly_min_extrap)
    }
    const double f = std::pow(1.020, p - NSC / 2);
    const double ccx = 0.5 * (tk.pos[0] + tk.pos[2]), ccy = 0.5 * (tk.pos[1] + tk.pos[3]);
    const double nhw = 0.5 * (tk.pos[2] - tk.pos[0]) * f, nhh = 0.5 * (tk.pos[3] - tk.pos[1]) * f;
    tk.pos[0] = ccx - nhw; tk.pos[1] = ccy - nhh; tk.pos[2] = ccx + nhw; tk.pos[3] = ccy + nhh;
    cd Gs[NSC];
    scale_target_hat(p, Gs);
    double bsum[NSC] = {0};
    for (int j = 0; j < SF; ++j)
      for (int m = 0; m < NSC; ++m) {
        const cd fv = Fs[(size_t)j * NSC + m];
        tk.As[(size_t)j * NSC + m] = (1 - 0.025) * tk.As[(size_t)j * NSC + m] + 0.025 * (Gs[m] * fv);
        bsum[m] += std::norm(fv);
      }
    for (int m = 0; m < NSC; ++m) tk.Bs[m] = (1 - 0.025) * tk.Bs[m] + 0.025 * bsum[m];
  }
  return psr;
}
}  // namespace

extern "C" void* pvc_trackers_create(int capacity, int use_scale) {
  Bank* b = new Bank();
  b->tr.resize(capacity);
  b->use_scale = use_scale != 0;
  return b;
}
extern "C" void pvc_trackers_destroy(void* bank) { delete static_cast<Bank*>(bank); }

extern "C" int pvc_trackers_start(void* bank, const uint8_t* rgb, int H, int W, const int* ids, const double* rects, int n) {
  Bank* b = static_cast<Bank*>(bank);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n; ++i) start_one(*b, b->tr[ids[i]], rgb, H, W, rects + 4 * i);
  return 0;
}
extern "C" int pvc_trackers_update(void* bank, const uint8_t* frames, int H, int W, const int* frame_idx, const int* ids, int n,
                                   double* psr) {
  Bank* b = static_cast<Bank*>(bank);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < n; ++i) {
    const uint8_t* rgb = frames + (size_t)(frame_idx ? frame_idx[i] : 0) * H * W * 3;
    psr[i] = update_one(*b, b->tr[ids[i]], rgb, H, W);
  }
  return 0;
}
extern "C" int pvc_trackers_position(void* bank, int id, double* ltrb) {
  Bank* b = static_cast<Bank*>(bank);
  memcpy(ltrb, b->tr[id].pos, sizeof(double) * 4);
  return 0;
}
