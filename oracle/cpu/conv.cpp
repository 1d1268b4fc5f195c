// ORACLE (test infrastructure) — fp32 convolution stacks of the C++ CPU restatement (see pvcpu.h).
//
// Restates dlib's CPU DNN path for the two networks on the face path (con_ + affine_ + relu_, max_pool_, avg_pool_,
// add_prev_, fc_; SURVEY.md App. A.1 / A.4, mirrored from oracle/nets.py) as a blocked direct convolution:
// activations are NHWC float32 with channels padded to 16 and a zero halo (so the inner loop never tests bounds),
// output channels are vectorised 16 wide (GCC vector extensions: AVX-512 / AVX2 / SSE / NEON, whatever -march gives),
// T output pixels x NB channel blocks are accumulated in registers, rows are spread over OpenMP threads.
// This file is compiled with FMA contraction allowed; results agree with torch's fp32 conv2d within float rounding.
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "pvcpu.h"
#include "tensor.h"

namespace pvc {

static inline float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

struct ConvLayer {
  int cout, cin, k, stride, pad, relu, nb;   // nb = ceil(cout/16)
  std::vector<float> w;                      // [k][k][cin][nb*16]
  std::vector<float> w_c1;                   // cout == 1: [k][k][cin_p16]
  std::vector<float> scale, shift;           // [nb*16]
};

struct Net {
  std::vector<ConvLayer> layers;
};

// ------------------------------------------------------------------------------------------------------------------
template <int NB, int T>
static void conv_rows(const Tensor& in, Tensor& out, const ConvLayer& L, bool round_bf16, int n, int y) {
  const int s = L.stride, k = L.k, pad = L.pad;
  const int OW = out.W;
  const long row_stride = in.row_stride();
  const int px_unit = in.Cp;          // floats per input pixel
  const int px_stride = px_unit * s;  // floats between the inputs of adjacent outputs
  for (int x0 = 0; x0 < OW; x0 += T) {
    v16 acc[T][NB];
    for (int t = 0; t < T; ++t)
      for (int b = 0; b < NB; ++b) acc[t][b] = v16_zero();
    const float* ip = in.at(n, y * s - pad, x0 * s - pad);
    for (int kh = 0; kh < k; ++kh) {
      const float* irow = ip + (long)kh * row_stride;
      for (int kw = 0; kw < k; ++kw) {
        const float* ipx = irow + (long)kw * px_unit;
        const float* wp = L.w.data() + ((long)(kh * k + kw) * L.cin) * (NB * 16);
        for (int ci = 0; ci < L.cin; ++ci) {
          v16 wv[NB];
#pragma GCC unroll 4
          for (int b = 0; b < NB; ++b) wv[b] = *reinterpret_cast<const v16u*>(wp + (long)ci * NB * 16 + b * 16);
#pragma GCC unroll 16
          for (int t = 0; t < T; ++t) {
            const float a = ipx[(long)t * px_stride + ci];
#pragma GCC unroll 4
            for (int b = 0; b < NB; ++b) acc[t][b] += wv[b] * a;
          }
        }
      }
    }
    const int nt = std::min(T, OW - x0);
    for (int t = 0; t < nt; ++t) {
      float* op = out.at(n, y, x0 + t);
      for (int b = 0; b < NB; ++b) {
        v16 v = acc[t][b] * *reinterpret_cast<const v16u*>(L.scale.data() + b * 16) +
                *reinterpret_cast<const v16u*>(L.shift.data() + b * 16);
        if (L.relu) v = v16_max0(v);
        if (round_bf16) {
          float tmp[16];
          memcpy(tmp, &v, 64);
          for (int i = 0; i < 16; ++i) tmp[i] = bf16_round(tmp[i]);
          memcpy(&v, tmp, 64);
        }
        *reinterpret_cast<v16u*>(op + b * 16) = v;
      }
    }
  }
}

// single output channel (the detector's 9x9 layer): vectorise over input channels instead
static void conv_rows_c1(const Tensor& in, Tensor& out, const ConvLayer& L, int n, int y, float* dst_row) {
  const int k = L.k, pad = L.pad;
  const int OW = out.W;
  const long row_stride = in.row_stride();
  const int cb = in.Cp / 16;
  constexpr int T = 8;
  for (int x0 = 0; x0 < OW; x0 += T) {
    v16 acc[T];
    for (int t = 0; t < T; ++t) acc[t] = v16_zero();
    const float* ip = in.at(n, y - pad, x0 - pad);
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const float* ipx = ip + (long)kh * row_stride + (long)kw * in.Cp;
        const float* wp = L.w_c1.data() + (long)(kh * k + kw) * in.Cp;
        for (int c = 0; c < cb; ++c) {
          const v16 wv = *reinterpret_cast<const v16u*>(wp + c * 16);
#pragma GCC unroll 8
          for (int t = 0; t < T; ++t) acc[t] += *reinterpret_cast<const v16u*>(ipx + (long)t * in.Cp + c * 16) * wv;
        }
      }
    const int nt = std::min(T, OW - x0);
    for (int t = 0; t < nt; ++t) {
      float tmp[16];
      memcpy(tmp, &acc[t], 64);
      float sum = 0.f;
      for (int i = 0; i < 16; ++i) sum += tmp[i];
      float v = sum * L.scale[0] + L.shift[0];
      if (L.relu) v = v > 0.f ? v : 0.f;
      dst_row[x0 + t] = v;
    }
  }
}

template <int NB>
static void conv_layer_nb(const Tensor& in, Tensor& out, const ConvLayer& L, bool round_bf16) {
  const int N = out.N, OH = out.H;
#pragma omp parallel for collapse(2) schedule(dynamic, 2)
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < OH; ++y) {
      if (NB == 1) conv_rows<1, 14>(in, out, L, round_bf16, n, y);
      else if (NB == 2) conv_rows<2, 12>(in, out, L, round_bf16, n, y);
      else if (NB == 3) conv_rows<3, 8>(in, out, L, round_bf16, n, y);
      else conv_rows<4, 6>(in, out, L, round_bf16, n, y);
    }
}

// out must be allocated by the caller: [N][OH][OW][cout] with the halo the NEXT layer needs
void conv_forward(const Tensor& in, Tensor& out, const ConvLayer& L, bool round_bf16) {
  if (L.cout == 1) {
    const int N = out.N, OH = out.H;
#pragma omp parallel for collapse(2) schedule(dynamic, 2)
    for (int n = 0; n < N; ++n)
      for (int y = 0; y < OH; ++y) {
        std::vector<float> row(out.W + 8);
        conv_rows_c1(in, out, L, n, y, row.data());
        for (int x = 0; x < out.W; ++x) out.at(n, y, x)[0] = row[x];
      }
    return;
  }
  // channel blocks beyond 4 (64 channels) are processed in groups of 4 by offsetting weights / outputs
  if (L.nb <= 4) {
    switch (L.nb) {
      case 1: conv_layer_nb<1>(in, out, L, round_bf16); break;
      case 2: conv_layer_nb<2>(in, out, L, round_bf16); break;
      case 3: conv_layer_nb<3>(in, out, L, round_bf16); break;
      default: conv_layer_nb<4>(in, out, L, round_bf16); break;
    }
    return;
  }
  // wide layers (embedder levels with 128 / 256 channels): slice the output channels into groups of 64
  for (int g0 = 0; g0 < L.nb; g0 += 4) {
    const int gb = std::min(4, L.nb - g0);
    ConvLayer S;
    S.cout = gb * 16; S.cin = L.cin; S.k = L.k; S.stride = L.stride; S.pad = L.pad; S.relu = L.relu; S.nb = gb;
    S.w.resize((size_t)L.k * L.k * L.cin * gb * 16);
    for (long r = 0; r < (long)L.k * L.k * L.cin; ++r)
      memcpy(&S.w[r * gb * 16], &L.w[r * L.nb * 16 + g0 * 16], sizeof(float) * gb * 16);
    S.scale.assign(L.scale.begin() + g0 * 16, L.scale.begin() + (g0 + gb) * 16);
    S.shift.assign(L.shift.begin() + g0 * 16, L.shift.begin() + (g0 + gb) * 16);
    Tensor view = out.channel_view(g0 * 16, gb * 16);
    switch (gb) {
      case 1: conv_layer_nb<1>(in, view, S, round_bf16); break;
      case 2: conv_layer_nb<2>(in, view, S, round_bf16); break;
      case 3: conv_layer_nb<3>(in, view, S, round_bf16); break;
      default: conv_layer_nb<4>(in, view, S, round_bf16); break;
    }
  }
}

}  // namespace pvc

using namespace pvc;

extern "C" void* pvc_net_create(void) { return new Net(); }
extern "C" void pvc_net_destroy(void* net) { delete static_cast<Net*>(net); }

extern "C" int pvc_net_add_conv(void* net, int cout, int cin, int k, int stride, int pad, const float* w, const float* scale,
                                const float* shift, int relu, int round_bf16) {
  Net* N = static_cast<Net*>(net);
  ConvLayer L;
  L.cout = cout; L.cin = cin; L.k = k; L.stride = stride; L.pad = pad; L.relu = relu;
  L.nb = (cout + 15) / 16;
  auto wv = [&](int co, int ci, int kh, int kw) {
    float v = w[(((long)co * cin + ci) * k + kh) * k + kw];
    return round_bf16 ? bf16_round(v) : v;
  };
  if (cout == 1) {
    const int cp = (cin + 15) / 16 * 16;
    L.w_c1.assign((size_t)k * k * cp, 0.f);
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw)
        for (int ci = 0; ci < cin; ++ci) L.w_c1[(size_t)(kh * k + kw) * cp + ci] = wv(0, ci, kh, kw);
  } else {
    L.w.assign((size_t)k * k * cin * L.nb * 16, 0.f);
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw)
        for (int ci = 0; ci < cin; ++ci)
          for (int co = 0; co < cout; ++co) L.w[(((size_t)(kh * k + kw) * cin + ci) * L.nb * 16) + co] = wv(co, ci, kh, kw);
  }
  L.scale.assign(L.nb * 16, 0.f);
  L.shift.assign(L.nb * 16, 0.f);
  for (int co = 0; co < cout; ++co) { L.scale[co] = scale[co]; L.shift[co] = shift[co]; }
  N->layers.push_back(std::move(L));
  return (int)N->layers.size() - 1;
}

extern "C" int pvc_detector_out_size(void* net, int n) {
  Net* N = static_cast<Net*>(net);
  for (const ConvLayer& L : N->layers) n = (n + 2 * L.pad - L.k) / L.stride + 1;
  return n;
}

extern "C" int pvc_detector_forward(void* net, const uint8_t* plane, int Hp, int Wp, const float* mean3, float pixel_scale,
                                    int round_bf16, float* scores, int* oh, int* ow) {
  Net* N = static_cast<Net*>(net);
  const int nl = (int)N->layers.size();
  if (nl == 0) return -1;
  // input: (v - mean) * scale inside tiles (A != 0), 0 in the padding; 3 channels padded to 4
  Tensor cur(1, Hp, Wp, 3, N->layers[0].pad, /*cp=*/4);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < Hp; ++y) {
    const uint8_t* src = plane + (long)y * Wp * 4;
    for (int x = 0; x < Wp; ++x) {
      float* d = cur.at(0, y, x);
      const bool in_tile = src[4 * x + 3] != 0;
      for (int c = 0; c < 3; ++c) {
        float v = in_tile ? ((float)src[4 * x + c] - mean3[c]) * pixel_scale : 0.f;
        d[c] = round_bf16 ? bf16_round(v) : v;
      }
      d[3] = 0.f;
    }
  }
  for (int i = 0; i < nl; ++i) {
    const ConvLayer& L = N->layers[i];
    const int OH = (cur.H + 2 * L.pad - L.k) / L.stride + 1, OW = (cur.W + 2 * L.pad - L.k) / L.stride + 1;
    const int next_pad = i + 1 < nl ? N->layers[i + 1].pad : 0;
    Tensor out(1, OH, OW, L.cout, next_pad);
    conv_forward(cur, out, L, round_bf16 && i + 1 < nl);
    cur = std::move(out);
  }
  *oh = cur.H;
  *ow = cur.W;
  for (int y = 0; y < cur.H; ++y)
    for (int x = 0; x < cur.W; ++x) scores[(long)y * cur.W + x] = cur.at(0, y, x)[0];
  return 0;
}

// ---- embedder graph (oracle/nets.py embed_forward) -------------------------------------------------------------
static void round_tensor(Tensor& t) {
#pragma omp parallel for collapse(2)
  for (int n = 0; n < t.N; ++n)
    for (int y = 0; y < t.H; ++y)
      for (int x = 0; x < t.W; ++x) {
        float* p = t.at(n, y, x);
        for (int c = 0; c < t.C; ++c) p[c] = bf16_round(p[c]);
      }
}

static Tensor maxpool3s2(const Tensor& in, int next_pad) {
  const int OH = (in.H - 3) / 2 + 1, OW = (in.W - 3) / 2 + 1;
  Tensor out(in.N, OH, OW, in.C, next_pad);
#pragma omp parallel for collapse(2)
  for (int n = 0; n < in.N; ++n)
    for (int y = 0; y < OH; ++y)
      for (int x = 0; x < OW; ++x) {
        float* o = out.at(n, y, x);
        for (int c = 0; c < in.C; ++c) {
          float m = -INFINITY;
          for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) m = std::max(m, in.at(n, 2 * y + dy, 2 * x + dx)[c]);
          o[c] = m;
        }
      }
  return out;
}

static Tensor avgpool2s2(const Tensor& in) {
  const int OH = (in.H - 2) / 2 + 1, OW = (in.W - 2) / 2 + 1;
  Tensor out(in.N, OH, OW, in.C, 0);
  for (int n = 0; n < in.N; ++n)
    for (int y = 0; y < OH; ++y)
      for (int x = 0; x < OW; ++x) {
        float* o = out.at(n, y, x);
        for (int c = 0; c < in.C; ++c)
          o[c] = (in.at(n, 2 * y, 2 * x)[c] + in.at(n, 2 * y, 2 * x + 1)[c] + in.at(n, 2 * y + 1, 2 * x)[c] +
                  in.at(n, 2 * y + 1, 2 * x + 1)[c]) * 0.25f;
      }
  return out;
}

// relu(zero_extend(a) + zero_extend(b)) into a tensor with halo `next_pad`
static Tensor add_relu(const Tensor& a, const Tensor& b, int next_pad) {
  const int C = std::max(a.C, b.C), H = std::max(a.H, b.H), W = std::max(a.W, b.W);
  Tensor out(a.N, H, W, C, next_pad);
  for (int n = 0; n < a.N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float* o = out.at(n, y, x);
        for (int c = 0; c < C; ++c) {
          float v = 0.f;
          if (y < a.H && x < a.W && c < a.C) v += a.at(n, y, x)[c];
          if (y < b.H && x < b.W && c < b.C) v += b.at(n, y, x)[c];
          o[c] = v > 0.f ? v : 0.f;
        }
      }
  return out;
}

extern "C" int pvc_embed_forward(void* net, const int* block_down, int n_blocks, const float* fc, const uint8_t* chips, int M,
                                 int S, const float* mean3, float pixel_scale, int round_bf16, float* out) {
  Net* N = static_cast<Net*>(net);
  if ((int)N->layers.size() != 1 + 2 * n_blocks) return -1;
  const bool rb = round_bf16 != 0;
  Tensor x(M, S, S, 3, 0, 4);
  for (int n = 0; n < M; ++n)
    for (int y = 0; y < S; ++y)
      for (int xx = 0; xx < S; ++xx) {
        const uint8_t* src = chips + (((long)n * S + y) * S + xx) * 3;
        float* d = x.at(n, y, xx);
        for (int c = 0; c < 3; ++c) {
          float v = ((float)src[c] - mean3[c]) * pixel_scale;
          d[c] = rb ? bf16_round(v) : v;
        }
        d[3] = 0.f;
      }
  auto conv = [&](const Tensor& in, const ConvLayer& L, int next_pad, bool round_out) {
    const int OH = (in.H + 2 * L.pad - L.k) / L.stride + 1, OW = (in.W + 2 * L.pad - L.k) / L.stride + 1;
    Tensor o(in.N, OH, OW, L.cout, next_pad);
    conv_forward(in, o, L, round_out);
    return o;
  };
  Tensor c1 = conv(x, N->layers[0], 0, rb);
  // the first block decides the halo of the pooled tensor
  auto halo_for = [&](int blk) { return blk < n_blocks ? (block_down[blk] ? 0 : 1) : 0; };
  Tensor cur = maxpool3s2(c1, halo_for(0));
  for (int i = 0; i < n_blocks; ++i) {
    const ConvLayer& A = N->layers[1 + 2 * i];
    const ConvLayer& B = N->layers[2 + 2 * i];
    Tensor t = conv(cur, A, 1, rb);
    Tensor u = conv(t, B, 0, false);
    Tensor nxt;
    if (block_down[i]) {
      Tensor s = avgpool2s2(cur);
      if (rb) round_tensor(s);
      nxt = add_relu(s, u, halo_for(i + 1));
    } else {
      nxt = add_relu(cur, u, halo_for(i + 1));
    }
    if (rb) round_tensor(nxt);
    cur = std::move(nxt);
  }
  for (int n = 0; n < M; ++n) {
    std::vector<float> g(cur.C, 0.f);
    for (int c = 0; c < cur.C; ++c) {
      float sum = 0.f;
      for (int y = 0; y < cur.H; ++y)
        for (int xx = 0; xx < cur.W; ++xx) sum += cur.at(n, y, xx)[c];
      g[c] = sum / (float)(cur.H * cur.W);
    }
    for (int o = 0; o < 128; ++o) {
      float sum = 0.f;
      for (int c = 0; c < cur.C; ++c) sum += g[c] * fc[(long)o * cur.C + c];
      out[(long)n * 128 + o] = sum;
    }
  }
  return 0;
}
