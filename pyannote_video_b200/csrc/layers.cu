// layers.cu — the bandwidth-bound kernels around the srgemm convolutions:
//   pack_gathered   RGBA u8 image -> first-layer "gathered" bf16 rows (normalise (v-mean)/256)
//   maxpool3x3s2    dlib max_pool<3,3,2,2> of the embedder (after conv1)
//   avgpool_skip    dlib avg_pool<2,2,2,2> skip path of ares_down blocks (reads the parity planes)
//   embed_head      avg_pool_everything + fc_no_bias<128>
// Reference call site of the whole network: face_recognition_.compute_face_descriptor,
// pyannote/video/face/face.py:74-75 (dlib anet_type).
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

__device__ __forceinline__ long long row_of(const PvRowMap& m, uint32_t n, uint32_t y, uint32_t x) {
  if (m.kind == 2) {   // faces side by side: image g = n / F holds faces g*F .. g*F+F-1 at column pitch px; py = rows per image
    const uint32_t F = (uint32_t)m.img, g = n / F, f = n - g * F;
    return ((long long)g * m.py + y) * m.w + (long long)f * m.px + x;
  }
  const uint32_t Y = y + m.py, X = x + m.px;
  if (m.kind == 0) return (long long)n * m.img + (long long)Y * m.w + X;
  const uint32_t plane = ((Y & 1u) << 1) | (X & 1u);
  return (long long)plane * m.plane_rows + (long long)n * m.img + (long long)(Y >> 1) * m.w + (X >> 1);
}

// ---------------------------------------------------------------------------------------------
// pack_gathered: one thread per output row (ph, n, i, j):
//   row[k*3 + c] = A(n, 2i+ph, 2j+k) ? (img[n, 2i+ph, 2j+k, c] - mean[c]) / 256 : 0    k < kw
// img is RGBA u8 [B, H, W, 4]; A == 0 marks pyramid padding (normalised value exactly 0).
// ---------------------------------------------------------------------------------------------
template <int KW, int COLS>
__global__ void pack_gathered_kernel(const uchar4* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H,
                                     int W, int Hq, int Wq, long long layout_plane_rows, float m0, float m1,
                                     float m2) {
  const long long plane_rows = (long long)B * Hq * Wq;  // rows of the B images actually present
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * plane_rows) return;
  const int ph = (int)(idx / plane_rows);
  long long r = idx - (long long)ph * plane_rows;
  const long long orow = (long long)ph * layout_plane_rows + r;
  const int n = (int)(r / ((long long)Hq * Wq));
  r -= (long long)n * Hq * Wq;
  const int i = (int)(r / Wq);
  const int j = (int)(r - (long long)i * Wq);
  const int y = 2 * i + ph;
  __align__(16) __nv_bfloat16 v[COLS];
#pragma unroll
  for (int k = 0; k < COLS; ++k) v[k] = __float2bfloat16(0.f);
  if (y < H) {
    const uchar4* src = img + ((long long)n * H + y) * W;
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      const int x = 2 * j + k;
      if (x < W) {
        const uchar4 p = src[x];
        if (p.w) {
          v[k * 3 + 0] = __float2bfloat16(__fmul_rn(__fsub_rn((float)p.x, m0), 0.00390625f));
          v[k * 3 + 1] = __float2bfloat16(__fmul_rn(__fsub_rn((float)p.y, m1), 0.00390625f));
          v[k * 3 + 2] = __float2bfloat16(__fmul_rn(__fsub_rn((float)p.z, m2), 0.00390625f));
        }
      }
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out + orow * COLS);
  const uint4* s = reinterpret_cast<const uint4*>(v);
#pragma unroll
  for (int q = 0; q < COLS / 8; ++q) dst[q] = s[q];
}

// ---------------------------------------------------------------------------------------------
// plane_to_pixrows: RGBA u8 [B,H,W,4] -> normalised bf16 RGBX pixels split by row parity
//   out[((ph*Bcap + n)*Hq + i)*W + x] = A ? ((r,g,b) - mean)/256 : 0, X = 0     (y = 2i + ph)
// The first conv then reads 8-pixel runs of this buffer in place (overlapping TMA rows).
// ---------------------------------------------------------------------------------------------
__global__ void plane_to_pixrows_kernel(const uchar4* __restrict__ img, uint2* __restrict__ out, int B, int H, int W,
                                        int Hq, long long plane_px, float m0, float m1, float m2) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = 2ll * B * Hq * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  long long r = idx / W;
  const int i = (int)(r % Hq);
  r /= Hq;
  const int n = (int)(r % B);
  const int ph = (int)(r / B);
  const int y = 2 * i + ph;
  uint2 o = make_uint2(0u, 0u);
  if (y < H) {
    const uchar4 p = img[((long long)n * H + y) * W + x];
    if (p.w) {
      o.x = pv_pack_bf16x2(__fmul_rn(__fsub_rn((float)p.x, m0), 0.00390625f), __fmul_rn(__fsub_rn((float)p.y, m1), 0.00390625f));
      o.y = pv_pack_bf16x2(__fmul_rn(__fsub_rn((float)p.z, m2), 0.00390625f), 0.f);
    }
  }
  out[(long long)ph * plane_px + ((long long)n * Hq + i) * W + x] = o;
}

// ---------------------------------------------------------------------------------------------
// maxpool 3x3 stride 2 pad 0 on padded-layout (pad 0) bf16 [B, H, W, C]; 8 channels per thread
// ---------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B,
                                    int H, int W, int C, int OH, int OW, PvRowMap dst) {
  const int c8 = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * c8;
  if (idx >= total) return;
  const int cc = (int)(idx % c8);
  long long r = idx / c8;
  const int ox = (int)(r % OW);
  r /= OW;
  const int oy = (int)(r % OH);
  const int n = (int)(r / OH);
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx) {
      const int y = 2 * oy + dy, x = 2 * ox + dx;
      if (y < H && x < W) {
        const uint4 q = *reinterpret_cast<const uint4*>(in + (((long long)n * H + y) * W + x) * C + cc * 8);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
          m[2 * k] = fmaxf(m[2 * k], __bfloat162float(h.x));
          m[2 * k + 1] = fmaxf(m[2 * k + 1], __bfloat162float(h.y));
        }
      }
    }
  uint4 o;
  o.x = pv_pack_bf16x2(m[0], m[1]);
  o.y = pv_pack_bf16x2(m[2], m[3]);
  o.z = pv_pack_bf16x2(m[4], m[5]);
  o.w = pv_pack_bf16x2(m[6], m[7]);
  *reinterpret_cast<uint4*>(out + row_of(dst, n, oy, ox) * dst.cols + cc * 8) = o;
}

// ---------------------------------------------------------------------------------------------
// avgpool 2x2 stride 2 pad 0 of a parity-layout tensor (the four planes ARE the four taps),
// channels zero-extended Cin -> Cout.  Writes the skip tensor S and relu(S) into the block output
// (the conv epilogue later overwrites the cells the conv branch covers: dlib add_prev zero-extends
// the smaller operand).
// ---------------------------------------------------------------------------------------------
__global__ void avgpool_skip_kernel(const __nv_bfloat16* __restrict__ in, int Cin, long long in_plane_rows,
                                    int in_hq, int in_wq, __nv_bfloat16* __restrict__ skip,
                                    __nv_bfloat16* __restrict__ out, int B, int OH, int OW, int Cout, PvRowMap dst) {
  const int c8 = Cout / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * OH * OW * c8;
  if (idx >= total) return;
  const int cc = (int)(idx % c8);
  long long r = idx / c8;
  const int ox = (int)(r % OW);
  r /= OW;
  const int oy = (int)(r % OH);
  const int n = (int)(r / OH);
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (cc * 8 < Cin) {
    const long long base = ((long long)n * in_hq + oy) * in_wq + ox;
    for (int p = 0; p < 4; ++p) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + ((long long)p * in_plane_rows + base) * Cin + cc * 8);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        a[2 * k] = __fadd_rn(a[2 * k], __bfloat162float(h.x));
        a[2 * k + 1] = __fadd_rn(a[2 * k + 1], __bfloat162float(h.y));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __fmul_rn(a[k], 0.25f);
  }
  const long long drow = row_of(dst, n, oy, ox) * dst.cols + cc * 8;
  uint4 o;
  o.x = pv_pack_bf16x2(a[0], a[1]);
  o.y = pv_pack_bf16x2(a[2], a[3]);
  o.z = pv_pack_bf16x2(a[4], a[5]);
  o.w = pv_pack_bf16x2(a[6], a[7]);
  *reinterpret_cast<uint4*>(skip + drow) = o;
  uint4 p;
  p.x = pv_pack_bf16x2(fmaxf(a[0], 0.f), fmaxf(a[1], 0.f));
  p.y = pv_pack_bf16x2(fmaxf(a[2], 0.f), fmaxf(a[3], 0.f));
  p.z = pv_pack_bf16x2(fmaxf(a[4], 0.f), fmaxf(a[5], 0.f));
  p.w = pv_pack_bf16x2(fmaxf(a[6], 0.f), fmaxf(a[7], 0.f));
  *reinterpret_cast<uint4*>(out + drow) = p;
}

// ---------------------------------------------------------------------------------------------
// "packed rows" tensors of the embedder's levels 4 / 3 (csrc/rsconv.cu): [G images][H][Wp][C] bf16, an image row holds F
// faces side by side at column pitch P = W + 1 (one zero column between neighbours = the 3x3 convs' padding).
//   pr_avgpool: dlib avg_pool<2,2,2,2> of every face (the ares_down skip path): in [G,H,Wp,Cin] -> out [G,OH,OWp,Cout],
//               channels >= Cin written as zeros (add_prev zero-extends), gap columns untouched (they stay zero).
//   pr_unpack:  packed rows -> rows of a PvRowMap (hand-over to the srgemm layers of level 2).
// ---------------------------------------------------------------------------------------------
__global__ void pr_avgpool_kernel(const __nv_bfloat16* __restrict__ in, int H, int Wp, int Cin, int P, __nv_bfloat16* __restrict__ out,
                                  int G, int F, int OH, int OW, int OWp, int OP, int Cout) {
  const int c8 = Cout / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)G * F * OH * OW * c8;
  if (idx >= total) return;
  const int cc = (int)(idx % c8);
  long long r = idx / c8;
  const int ox = (int)(r % OW);
  r /= OW;
  const int oy = (int)(r % OH);
  r /= OH;
  const int f = (int)(r % F);
  const int g = (int)(r / F);
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (cc * 8 < Cin) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int y = 2 * oy + (t >> 1), x = f * P + 2 * ox + (t & 1);
      const uint4 q = *reinterpret_cast<const uint4*>(in + (((long long)g * H + y) * Wp + x) * Cin + cc * 8);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        a[2 * k] = __fadd_rn(a[2 * k], __bfloat162float(h.x));
        a[2 * k + 1] = __fadd_rn(a[2 * k + 1], __bfloat162float(h.y));
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = __fmul_rn(a[k], 0.25f);
  }
  uint4 o;
  o.x = pv_pack_bf16x2(a[0], a[1]);
  o.y = pv_pack_bf16x2(a[2], a[3]);
  o.z = pv_pack_bf16x2(a[4], a[5]);
  o.w = pv_pack_bf16x2(a[6], a[7]);
  *reinterpret_cast<uint4*>(out + (((long long)g * OH + oy) * OWp + f * OP + ox) * Cout + cc * 8) = o;
}

__global__ void pr_unpack_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int Wp, int P, int F, int C, int B,
                                 __nv_bfloat16* __restrict__ out, PvRowMap dst) {
  const int c8 = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W * c8;
  if (idx >= total) return;
  const int cc = (int)(idx % c8);
  long long r = idx / c8;
  const int x = (int)(r % W);
  r /= W;
  const int y = (int)(r % H);
  const int n = (int)(r / H);
  const int g = n / F, f = n - g * F;
  const uint4 q = *reinterpret_cast<const uint4*>(in + (((long long)g * H + y) * Wp + f * P + x) * C + cc * 8);
  *reinterpret_cast<uint4*>(out + row_of(dst, n, y, x) * dst.cols + cc * 8) = q;
}

// ---------------------------------------------------------------------------------------------
// embed head: global average over HxW (padded layout pad 0), then y = fc[128,256] . g
// one CTA of 256 threads per face
// ---------------------------------------------------------------------------------------------
__global__ void embed_head_kernel(const __nv_bfloat16* __restrict__ in, int HW, int C, const float* __restrict__ fc,
                                  float* __restrict__ out, int D) {
  __shared__ float g[256];
  const int n = blockIdx.x;
  const int c = threadIdx.x;
  if (c < C) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s = __fadd_rn(s, __bfloat162float(in[((long long)n * HW + p) * C + c]));
    g[c] = __fdiv_rn(s, (float)HW);
  }
  __syncthreads();
  // 2 threads per output would be faster; the layer is 32k MAC per face — irrelevant.
  if (c < D) {
    float acc = 0.f;
    const float* w = fc + (long long)c * C;
    for (int k = 0; k < C; ++k) acc = fmaf(w[k], g[k], acc);
    out[(long long)n * D + c] = acc;
  }
}

}  // namespace

extern "C" int pv_pack_gathered(const void* rgba, void* out, int B, int H, int W, int kw, int64_t layout_plane_rows,
                                const float* mean_host, void* stream) {
  PV_REQUIRE(rgba && out && mean_host, "pv_pack_gathered: null argument");
  PV_REQUIRE(kw == 5 || kw == 7, "pv_pack_gathered: kw=%d (5 or 7)", kw);
  const int Hq = (H + 1) / 2, Wq = (W + 1) / 2;
  const long long rows = 2ll * B * Hq * Wq;
  const int threads = 256;
  const long long blocks = (rows + threads - 1) / threads;
  PV_REQUIRE(blocks < (1ll << 31), "pv_pack_gathered: too many rows");
  PV_REQUIRE(layout_plane_rows >= (long long)B * Hq * Wq, "pv_pack_gathered: layout_plane_rows too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (kw == 5)
    pack_gathered_kernel<5, 16><<<(unsigned)blocks, threads, 0, s>>>(
        static_cast<const uchar4*>(rgba), static_cast<__nv_bfloat16*>(out), B, H, W, Hq, Wq, layout_plane_rows, mean_host[0],
        mean_host[1], mean_host[2]);
  else
    pack_gathered_kernel<7, 32><<<(unsigned)blocks, threads, 0, s>>>(
        static_cast<const uchar4*>(rgba), static_cast<__nv_bfloat16*>(out), B, H, W, Hq, Wq, layout_plane_rows, mean_host[0],
        mean_host[1], mean_host[2]);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_plane_to_pixrows(const void* rgba, void* out, int B, int H, int W, int64_t layout_plane_rows,
                                   const float* mean_host, void* stream) {
  PV_REQUIRE(rgba && out && mean_host, "pv_plane_to_pixrows: null argument");
  PV_REQUIRE(W % 2 == 0, "pv_plane_to_pixrows: W=%d must be even", W);
  const int Hq = (H + 1) / 2;
  const long long total = 2ll * B * Hq * W;
  PV_REQUIRE(layout_plane_rows * 2 >= (long long)B * Hq * W, "pv_plane_to_pixrows: layout too small");
  const int threads = 256;
  plane_to_pixrows_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uchar4*>(rgba), static_cast<uint2*>(out), B, H, W, Hq, layout_plane_rows * 2, mean_host[0],
      mean_host[1], mean_host[2]);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_maxpool3x3s2(const void* in, void* out, int B, int H, int W, int C, const PvRowMap* dst,
                               void* stream) {
  PV_REQUIRE(in && out && dst, "pv_maxpool3x3s2: null argument");
  PV_REQUIRE(C % 8 == 0 && dst->cols >= C, "pv_maxpool3x3s2: C=%d", C);
  const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
  const long long total = (long long)B * OH * OW * (C / 8);
  const int threads = 256;
  maxpool3x3s2_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), B, H, W, C, OH, OW, *dst);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_avgpool_skip(const void* in, int Cin, int64_t in_plane_rows, int in_hq, int in_wq, void* skip,
                               void* out, int B, int OH, int OW, int Cout, const PvRowMap* dst, void* stream) {
  PV_REQUIRE(in && skip && out && dst, "pv_avgpool_skip: null argument");
  PV_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cout >= Cin && dst->cols == Cout, "pv_avgpool_skip: channels");
  const long long total = (long long)B * OH * OW * (Cout / 8);
  const int threads = 256;
  avgpool_skip_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), Cin, in_plane_rows, in_hq, in_wq, static_cast<__nv_bfloat16*>(skip),
      static_cast<__nv_bfloat16*>(out), B, OH, OW, Cout, *dst);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_pr_avgpool(const void* in, int G, int F, int H, int W, int Wp, int Cin, void* out, int OWp, int Cout,
                             void* stream) {
  PV_REQUIRE(in && out, "pv_pr_avgpool: null argument");
  PV_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cout >= Cin && (W + 1) % 2 == 0, "pv_pr_avgpool: channels / face pitch");
  const int OH = (H - 2) / 2 + 1, OW = (W - 2) / 2 + 1, P = W + 1, OP = P / 2;
  PV_REQUIRE(F * P <= Wp && F * OP <= OWp, "pv_pr_avgpool: pitches");
  const long long total = (long long)G * F * OH * OW * (Cout / 8);
  pr_avgpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), H, Wp, Cin, P, static_cast<__nv_bfloat16*>(out), G, F, OH, OW, OWp, OP, Cout);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_pr_unpack(const void* in, int B, int F, int H, int W, int Wp, int C, void* out, const PvRowMap* dst,
                            void* stream) {
  PV_REQUIRE(in && out && dst, "pv_pr_unpack: null argument");
  PV_REQUIRE(C % 8 == 0 && dst->cols >= C && F * (W + 1) <= Wp, "pv_pr_unpack: geometry");
  const long long total = (long long)B * H * W * (C / 8);
  pr_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), H, W, Wp, W + 1, F, C, B, static_cast<__nv_bfloat16*>(out), *dst);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_embed_head(const void* in, int B, int HW, int C, const float* fc, float* out, int D, void* stream) {
  PV_REQUIRE(in && fc && out, "pv_embed_head: null argument");
  PV_REQUIRE(C <= 256 && D <= 256, "pv_embed_head: C=%d D=%d", C, D);
  embed_head_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(in), HW, C,
                                                                      fc, out, D);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
