// c12.cu — detector conv1 + conv2 as ONE row-streaming strip kernel (sm_100a, tcgen05 + TMEM + TMA).
//
// Replaces the first two `con` layers of dlib's CNN/MMOD face detector (con<16,5,5,2,2> and con<32,5,5,2,2>, each
// followed by affine + relu) behind face_detector_(rgb, 1), pyannote/video/face/face.py:66.  Separately
// (csrc/conv1_fused.cu + csrc/rsconv.cu) conv1 writes 232 MB per 1080p frame that conv2 reads straight back, and
// conv1 is paced by its per-tile hand-offs (profiles/README.md).  Here the conv1 activations never leave the SM:
//
//   work item   = 124 conv2 output columns x L conv2 rows of one image
//                 = 252 conv1 columns (two M tiles of 126 outputs) = 512 plane pixels per plane row
//   stage       = a QUAD of 4 plane rows (2 TMA boxes of 256 px x 4 rows, raw RGBA u8) — one hand-off per quad
//   converters  : 8 warps normalise the quad once into bf16 RGB0 pixel rows (8 B / pixel); the conv1 A operand rows
//                 (6-pixel windows at a 2-pixel step = 16 B) OVERLAP in that buffer: no im2col (as in conv1_fused.cu)
//   conv1 MMAs  : per plane row t one "main" MMA (kw 0..3 x RGB0 = K 16) against the filter rows of ALL conv1 rows
//                 that read it side by side (kh = t - 2r: N = 48 for even t, 32 for odd t), plus one "pair" MMA per two
//                 plane rows for kw = 4 (K = 2 x 8) — 8 MMAs per tile and quad; accumulators = a ring of TMEM row
//                 slots (16 columns per conv1 row and tile)
//   conv1 epilogue: 8 warps, TMEM -> affine -> ReLU -> bf16 -> SHARED memory, written directly in the layout of
//                 conv2's A operand: pixel-PAIR rows of 64 B (2 x 16 channels) with the 64-byte swizzle applied by hand
//   conv2 MMAs  : exactly rsconv<16,32,5,5,2>: per conv1 row 5 MMAs (kw) of N = 96 / 64 into a ring of 8 TMEM row
//                 slots of 32 columns
//   conv2 epilogue: 4 warps, TMEM -> affine -> ReLU -> bf16 -> NHWC global (64 B per pixel)
//
// Accumulator rows are handed back CLEARED by the epilogues (tcgen05.st of zeros after the tcgen05.ld), so every MMA
// accumulates and the first contribution to a row needs no MMA of its own.  Every item starts at conv1 ring slot 0
// (per-slot barrier parities are tracked in bit masks), which makes the conv1 schedule of an interior quad a
// compile-time sequence (template on q mod 4): 6 MMAs per tile and quad, ring-wrap splits only when q mod 4 == 0.
//
// TMEM: conv1 tile 0 columns [0,128), tile 1 [128,256), conv2 [256,512): one CTA per SM, persistent.
// Warps: 0 TMA, 1 / 3 conv1 MMA issuers (tile 0 / 1; 3 also owns TMEM), 2 conv2 MMA issuer, 4..11 converters,
// 12..19 conv1 epilogue, 20..23 conv2 epilogue.  All waits are bounded (pv_mbar_wait traps and raises the error flag).
#include <cuda.h>
#include <atomic>
#include <cstdlib>
#include <vector>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kW2 = 124;                 // valid conv2 outputs per strip
constexpr int kT1 = 126;                 // conv1 outputs per M tile (2 tiles per strip)
constexpr int kRawHalf = 256;            // pixels per TMA box row
constexpr int kQuad = 4;                 // plane rows per stage
constexpr int kRawHalfBytes = kQuad * kRawHalf * 4;     // 4 KB
constexpr int kRawStage = 2 * kRawHalfBytes;            // 8 KB
constexpr int kRawRing = 3;
constexpr int kPxRow = (2 * kRawHalf + 8) * 8;          // 4160 B: 520 bf16 RGB0 pixels, the trailing 8 stay zero
constexpr int kPxStage = kQuad * kPxRow;                // 16640 B
constexpr int kPxRing = 3;
constexpr int kA2RowB = 64;                              // conv2 A operand row: one conv1 pixel pair x 16 channels
constexpr int kA2Entry = 8704;                           // 130 pair rows (8320 B) rounded up to 512 B
constexpr int kA2Ring = 8;
constexpr int kNS1 = 8;                                  // conv1 TMEM row slots per tile (16 columns each)
constexpr int kNS2 = 8;                                  // conv2 TMEM row slots (32 columns each)
constexpr int kC1 = 16, kC2 = 32;
constexpr int kW1Bytes = 8 * 512;                        // conv1 weight image: tiles M0 (3 blocks), M1 (2), P (3); block = 16 rows x 32 B
constexpr int kW1M0 = 0, kW1M1 = 3 * 512, kW1P = 5 * 512;
constexpr int kW2Tile0 = 3 * kC2 * 32, kW2Tile1 = 2 * kC2 * 32;          // conv2 weight tiles of parity class 0 / 1
constexpr int kW2Q1Base = 5 * kW2Tile0;
constexpr int kW2Bytes = 5 * (kW2Tile0 + kW2Tile1);      // 25600
constexpr int kThreads = 768;
constexpr int kConvWarps = 8, kEp1Warps = 8, kEp2Warps = 4;

// shared-memory map (offsets from the 1 KB-aligned base)
constexpr int kOffA2 = 0;
constexpr int kOffW2 = kOffA2 + kA2Ring * kA2Entry;
constexpr int kOffW1 = kOffW2 + kW2Bytes;
constexpr int kOffPx = kOffW1 + kW1Bytes;
constexpr int kOffRaw = kOffPx + kPxRing * kPxStage;
constexpr int kOffBar = kOffRaw + kRawRing * kRawStage;
constexpr int kNumBars = 2 * kRawRing + 2 * kPxRing + 4 * kNS1 + 2 * kA2Ring + 2 * kNS2;
constexpr int kOffFl = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffFl + (2 * kC1 + 2 * kC2) * 4 + 16 + 1024;
static_assert(kOffW2 % 1024 == 0 && kOffW1 % 1024 == 0 && kOffPx % 128 == 0 && kOffRaw % 128 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kA2Entry % 512 == 0 && kPxRow % 16 == 0, "alignment");

struct C12Params {
  CUtensorMap raw;         // uint32 [B*Hp rows, Wp px], box 256 x 4
  const uint8_t* w1_img;   // 4096 B (layout: pack_c12_w1 in detconv.py / header)
  const uint8_t* w2_img;   // 25600 B = rsconv<16,32,5,5,2> weight image
  const float* scale1; const float* shift1;   // [16]
  const float* scale2; const float* shift2;   // [32]
  __nv_bfloat16* out;      // [B, OH2, out_pitch, 32]
  int B, Hp, Wp;
  int OH2, OW2, out_pitch;
  int strips, segs, seg_rows;
  int num_items;
  float c0, c1, c2;        // -mean/256
  int* err;
  long long* dbg;          // optional [grid][16] role cycle counters
};

__device__ __forceinline__ void c12_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

// un-swizzled K-major operand (conv1 A): LBO = byte step between the two 8-element K chunks, SBO = 8-row group step
__device__ __forceinline__ uint64_t c12_desc_plain(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}
// swizzled K-major operand: hi word (SBO, version 1, layout type), lo word = (addr >> 4) | LBO 1
__host__ __device__ constexpr uint32_t c12_desc_hi(int sbo_bytes, uint32_t layout) {
  return (uint32_t)((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout << 29);
}

constexpr uint32_t kIdesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);   // + (N >> 3) << 17
constexpr uint32_t kBHi = c12_desc_hi(256, 6u);        // weight tiles: 32-byte rows, 32-byte swizzle, 8 rows = 256 B
constexpr uint32_t kA2Hi = c12_desc_hi(512, 4u);       // conv2 A: 64-byte rows, 64-byte swizzle, 8 rows = 512 B

constexpr uint32_t kPxRow16 = (uint32_t)(kPxRow >> 4);
__host__ __device__ constexpr uint32_t c12_idesc(int n) { return kIdesc0 | ((uint32_t)(n >> 3) << 17); }

// conv1, generic (edge quads): rows [r_first, r_first + nblk) of the item take blocks 0..nblk-1 of a weight tile; clamp to
// [0, r_max], split at the TMEM ring wrap (row r of an item lives in slot r & 7).  All MMAs accumulate.
__device__ __forceinline__ void c1_issue(uint32_t dbase, uint64_t adesc, uint32_t b_lo, int r_first, int nblk, int r_max) {
  const int lo = r_first > 0 ? r_first : 0;
  int hi = r_first + nblk - 1;
  if (hi > r_max) hi = r_max;
  const int n = hi - lo + 1;
  if (n > 0) {
    const uint32_t s_lo = (uint32_t)lo & (uint32_t)(kNS1 - 1);
    const int n1 = n < (int)(kNS1 - s_lo) ? n : (int)(kNS1 - s_lo);
    const int n2 = n - n1;
    const uint64_t b1 = ((uint64_t)kBHi << 32) | (uint64_t)(b_lo + (uint32_t)(lo - r_first) * 32u);
    c12_umma(dbase + s_lo * kC1, adesc, b1, c12_idesc(n1 * kC1), 1u);
    if (n2 > 0) c12_umma(dbase, adesc, b1 + (uint64_t)(n1 * 32), c12_idesc(n2 * kC1), 1u);
  }
}

// conv1, compile-time: rows at ring slots S0 .. S0+NB-1 (mod 8) take blocks 0..NB-1 of the tile at b_lo
template <int S0, int NB>
__device__ __forceinline__ void c1_group(uint32_t dbase, uint64_t adesc, uint32_t b_lo) {
  constexpr int s0 = S0 & (kNS1 - 1);
  constexpr int n1 = (s0 + NB <= kNS1) ? NB : kNS1 - s0;
  constexpr int n2 = NB - n1;
  const uint64_t b1 = ((uint64_t)kBHi << 32) | (uint64_t)b_lo;
  c12_umma(dbase + (uint32_t)(s0 * kC1), adesc, b1, c12_idesc(n1 * kC1), 1u);
  if constexpr (n2 > 0) c12_umma(dbase, adesc, b1 + (uint64_t)(n1 * 32), c12_idesc(n2 * kC1), 1u);
}

__device__ __forceinline__ void c1_open(uint64_t* bar_e, uint32_t& emask, int slot, int* err) {
  pv_mbar_wait(&bar_e[slot], (emask >> slot) & 1u, err, 3);
  emask ^= 1u << slot;
  pv_tc_fence_after();
}

// one interior quad (rows 2q-2 .. 2q+1 all inside the item) of one tile; QM = q & 3 fixes every TMEM slot
template <int QM>
__device__ __forceinline__ void c1_quad_static(uint32_t dbase, uint64_t ad0, uint64_t pd0, uint32_t w1_16, uint64_t* bar_e,
                                               uint64_t* bar_f, uint32_t& emask, int* err) {
  constexpr int b = (2 * QM + 6) & 7;                     // slot of row 2q-2
  c1_open(bar_e, emask, (b + 2) & 7, err);                // even plane row 4q opens row 2q
  c1_group<b, 3>(dbase, ad0, w1_16 + (uint32_t)(kW1M0 >> 4));
  c1_group<b + 1, 2>(dbase, ad0 + kPxRow16, w1_16 + (uint32_t)(kW1M1 >> 4));
  c1_group<b, 3>(dbase, pd0, w1_16 + (uint32_t)(kW1P >> 4));
  pv_umma_commit(&bar_f[b]);                              // row 2q-2 complete
  c1_open(bar_e, emask, (b + 3) & 7, err);                // plane row 4q+2 opens row 2q+1
  c1_group<b + 1, 3>(dbase, ad0 + 2 * kPxRow16, w1_16 + (uint32_t)(kW1M0 >> 4));
  c1_group<b + 2, 2>(dbase, ad0 + 3 * kPxRow16, w1_16 + (uint32_t)(kW1M1 >> 4));
  c1_group<b + 1, 3>(dbase, pd0 + 2 * kPxRow16, w1_16 + (uint32_t)(kW1P >> 4));
  pv_umma_commit(&bar_f[(b + 1) & 7]);                    // row 2q-1 complete
}

// one RGBA u8 pixel -> bf16 RGB0: (v - mean)/256 == fma(v, 2^-8, -mean*2^-8) bit for bit (power-of-two scaling commutes
// with rounding); u8 -> f32 without the conversion unit: prmt builds 0x4B0000vv = 2^23 + v, minus 2^23 is exact.
// alpha 0 = pyramid padding / TMA zero fill -> exact 0
template <int SEL>
__device__ __forceinline__ float c12_u8f(uint32_t u) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(u), "r"(0x4B000000u), "n"(0x7540 | SEL));
  return __fsub_rn(__uint_as_float(r), 8388608.0f);
}
__device__ __forceinline__ void c12_convert(uint32_t u, float c0, float c1, float c2, uint32_t& lo, uint32_t& hi) {
  const float cr = fmaf(c12_u8f<0>(u), 0.00390625f, c0);
  const float cg = fmaf(c12_u8f<1>(u), 0.00390625f, c1);
  const float cb = fmaf(c12_u8f<2>(u), 0.00390625f, c2);
  const bool a = (u >> 24) != 0u;
  lo = a ? pv_pack_bf16x2(cr, cg) : 0u;
  hi = a ? pv_pack_bf16x2(cb, 0.f) : 0u;
}

struct ItemGeo {
  int b, strip, ra, rb;     // image, strip, conv2 rows [ra, rb)
  __device__ __forceinline__ void set(const C12Params& p, int item) {
    const int per_img = p.segs * p.strips;
    b = item / per_img;
    const int r = item - b * per_img;
    const int seg = r / p.strips;
    strip = r - seg * p.strips;
    ra = seg * p.seg_rows;
    rb = min(ra + p.seg_rows, p.OH2);
  }
};

__global__ void __launch_bounds__(kThreads, 1) c12_kernel(const __grid_constant__ C12Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a2 = smem + kOffA2;
  uint8_t* w2s = smem + kOffW2;
  uint8_t* w1s = smem + kOffW1;
  uint8_t* pxb = smem + kOffPx;
  uint8_t* raw = smem + kOffRaw;
  uint64_t* bar_rawf = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* bar_rawe = bar_rawf + kRawRing;
  uint64_t* bar_pxf = bar_rawe + kRawRing;
  uint64_t* bar_pxe = bar_pxf + kPxRing;
  uint64_t* bar_c1f = bar_pxe + kPxRing;          // [tile][slot]
  uint64_t* bar_c1e = bar_c1f + 2 * kNS1;
  uint64_t* bar_a2f = bar_c1e + 2 * kNS1;
  uint64_t* bar_a2e = bar_a2f + kA2Ring;
  uint64_t* bar_c2f = bar_a2e + kA2Ring;
  uint64_t* bar_c2e = bar_c2f + kNS2;
  float* s_scale1 = reinterpret_cast<float*>(smem + kOffFl);
  float* s_shift1 = s_scale1 + kC1;
  float* s_scale2 = s_shift1 + kC1;
  float* s_shift2 = s_scale2 + kC2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_shift2 + kC2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- setup: weights (generic copies), pad pixels, barriers, TMEM ----
  for (int i = threadIdx.x; i < kW2Bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(w2s)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2_img) + i);
  for (int i = threadIdx.x; i < kW1Bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(w1s)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1_img) + i);
  for (int i = threadIdx.x; i < kPxRing * kQuad * 8; i += kThreads) {       // trailing 8 pixels of every pixel row
    const int row = i >> 3, q = i & 7;
    *reinterpret_cast<uint2*>(pxb + row * kPxRow + (2 * kRawHalf + q) * 8) = make_uint2(0u, 0u);
  }
  for (int i = threadIdx.x; i < kA2Ring * kA2Entry / 16; i += kThreads)      // pair rows 126..129 are never written: keep them finite
    reinterpret_cast<uint4*>(a2)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x < kC1) {
    s_scale1[threadIdx.x] = p.scale1[threadIdx.x];
    s_shift1[threadIdx.x] = p.shift1[threadIdx.x];
  }
  if (threadIdx.x < kC2) {
    s_scale2[threadIdx.x] = p.scale2[threadIdx.x];
    s_shift2[threadIdx.x] = p.shift2[threadIdx.x];
  }
  if (warp == 0 && lane == 0) {
    pv_tma_prefetch_desc(&p.raw);
    for (int i = 0; i < kRawRing; ++i) { pv_mbar_init(&bar_rawf[i], 1); pv_mbar_init(&bar_rawe[i], kConvWarps); }
    for (int i = 0; i < kPxRing; ++i) { pv_mbar_init(&bar_pxf[i], kConvWarps); pv_mbar_init(&bar_pxe[i], 2); }
    for (int i = 0; i < 2 * kNS1; ++i) { pv_mbar_init(&bar_c1f[i], 1); pv_mbar_init(&bar_c1e[i], kEp1Warps / 2); }
    for (int i = 0; i < kA2Ring; ++i) { pv_mbar_init(&bar_a2f[i], kEp1Warps); pv_mbar_init(&bar_a2e[i], 1); }
    for (int i = 0; i < kNS2; ++i) { pv_mbar_init(&bar_c2f[i], 1); pv_mbar_init(&bar_c2e[i], kEp2Warps); }
    pv_fence_mbar_init();
  }
  pv_fence_proxy_async();                  // weights / zero pads were written through the generic proxy
  if (warp == 3) pv_tmem_alloc(s_tmem, 512);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // all accumulator rows start cleared (every MMA accumulates)
  if (warp >= 4 + kConvWarps) {
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4 + kConvWarps + kEp1Warps) {
      const int j = (warp - (4 + kConvWarps)) >> 2;
#pragma unroll
      for (int c = 0; c < kNS1; ++c) pv_tmem_st16_zero(lane_base + (uint32_t)(j * 128 + c * 16));
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) pv_tmem_st16_zero(lane_base + 256u + (uint32_t)(c * 16));
    }
    pv_tmem_st_wait();
  }
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer: one quad of plane rows per stage =====================
    if (pv_elect_one()) {
      int rs = 0;
      uint32_t rph = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        ItemGeo g;
        g.set(p, item);
        const int nq = (g.rb - g.ra) + 3;
        const int x0 = 4 * kW2 * g.strip;                       // first plane pixel of the strip
        const int y0 = g.b * p.Hp + 4 * g.ra;
        for (int q = 0; q < nq; ++q) {
          pv_mbar_wait(&bar_rawe[rs], rph ^ 1u, p.err, 1);
          pv_mbar_arrive_expect_tx(&bar_rawf[rs], (uint32_t)kRawStage);
          uint8_t* dst = raw + rs * kRawStage;
          pv_tma_load_2d(dst, &p.raw, &bar_rawf[rs], x0, y0 + 4 * q);
          pv_tma_load_2d(dst + kRawHalfBytes, &p.raw, &bar_rawf[rs], x0 + kRawHalf, y0 + 4 * q);
          if (++rs == kRawRing) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== conv1 MMA issuers: warp 1 -> M tile 0, warp 3 -> M tile 1 =====================
    if (pv_elect_one()) {
      const int j = warp == 1 ? 0 : 1;
      const uint32_t w1_16 = ((pv_smem_u32(w1s) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t dbase = tmem_base + (uint32_t)(j * 128);
      uint64_t* bar_e = bar_c1e + j * kNS1;
      uint64_t* bar_f = bar_c1f + j * kNS1;
      uint32_t emask = 0xFFu;                                   // parity to wait for on bar_e[slot]
      int ps = 0;
      uint32_t pph = 0;
      long long d_wpx = 0, d_issue = 0, d_quads = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        ItemGeo g;
        g.set(p, item);
        const int L = g.rb - g.ra;
        const int r_max = 2 * L + 2;                            // conv1 rows 0 .. 2L+2 of the item, row r in slot r & 7
        const int nq = L + 3;
        for (int q = 0; q < nq; ++q) {
          const long long t0 = p.dbg ? clock64() : 0;
          pv_mbar_wait(&bar_pxf[ps], pph, p.err, 2);
          pv_tc_fence_after();
          const long long t1 = p.dbg ? clock64() : 0;
          const uint32_t px0 = pv_smem_u32(pxb + ps * kPxStage) + (uint32_t)(j * kT1 * 16);
          const uint64_t ad0 = c12_desc_plain(px0, 16u, 128u);
          const uint64_t pd0 = c12_desc_plain(px0 + 32u, (uint32_t)kPxRow, 128u);
          if (q >= 1 && q <= L) {
            switch (q & 3) {
              case 0: c1_quad_static<0>(dbase, ad0, pd0, w1_16, bar_e, bar_f, emask, p.err); break;
              case 1: c1_quad_static<1>(dbase, ad0, pd0, w1_16, bar_e, bar_f, emask, p.err); break;
              case 2: c1_quad_static<2>(dbase, ad0, pd0, w1_16, bar_e, bar_f, emask, p.err); break;
              default: c1_quad_static<3>(dbase, ad0, pd0, w1_16, bar_e, bar_f, emask, p.err); break;
            }
          } else {
            // first quad / the quads past the last conv1 row: same sequence with the row range clamped
#pragma unroll
            for (int tt = 0; tt < kQuad; ++tt) {
              const int r_top = 2 * q + (tt >> 1);
              if ((tt & 1) == 0) {
                if (r_top <= r_max) c1_open(bar_e, emask, r_top & (kNS1 - 1), p.err);
                c1_issue(dbase, ad0 + (uint64_t)(tt * kPxRow16), w1_16 + (uint32_t)(kW1M0 >> 4), r_top - 2, 3, r_max);
              } else {
                c1_issue(dbase, ad0 + (uint64_t)(tt * kPxRow16), w1_16 + (uint32_t)(kW1M1 >> 4), r_top - 1, 2, r_max);
                c1_issue(dbase, pd0 + (uint64_t)((tt - 1) * kPxRow16), w1_16 + (uint32_t)(kW1P >> 4), r_top - 2, 3, r_max);
                const int rc = r_top - 2;
                if (rc >= 0 && rc <= r_max) pv_umma_commit(&bar_f[rc & (kNS1 - 1)]);
              }
            }
          }
          pv_umma_commit(&bar_pxe[ps]);
          if (++ps == kPxRing) { ps = 0; pph ^= 1u; }
          if (p.dbg) { d_wpx += t1 - t0; d_issue += clock64() - t1; ++d_quads; }
        }
      }
      if (p.dbg) {
        long long* d = p.dbg + (long long)blockIdx.x * 16;
        if (j == 0) { d[0] = d_wpx; d[2] = d_issue; d[3] = d_quads; } else { d[1] = d_issue; }
      }
    }
  } else if (warp == 2) {
    // ===================== conv2 MMA issuer (rsconv<16,32,5,5,2> over the shared-memory conv1 rows) =====================
    if (pv_elect_one()) {
      const uint32_t w2_16 = (pv_smem_u32(w2s) & 0x3FFFFu) >> 4;
      const uint32_t a2_lo = ((pv_smem_u32(a2) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t d2 = tmem_base + 256u;
      constexpr uint32_t BLK16 = (uint32_t)(kC2 * 32 >> 4);
      int e = 0;
      uint32_t eph = 0;
      int slot_base = 0;
      uint32_t par_base = 0;
      long long d_wa = 0, d_ws = 0, d_issue = 0, d_rows = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        ItemGeo g;
        g.set(p, item);
        const int L = g.rb - g.ra;
        const int t1 = 2 * (L - 1) + 4;
        int slot_top = slot_base;
        uint32_t par_top = par_base;
        for (int t = 0; t <= t1; ++t) {
          const int q = t & 1;
          const bool odd = q == 1;
          const int nqr = odd ? 2 : 3;
          const int r_top = t >> 1;
          const int r_first = r_top - (nqr - 1);
          const int r_hi = min(r_top, L - 1);
          const int r_lo = max(r_first, 0);
          const bool new_row = !odd && r_top <= L - 1;
          const long long c0 = p.dbg ? clock64() : 0;
          if (new_row) {
            pv_mbar_wait(&bar_c2e[slot_top], par_top ^ 1u, p.err, 4);
          }
          const long long c1 = p.dbg ? clock64() : 0;
          pv_mbar_wait(&bar_a2f[e], eph, p.err, 5);
          pv_tc_fence_after();
          const long long c2 = p.dbg ? clock64() : 0;
          const int n_all = r_hi - r_lo + 1;
          int slot_lo = slot_top - (r_top - r_lo);
          if (slot_lo < 0) slot_lo += kNS2;
          const int n1 = min(n_all, kNS2 - slot_lo);
          const int n2 = n_all - n1;
          const uint32_t dd1 = d2 + (uint32_t)(slot_lo * kC2);
          const uint32_t id1 = kIdesc0 | ((uint32_t)(n1 * kC2 >> 3) << 17);
          const uint32_t id2 = kIdesc0 | ((uint32_t)(n2 * kC2 >> 3) << 17);
          const uint32_t tile16 = (uint32_t)((odd ? kW2Tile1 : kW2Tile0) >> 4);
          const uint32_t bb1 = (w2_16 + (odd ? (uint32_t)(kW2Q1Base >> 4) : 0u) + (uint32_t)(r_lo - r_first) * BLK16) | (1u << 16);
          const uint32_t bb2 = bb1 + (uint32_t)n1 * BLK16;
          const uint32_t a0 = a2_lo + (uint32_t)e * (uint32_t)(kA2Entry >> 4);
#pragma unroll
          for (int kw = 0; kw < 5; ++kw) {
            const uint32_t a_lo = a0 + (uint32_t)(((kw >> 1) * kA2RowB + (kw & 1) * 32) >> 4);
            const uint64_t da = ((uint64_t)kA2Hi << 32) | a_lo;
            const uint32_t boff = (uint32_t)kw * tile16;
            c12_umma(dd1, da, ((uint64_t)kBHi << 32) | (bb1 + boff), id1, 1u);
            if (n2 > 0) c12_umma(d2, da, ((uint64_t)kBHi << 32) | (bb2 + boff), id2, 1u);
          }
          pv_umma_commit(&bar_a2e[e]);
          if (!odd) {
            const int rc = r_top - 2;                      // the row whose kh = 4 this was
            if (rc >= 0 && rc < L) {
              int sc = slot_top - 2;
              if (sc < 0) sc += kNS2;
              pv_umma_commit(&bar_c2f[sc]);
            }
          }
          if (++e == kA2Ring) { e = 0; eph ^= 1u; }
          if (odd && ++slot_top == kNS2) { slot_top = 0; par_top ^= 1u; }
          if (p.dbg) { d_ws += c1 - c0; d_wa += c2 - c1; d_issue += clock64() - c2; ++d_rows; }
        }
        const int adv = (slot_base + L) / kNS2;
        slot_base = (slot_base + L) % kNS2;
        par_base ^= (uint32_t)(adv & 1);
      }
      if (p.dbg) {
        long long* d = p.dbg + (long long)blockIdx.x * 16;
        d[4] = d_wa; d[5] = d_ws; d[6] = d_issue; d[7] = d_rows;
      }
    }
  } else if (warp < 4 + kConvWarps) {
    // ===================== converters: raw RGBA u8 -> bf16 RGB0, every plane pixel once =====================
    const int tc = threadIdx.x - 128;                 // 0..255: pixel pair (2 tc, 2 tc + 1) of every row of the quad
    const int half = tc >> 7;
    const uint32_t src_off = (uint32_t)(half * kRawHalfBytes + (tc & 127) * 8);
    const uint32_t dst_off = (uint32_t)(tc * 16);
    int rs = 0, ps = 0;
    uint32_t rph = 0, pph = 0;
    long long k_wr = 0, k_we = 0, k_work = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      ItemGeo g;
      g.set(p, item);
      const int nq = (g.rb - g.ra) + 3;
      for (int q = 0; q < nq; ++q) {
        const long long k0 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_rawf[rs], rph, p.err, 6);
        const long long k1 = p.dbg ? clock64() : 0;
        uint2 v[kQuad];
        const uint8_t* rb = raw + rs * kRawStage + src_off;
#pragma unroll
        for (int r = 0; r < kQuad; ++r) v[r] = *reinterpret_cast<const uint2*>(rb + r * (kRawHalf * 4));
        pv_mbar_wait(&bar_pxe[ps], pph ^ 1u, p.err, 7);
        const long long k2 = p.dbg ? clock64() : 0;
        uint8_t* db = pxb + ps * kPxStage + dst_off;
#pragma unroll
        for (int r = 0; r < kQuad; ++r) {
          uint4 o;
          c12_convert(v[r].x, p.c0, p.c1, p.c2, o.x, o.y);
          c12_convert(v[r].y, p.c0, p.c1, p.c2, o.z, o.w);
          *reinterpret_cast<uint4*>(db + r * kPxRow) = o;
        }
        pv_fence_proxy_async();                          // generic-proxy smem writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) {
          pv_mbar_arrive(&bar_pxf[ps]);
          pv_mbar_arrive(&bar_rawe[rs]);
        }
        if (++rs == kRawRing) { rs = 0; rph ^= 1u; }
        if (++ps == kPxRing) { ps = 0; pph ^= 1u; }
        if (p.dbg) { k_wr += k1 - k0; k_we += k2 - k1; k_work += clock64() - k2; }
      }
    }
    if (p.dbg && threadIdx.x == 128) {
      long long* d = p.dbg + (long long)blockIdx.x * 16;
      d[8] = k_wr; d[9] = k_we; d[10] = k_work;
    }
  } else if (warp < 4 + kConvWarps + kEp1Warps) {
    // ===================== conv1 epilogue: TMEM -> affine -> ReLU -> bf16 -> conv2's A operand in shared memory =====================
    const int j = (warp - (4 + kConvWarps)) >> 2;       // M tile
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const bool valid = m < kT1;
    const int x1 = j * kT1 + m;                          // conv1 column inside the strip
    const int pr = x1 >> 1;
    const uint32_t sw = (uint32_t)((pr >> 1) & 3);
    const uint32_t off0 = (uint32_t)(pr * kA2RowB) + ((((uint32_t)(x1 & 1) * 2u + 0u) ^ sw) << 4);
    const uint32_t off1 = (uint32_t)(pr * kA2RowB) + ((((uint32_t)(x1 & 1) * 2u + 1u) ^ sw) << 4);
    uint64_t* bar_e = bar_c1e + j * kNS1;
    uint64_t* bar_f = bar_c1f + j * kNS1;
    uint32_t fmask = 0u;                                  // parity to wait for on bar_f[slot]
    int e = 0;
    uint32_t eph = 0;
    long long e_wt = 0, e_wa = 0, e_work = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      ItemGeo g;
      g.set(p, item);
      const int nr = 2 * (g.rb - g.ra) + 3;
      for (int r = 0; r < nr; ++r) {
        const uint32_t slot = (uint32_t)r & (kNS1 - 1);    // every item starts at slot 0
        const long long q0 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_f[slot], (fmask >> slot) & 1u, p.err, 8);
        fmask ^= 1u << slot;
        pv_tc_fence_after();
        const long long q1 = p.dbg ? clock64() : 0;
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(j * 128) + slot * kC1;
        pv_tmem_ld16(taddr, v);
        pv_tmem_ld_wait();
        pv_tmem_st16_zero(taddr);                          // hand the row back cleared
        pv_tmem_st_wait();
        pv_tc_fence_before();
        __syncwarp();
        if (lane == 0) pv_mbar_arrive(&bar_e[slot]);       // accumulator row is in registers: release the slot
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float f0 = fmaf(__uint_as_float(v[2 * k]), s_scale1[2 * k], s_shift1[2 * k]);
          float f1 = fmaf(__uint_as_float(v[2 * k + 1]), s_scale1[2 * k + 1], s_shift1[2 * k + 1]);
          o[k] = pv_pack_bf16x2(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
        }
        pv_mbar_wait(&bar_a2e[e], eph ^ 1u, p.err, 9);
        const long long q2 = p.dbg ? clock64() : 0;
        if (valid) {
          uint8_t* eb = a2 + e * kA2Entry;
          *reinterpret_cast<uint4*>(eb + off0) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(eb + off1) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        pv_fence_proxy_async();
        __syncwarp();
        if (lane == 0) pv_mbar_arrive(&bar_a2f[e]);
        if (++e == kA2Ring) { e = 0; eph ^= 1u; }
        if (p.dbg) { e_wt += q1 - q0; e_wa += q2 - q1; e_work += clock64() - q2; }
      }
    }
    if (p.dbg && threadIdx.x == 32 * (4 + kConvWarps)) {
      long long* d = p.dbg + (long long)blockIdx.x * 16;
      d[11] = e_wt; d[12] = e_wa; d[13] = e_work;
    }
  } else {
    // ===================== conv2 epilogue: TMEM -> affine -> ReLU -> bf16 -> NHWC global =====================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    uint32_t cnt = 0;
    long long f_wt = 0, f_work = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      ItemGeo g;
      g.set(p, item);
      const int ox = g.strip * kW2 + m;
      const bool xvalid = m < kW2 && ox < p.OW2;
      for (int r = g.ra; r < g.rb; ++r, ++cnt) {
        const uint32_t slot = cnt % kNS2;
        const long long q0 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_c2f[slot], (cnt / kNS2) & 1u, p.err, 10);
        pv_tc_fence_after();
        const long long q1 = p.dbg ? clock64() : 0;
        uint32_t v[2][16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + 256u + slot * kC2;
        pv_tmem_ld16(taddr, v[0]);
        pv_tmem_ld16(taddr + 16, v[1]);
        pv_tmem_ld_wait();
        pv_tmem_st16_zero(taddr);                          // hand the row back cleared
        pv_tmem_st16_zero(taddr + 16);
        pv_tmem_st_wait();
        pv_tc_fence_before();
        __syncwarp();
        if (lane == 0) pv_mbar_arrive(&bar_c2e[slot]);
        if (xvalid) {
          __nv_bfloat16* dp = p.out + (((long long)g.b * p.OH2 + r) * p.out_pitch + ox) * kC2;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float f0 = fmaf(__uint_as_float(v[h][2 * k]), s_scale2[h * 16 + 2 * k], s_shift2[h * 16 + 2 * k]);
              float f1 = fmaf(__uint_as_float(v[h][2 * k + 1]), s_scale2[h * 16 + 2 * k + 1], s_shift2[h * 16 + 2 * k + 1]);
              o[k] = pv_pack_bf16x2(fmaxf(f0, 0.f), fmaxf(f1, 0.f));
            }
            pv_stg256(dp + h * 16, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
          }
        }
        if (p.dbg) { f_wt += q1 - q0; f_work += clock64() - q1; }
      }
    }
    if (p.dbg && threadIdx.x == 32 * (4 + kConvWarps + kEp1Warps)) {
      long long* d = p.dbg + (long long)blockIdx.x * 16;
      d[14] = f_wt; d[15] = f_work;
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct C12Plan {
  C12Params p;
  int num_sms = 0;
  int Bmax = 0;
  int* d_err = nullptr;
  long long* d_dbg = nullptr;
};

}  // namespace

extern "C" int pv_c12_create(const PvC12Desc* d, void** out_handle) {
  PV_REQUIRE(d && out_handle, "pv_c12_create: null argument");
  PV_REQUIRE(d->plane && d->w1_img && d->w2_img && d->scale1 && d->shift1 && d->scale2 && d->shift2 && d->out && d->mean_host,
             "pv_c12_create: null operand");
  PV_REQUIRE(d->B > 0 && d->Hp >= 13 && d->Wp >= 13 && d->Wp % 4 == 0, "pv_c12_create: bad plane %d x %d x %d (Wp must be a multiple of 4)",
             d->B, d->Hp, d->Wp);
  PV_REQUIRE(d->w1_bytes == kW1Bytes && d->w2_bytes == kW2Bytes, "pv_c12_create: weight images are %lld / %lld bytes, expected %d / %d",
             (long long)d->w1_bytes, (long long)d->w2_bytes, kW1Bytes, kW2Bytes);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(d->plane) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w1_img) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->w2_img) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->out) & 31) == 0,
             "pv_c12_create: operands must be 16-byte (output: 32-byte) aligned");
  const int OH1 = (d->Hp - 5) / 2 + 1, OW1 = (d->Wp - 5) / 2 + 1;
  const int OH2 = (OH1 - 5) / 2 + 1, OW2 = (OW1 - 5) / 2 + 1;
  PV_REQUIRE(OH2 > 0 && OW2 > 0, "pv_c12_create: empty output");
  PV_REQUIRE(d->out_pitch >= OW2, "pv_c12_create: output pitch %d < %d", d->out_pitch, OW2);
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PV_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
    PV_REQUIRE(qres == cudaDriverEntryPointSuccess && fp, "pv_c12_create: cuTensorMapEncodeTiled unavailable");
    enc = reinterpret_cast<EncodeTiledFn>(fp);
  }
  C12Plan* plan = new C12Plan();
  memset(&plan->p, 0, sizeof(C12Params));
  C12Params& p = plan->p;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    pv_set_error("pv_c12_create: no CUDA device: %s", cudaGetErrorString(e));
    delete plan;
    return PV_ERR_CUDA;
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)d->Wp, (cuuint64_t)d->B * (cuuint64_t)d->Hp};
    cuuint64_t gstride[1] = {(cuuint64_t)d->Wp * 4};
    cuuint32_t box[2] = {kRawHalf, kQuad};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.raw, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(d->plane), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pv_set_error("pv_c12_create: cuTensorMapEncodeTiled failed (CUresult %d, Wp=%d rows=%lld)", (int)r, d->Wp, (long long)d->B * d->Hp);
      delete plan;
      return PV_ERR_CUDA;
    }
  }
  if (cudaMalloc(&plan->d_err, sizeof(int)) != cudaSuccess) {
    pv_set_error("pv_c12_create: cudaMalloc failed");
    delete plan;
    return PV_ERR_CUDA;
  }
  cudaMemset(plan->d_err, 0, sizeof(int));
  if (getenv("PV_C12_DEBUG")) {
    cudaMalloc(&plan->d_dbg, sizeof(long long) * 16 * plan->num_sms);
    cudaMemset(plan->d_dbg, 0, sizeof(long long) * 16 * plan->num_sms);
  }
  p.dbg = plan->d_dbg;
  p.w1_img = static_cast<const uint8_t*>(d->w1_img);
  p.w2_img = static_cast<const uint8_t*>(d->w2_img);
  p.scale1 = d->scale1; p.shift1 = d->shift1;
  p.scale2 = d->scale2; p.shift2 = d->shift2;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.B = d->B; p.Hp = d->Hp; p.Wp = d->Wp;
  p.OH2 = OH2; p.OW2 = OW2; p.out_pitch = d->out_pitch;
  p.strips = (OW2 + kW2 - 1) / kW2;
  {
    // rows per work item: minimise ceil(items / CTAs) x (quads per item) among heights whose halo (3 quads) stays small
    long long best_cost = -1;
    int best_rows = OH2;
    for (int segs = 1; segs <= OH2; ++segs) {
      const int rows = (OH2 + segs - 1) / segs;
      if (rows < 16 && segs > 1) break;
      const int nseg = (OH2 + rows - 1) / rows;
      const long long items = (long long)d->B * nseg * p.strips;
      const long long waves = (items + plan->num_sms - 1) / plan->num_sms;
      const long long cost = waves * (rows + 3);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
    }
    p.seg_rows = best_rows;
    p.segs = (OH2 + best_rows - 1) / best_rows;
  }
  p.c0 = -d->mean_host[0] * 0.00390625f;
  p.c1 = -d->mean_host[1] * 0.00390625f;
  p.c2 = -d->mean_host[2] * 0.00390625f;
  p.err = plan->d_err;
  plan->Bmax = d->B;
  *out_handle = plan;
  return PV_OK;
}

extern "C" int pv_c12_run(void* handle, int B, void* stream) {
  PV_REQUIRE(handle, "pv_c12_run: null handle");
  C12Plan* plan = static_cast<C12Plan*>(handle);
  PV_REQUIRE(B > 0 && B <= plan->Bmax, "pv_c12_run: B=%d outside [1,%d]", B, plan->Bmax);
  C12Params p = plan->p;
  p.B = B;
  const long long ni = (long long)B * p.segs * p.strips;
  PV_REQUIRE(ni < (1ll << 31), "pv_c12_run: too many work items");
  p.num_items = (int)ni;
  const int grid = p.num_items < plan->num_sms ? p.num_items : plan->num_sms;
  static unsigned long long attr = 0;
  if (pv_attr_needed(&attr))
    PV_CUDA_CHECK(cudaFuncSetAttribute(c12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  c12_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_c12_info(void* handle, int* smem_bytes, int* strips, int* segs, int* seg_rows, int* oh2, int* ow2) {
  PV_REQUIRE(handle, "pv_c12_info: null handle");
  C12Plan* plan = static_cast<C12Plan*>(handle);
  if (smem_bytes) *smem_bytes = kSmemBytes;
  if (strips) *strips = plan->p.strips;
  if (segs) *segs = plan->p.segs;
  if (seg_rows) *seg_rows = plan->p.seg_rows;
  if (oh2) *oh2 = plan->p.OH2;
  if (ow2) *ow2 = plan->p.OW2;
  return PV_OK;
}

/* role timing of the last launch (PV_C12_DEBUG=1 at create time): out16 (HOST) = cycles summed over CTAs */
extern "C" int pv_c12_debug(void* handle, long long* out16) {
  PV_REQUIRE(handle && out16, "pv_c12_debug: null argument");
  C12Plan* plan = static_cast<C12Plan*>(handle);
  PV_REQUIRE(plan->d_dbg, "pv_c12_debug: plan was created without PV_C12_DEBUG");
  std::vector<long long> h(16 * plan->num_sms);
  PV_CUDA_CHECK(cudaMemcpy(h.data(), plan->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 16; ++k) out16[k] = 0;
  for (int i = 0; i < plan->num_sms; ++i)
    for (int k = 0; k < 16; ++k) out16[k] += h[16 * i + k];
  return PV_OK;
}

extern "C" int pv_c12_check(void* handle, void* stream) {
  PV_REQUIRE(handle, "pv_c12_check: null handle");
  C12Plan* plan = static_cast<C12Plan*>(handle);
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&flag, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    pv_set_error("pv_c12_check: %s", cudaGetErrorString(e));
    return PV_ERR_CUDA;
  }
  if (flag != 0) {
    pv_set_error("c12: device-side pipeline timeout (role code %d)", flag);
    cudaMemset(plan->d_err, 0, sizeof(int));
    return PV_ERR_DEVICE_TIMEOUT;
  }
  return PV_OK;
}

extern "C" int pv_c12_destroy(void* handle) {
  if (!handle) return PV_OK;
  C12Plan* plan = static_cast<C12Plan*>(handle);
  cudaFree(plan->d_err);
  if (plan->d_dbg) cudaFree(plan->d_dbg);
  delete plan;
  return PV_OK;
}
