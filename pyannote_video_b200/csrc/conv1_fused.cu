// conv1_fused.cu — first detector convolution (dlib con<16,5,5,2,2> on the tiled RGB pyramid, the
// input stage of face_detector_(rgb, 1), pyannote/video/face/face.py:66) reading the RGBA u8 plane
// directly.
//
// The generic path first writes a "gathered" bf16 matrix (16 B per plane pixel) and reads it back
// through TMA; with ~32 M plane pixels per 1080p frame that round trip is the largest HBM stream of
// the whole detector.  Here the plane is read once, as raw pixels, by TMA (asynchronous, deep ring);
// eight converter warps normalise every raw pixel ONCE — (v - mean)/256 as bf16 RGB0, 8 bytes — into
// a pixel-row buffer in shared memory, and tcgen05.mma reads its A operand straight out of that
// buffer: in the un-swizzled K-major layout the 8 rows of a core matrix are 16 bytes apart, which is
// exactly the 2-pixel step between neighbouring outputs of a stride-2 convolution, so the A rows
// (the 5-pixel windows) simply OVERLAP in shared memory and no im2col gather is ever materialised.
//   A row m of input row i = pixels 2m .. 2m+5 of buffer row i   (3 chunks of 2 pixels = 8 bf16)
// One input row serves every output row of the tile that reads it (kh = i - 2r), so its A operand is read
// ONCE and multiplied against the filter rows of all those output rows side by side (N = 16, 32 or 48):
//   11 "main" MMAs : chunks 0,1 of input row i (kw 0..3, LBO = 16 B)  x  [W_kh(r_lo) | .. | W_kh(r_hi)]
//    6 "pair" MMAs : chunk 2 (kw 4) of input rows (i, i+1) (LBO = one buffer row)
// = 17 MMAs per 4 x 126 outputs instead of 32 single-row ones (A traffic from shared memory -47 %).
// A tile is 126 consecutive outputs of FOUR consecutive output rows (2*125 + 5 = 255 input pixels fit one
// 256-pixel TMA box; rows 126/127 of each MMA are padding; 2*4 + 3 = 11 input rows).  Four rows per tile
// mean every input pixel is fetched and converted 11/8 = 1.4 times instead of 2.5 times, and the per-tile
// hand-offs (two mbarrier waits, one fence.proxy.async per converter warp) are paid once per 504 outputs.
// TMA's out-of-bounds zero fill is the plane's zero padding; tile coordinates advance without divisions.
#include <cuda.h>
#include <atomic>
#include <cstdlib>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kTileM = 128;
constexpr int kTileOut = 126;                     // valid outputs per tile
constexpr int kKH = 5;
constexpr int kN = 16;
constexpr int kRows = 4;                          // output rows per tile
constexpr int kInRows = 2 * kRows + 3;            // 11 input rows
constexpr int kRawW = 256;                        // pixels per raw row (TMA box)
constexpr int kRawBytes = kInRows * kRawW * 4;    // 11 KB
constexpr int kRawRing = 4;
constexpr int kPxRowBytes = (kRawW + 8) * 8;      // 264 pixels x 8 B; the 8 trailing pixels stay zero
constexpr int kPxSlotBytes = 23296;               // 11 rows (23232 B) rounded up to 128
constexpr int kPxRing = 2;
constexpr int kAcc = 4;                           // tiles in flight in TMEM (kRows accumulators each): 256 columns per CTA
constexpr int kNumMma = 17;                       // per tile: 11 main + 6 pair (see the header comment)

// ---- the MMA schedule of a tile (compile-time arithmetic, also run at setup to lay the weights out) ----
// order: main(4) and main(10) first — between them they touch every accumulator and open each with accumulate = 0
__host__ __device__ constexpr bool ent_pair(int e) { return e >= 11; }
__host__ __device__ constexpr int ent_row(int e) {
  return e == 0 ? 4 : (e == 1 ? 10 : (e < 11 ? (e - 2 < 4 ? e - 2 : e - 1) : 2 * (e - 11)));
}
__host__ __device__ constexpr int ent_rlo(int e) { return ent_row(e) - 3 > 0 ? (ent_row(e) - 3) >> 1 : 0; }
__host__ __device__ constexpr int ent_rhi(int e) {
  return ((ent_row(e) + (ent_pair(e) ? 1 : 0)) >> 1) < kRows - 1 ? ((ent_row(e) + (ent_pair(e) ? 1 : 0)) >> 1) : kRows - 1;
}
__host__ __device__ constexpr int ent_n(int e) { return (ent_rhi(e) - ent_rlo(e) + 1) * kN; }       // MMA N
__host__ __device__ constexpr int ent_boff(int e) { return e == 0 ? 0 : ent_boff(e - 1) + ent_n(e - 1) * 32; }
constexpr int kWBytes = ent_boff(kNumMma);        // all B tiles (N x 16 bf16 each)
static_assert(ent_rlo(0) == 0 && ent_rhi(0) == 2 && ent_rlo(1) == 3 && ent_rhi(1) == 3, "opening MMAs must cover all rows");
constexpr int kConvWarps = 8;
// warps: 0 = TMA, 1 = MMA, 2..9 = converters, 10..13 = epilogue
constexpr int kThreads = 32 * (2 + kConvWarps + 4);
static_assert(kPxSlotBytes >= kInRows * kPxRowBytes, "slot too small");
static_assert(kRows % 2 == 0, "the epilogue drains two rows per pass");

struct C1Params {
  CUtensorMap raw;       // uint32 [B*Hp rows, Wp cols], box 256 x 11 (default)
  const uint32_t* plane; // the same plane for the direct-load variant (PV_C1_TMA=0)
  int use_tma;
  int B, Hp, Wp;
  int oh, ow;            // valid output extent
  int oh_tiles;          // ceil(oh / kRows)
  int tiles_per_row;
  const __nv_bfloat16* w;  // [16][3][5][5]
  const float* scale;
  const float* shift;
  int relu;
  __nv_bfloat16* out;
  PvRowMap dst;
  float c0, c1, c2;      // -mean/256
  int num_tiles;
  int* err;
  long long* dbg;   // optional [grid][8] role cycle counters (PV_C1_DEBUG), nullptr = off
  int ablate;       // PV_C1_ABLATE (probe only, results become wrong): 1 no output stores, 2 no epilogue math/stores,
                    // 4 converters store without converting, 8 only the two opening MMAs per tile, 16 converters do not store,
                    // 32 no fence.proxy.async, 64 epilogue does not read TMEM, 128 no nanosleep back-off in the producer / epilogue waits
};

__device__ __forceinline__ long long row_of(const PvRowMap& m, uint32_t n, uint32_t y, uint32_t x) {
  const uint32_t Y = y + m.py, X = x + m.px;
  if (m.kind == 0) return (long long)n * m.img + (long long)Y * m.w + X;
  const uint32_t plane = ((Y & 1u) << 1) | (X & 1u);
  return (long long)plane * m.plane_rows + (long long)n * m.img + (long long)(Y >> 1) * m.w + (X >> 1);
}

// tile -> (image, tile row (kRows output rows), first output column), advanced by gridDim.x tiles without divisions
struct TileWalk {
  uint32_t n, oy, ct;
  uint32_t step_n, step_oy, step_ct;
  __device__ __forceinline__ void init(const C1Params& p, uint32_t tile, uint32_t step) {
    const uint32_t tpr = (uint32_t)p.tiles_per_row, oh = (uint32_t)p.oh_tiles;
    uint32_t r = tile / tpr;
    ct = tile - r * tpr;
    n = r / oh;
    oy = r - n * oh;
    r = step / tpr;
    step_ct = step - r * tpr;
    step_n = r / oh;
    step_oy = r - step_n * oh;
  }
  __device__ __forceinline__ void next(const C1Params& p) {
    ct += step_ct;
    oy += step_oy;
    n += step_n;
    if (ct >= (uint32_t)p.tiles_per_row) { ct -= (uint32_t)p.tiles_per_row; ++oy; }
    if (oy >= (uint32_t)p.oh_tiles) { oy -= (uint32_t)p.oh_tiles; ++n; }
  }
};

__device__ __forceinline__ void c1_prefetch_2d(const void* tmap, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1)
               : "memory");
}

// un-swizzled K-major operand: start address, LBO = byte step between the two 8-element K chunks,
// SBO = byte step between 8-row groups (cute::UMMA::SmemDescriptor, LayoutType::INTERLEAVE; version 1)
__device__ __forceinline__ uint64_t desc_kmajor_plain(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// the 17 MMAs of a tile, schedule entries forced to compile-time constants (template recursion)
template <int E>
__device__ __forceinline__ void issue_tile(uint32_t a0, uint32_t w0, uint32_t tmem_tile, uint32_t idesc0, uint32_t lead,
                                           uint32_t lead_rest) {
  if constexpr (E < kNumMma) {
    constexpr int i0 = ent_row(E), N = ent_n(E), rlo = ent_rlo(E), boff = ent_boff(E);
    constexpr bool pair = ent_pair(E);
    // (the last pair has no second input row: its second K chunk re-reads this row — finite data, zero weights)
    constexpr uint32_t a_off = pair ? (uint32_t)(i0 * kPxRowBytes + 32) : (uint32_t)(i0 * kPxRowBytes);
    constexpr uint32_t a_lbo = pair ? (i0 + 1 < kInRows ? (uint32_t)kPxRowBytes : 16u) : 16u;
    const uint64_t ad = desc_kmajor_plain(a0 + a_off, a_lbo, 128u);
    const uint64_t bd = desc_kmajor_plain(w0 + (uint32_t)boff, (uint32_t)(N * 16), 128u);
    pv_umma_bf16_pred(tmem_tile + (uint32_t)(rlo * kN), ad, bd, idesc0 | ((uint32_t)(N >> 3) << 17), E >= 2 ? 1u : 0u,
                      E >= 2 ? lead_rest : lead);
    issue_tile<E + 1>(a0, w0, tmem_tile, idesc0, lead, lead_rest);
  }
}

__global__ void __launch_bounds__(kThreads, 2) conv1_fused_kernel(const __grid_constant__ C1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* pxb = smem;                                // kPxRing slots of 5 bf16 pixel rows
  uint8_t* raw = pxb + kPxRing * kPxSlotBytes;        // kRawRing raw pixel blocks
  uint8_t* wsm = raw + kRawRing * kRawBytes;          // 8 B operands
  uint64_t* bar_rfull = reinterpret_cast<uint64_t*>(wsm + kWBytes);
  uint64_t* bar_rempty = bar_rfull + kRawRing;
  uint64_t* bar_full = bar_rempty + kRawRing;
  uint64_t* bar_empty = bar_full + kPxRing;
  uint64_t* bar_tfull = bar_empty + kPxRing;
  uint64_t* bar_tempty = bar_tfull + kAcc;
  float* s_scale = reinterpret_cast<float*>(bar_tempty + kAcc);
  float* s_shift = s_scale + kN;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_shift + kN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- setup ----
  // B tile of schedule entry e: rows nn = (r - r_lo)*16 + n, element (nn, k) at
  //   boff(e) + (k>>3)*(N*16) + (nn>>3)*128 + (nn&7)*16 + (k&7)*2          (un-swizzled K-major core matrices)
  // main: k = (kw 0..3, ch) of input row i, kh = i - 2r;  pair: k<8 -> (kw 4|pad, ch) of row i, k>=8 -> of row i+1
  int boff = 0;
  for (int e = 0; e < kNumMma; boff += ent_n(e) * 32, ++e) {
    const int i0 = ent_row(e), rlo = ent_rlo(e), N = ent_n(e);
    const bool pair = ent_pair(e);
    for (int idx = threadIdx.x; idx < N * 16; idx += kThreads) {
      const int nn = idx >> 4, k = idx & 15;
      const int r = rlo + (nn >> 4), n = nn & 15;
      const int ch = k & 3;
      int kh, kw;
      bool ok = ch < 3;
      if (!pair) {
        kw = k >> 2;
        kh = i0 - 2 * r;
      } else {
        const int ii = i0 + (k >> 3);
        kw = 4 + ((k >> 2) & 1);
        kh = ii - 2 * r;
        ok = ok && kw == 4 && ii < kInRows;
      }
      ok = ok && kh >= 0 && kh < kKH;
      const __nv_bfloat16 v = ok ? p.w[((n * 3 + ch) * 5 + kh) * 5 + kw] : __float2bfloat16(0.f);
      *reinterpret_cast<__nv_bfloat16*>(wsm + boff + (k >> 3) * (N * 16) + (nn >> 3) * 128 + (nn & 7) * 16 + (k & 7) * 2) = v;
    }
  }
  // trailing 8 pixels of every buffer row: read by the padding rows of the MMA, must be finite
  for (int i = threadIdx.x; i < kPxRing * kInRows * 8; i += kThreads) {
    const int s = i / (kInRows * 8), rem = i - s * (kInRows * 8);
    const int kh = rem >> 3, q = rem & 7;
    *reinterpret_cast<uint2*>(pxb + s * kPxSlotBytes + kh * kPxRowBytes + (kRawW + q) * 8) = make_uint2(0u, 0u);
  }
  if (threadIdx.x < kN) {
    s_scale[threadIdx.x] = p.scale[threadIdx.x];
    s_shift[threadIdx.x] = p.shift[threadIdx.x];
  }
  if (warp == 0 && lane == 0) {
    pv_tma_prefetch_desc(&p.raw);
    for (int i = 0; i < kRawRing; ++i) {
      pv_mbar_init(&bar_rfull[i], 1);
      pv_mbar_init(&bar_rempty[i], kConvWarps);
    }
    for (int i = 0; i < kPxRing; ++i) {
      pv_mbar_init(&bar_full[i], kConvWarps);     // one arrival per converter warp
      pv_mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      pv_mbar_init(&bar_tfull[i], 1);
      pv_mbar_init(&bar_tempty[i], 4);
    }
    pv_fence_mbar_init();
  }
  pv_fence_proxy_async();                  // weights / zero pads written through the generic proxy, read by the tensor core
  if (warp == 1) pv_tmem_alloc(s_tmem, kAcc * kRows * kN);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== TMA: raw pixel blocks (PV_C1_TMA=1 only) =====================
    if (p.use_tma && pv_elect_one()) {
      TileWalk t, tp;
      t.init(p, blockIdx.x, gridDim.x);
      // the plane streams from HBM once: an L2 prefetch runs kPrefetch tiles ahead of the shared-memory ring so that
      // the ring's loads find their lines in L2 (ablation probe: the bare pipeline was paced by load latency)
      constexpr int kPrefetch = 8;
      tp.init(p, blockIdx.x, gridDim.x);
      int ahead = blockIdx.x;
      for (int k = 0; k < kPrefetch && ahead < p.num_tiles; ++k, ahead += gridDim.x) {
        if (k >= kRawRing) c1_prefetch_2d(&p.raw, (int32_t)(2 * kTileOut * tp.ct), (int32_t)(tp.n * (uint32_t)p.Hp + 2 * kRows * tp.oy));
        tp.next(p);
      }
      int rs = 0;
      uint32_t rphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (ahead < p.num_tiles) {
          c1_prefetch_2d(&p.raw, (int32_t)(2 * kTileOut * tp.ct), (int32_t)(tp.n * (uint32_t)p.Hp + 2 * kRows * tp.oy));
          tp.next(p);
          ahead += gridDim.x;
        }
        if (p.ablate & 128) pv_mbar_wait(&bar_rempty[rs], rphase ^ 1u, p.err, 1);
        else pv_mbar_wait_backoff(&bar_rempty[rs], rphase ^ 1u, p.err, 1, 64);
        pv_mbar_arrive_expect_tx(&bar_rfull[rs], kRawBytes);
        pv_tma_load_2d(raw + rs * kRawBytes, &p.raw, &bar_rfull[rs], (int32_t)(2 * kTileOut * t.ct),
                       (int32_t)(t.n * (uint32_t)p.Hp + 2 * kRows * t.oy));
        if (++rs == kRawRing) { rs = 0; rphase ^= 1u; }
        t.next(p);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t lead = pv_elect_one() ? 1u : 0u;
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileM >> 4) << 24);   // + (N >> 3) << 17
    const uint32_t w0 = pv_smem_u32(wsm);
    int slot = 0, buf = 0;
    uint32_t phase = 0, aphase = 0;
    long long m_wt = 0, m_wf = 0, m_is = 0, m_n = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long c0 = p.dbg ? clock64() : 0;
      pv_mbar_wait(&bar_tempty[buf], aphase ^ 1u, p.err, 2);
      const long long c1 = p.dbg ? clock64() : 0;
      pv_mbar_wait(&bar_full[slot], phase, p.err, 3);
      pv_tc_fence_after();
      const long long c2 = p.dbg ? clock64() : 0;
      const uint32_t a0 = pv_smem_u32(pxb + slot * kPxSlotBytes);
      issue_tile<0>(a0, w0, tmem_base + (uint32_t)(buf * kRows * kN), idesc0, lead, (p.ablate & 8) ? 0u : lead);
      pv_umma_commit_pred(&bar_empty[slot], lead);
      pv_umma_commit_pred(&bar_tfull[buf], lead);
      if (++slot == kPxRing) { slot = 0; phase ^= 1u; }
      if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
      if (p.dbg) { m_wt += c1 - c0; m_wf += c2 - c1; m_is += clock64() - c2; ++m_n; }
    }
    if (p.dbg && lead) {
      long long* d = p.dbg + (long long)blockIdx.x * 8;
      d[0] = m_wt; d[1] = m_wf; d[2] = m_is; d[3] = m_n;
    }
  } else if (warp < 2 + kConvWarps) {
    // ===================== converters: every raw pixel -> bf16 RGB0, once =====================
    const int ct = threadIdx.x - 64;               // pixel column 0..255 of the block
    int rs = 0, slot = 0;
    uint32_t rphase = 0, phase = 0;
    long long k_wr = 0, k_we = 0, k_work = 0;
    // Alternative (PV_C1_TMA=0): the converters fetch their raw pixels themselves — one coalesced 4-byte load per input
    // row, issued one tile AHEAD (registers).  Measured slower than the TMA ring (1030 vs 872 us per 8 frames on the same
    // box); kept as a probe: it shows that the TMA unit is not what paces the bare pipeline (profiles/README.md).
    TileWalk tw;
    tw.init(p, blockIdx.x, gridDim.x);
    const long long total_rows = (long long)p.B * p.Hp;
    uint32_t vn[kInRows];
    auto fetch = [&](const TileWalk& t) {
      const long long row0 = (long long)t.n * p.Hp + (long long)(2 * kRows) * t.oy;
      const uint32_t x = 2u * kTileOut * t.ct + (uint32_t)ct;
      const uint32_t* src = p.plane + row0 * p.Wp + x;
#pragma unroll
      for (int kh = 0; kh < kInRows; ++kh)
        vn[kh] = (x < (uint32_t)p.Wp && row0 + kh < total_rows) ? __ldg(src + (long long)kh * p.Wp) : 0u;
    };
    if (!p.use_tma && blockIdx.x < p.num_tiles) fetch(tw);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long k0 = p.dbg ? clock64() : 0;
      uint32_t v[kInRows];
      long long k1 = k0;
      if (p.use_tma) {
        pv_mbar_wait(&bar_rfull[rs], rphase, p.err, 5);
        k1 = p.dbg ? clock64() : 0;
        const uint32_t* rb = reinterpret_cast<const uint32_t*>(raw + rs * kRawBytes) + ct;
#pragma unroll
        for (int kh = 0; kh < kInRows; ++kh) v[kh] = rb[kh * kRawW];
      } else {
#pragma unroll
        for (int kh = 0; kh < kInRows; ++kh) v[kh] = vn[kh];
        tw.next(p);
        if (tile + (int)gridDim.x < p.num_tiles) fetch(tw);      // next tile's pixels travel while this one is converted
      }
      const long long k2 = p.dbg ? clock64() : 0;
      pv_mbar_wait(&bar_empty[slot], phase ^ 1u, p.err, 6);
      const long long k3 = p.dbg ? clock64() : 0;
      uint8_t* dstp = pxb + slot * kPxSlotBytes + ct * 8;
#pragma unroll
      for (int kh = 0; kh < kInRows; ++kh) {
        uint2 o;
        if (p.ablate & 4) {
          o.x = v[kh] & 0x3f003f00u;
          o.y = 0u;
        } else {
          // (v - mean)/256 == fma(v, 2^-8, -mean*2^-8) bit for bit (power-of-two scaling commutes with rounding)
          const float r = fmaf((float)(v[kh] & 255u), 0.00390625f, p.c0);
          const float g = fmaf((float)((v[kh] >> 8) & 255u), 0.00390625f, p.c1);
          const float b = fmaf((float)((v[kh] >> 16) & 255u), 0.00390625f, p.c2);
          const bool a = (v[kh] >> 24) != 0u;          // alpha 0 = pyramid padding / TMA zero fill -> exact 0
          o.x = a ? pv_pack_bf16x2(r, g) : 0u;
          o.y = a ? pv_pack_bf16x2(b, 0.f) : 0u;
        }
        if (!(p.ablate & 16)) *reinterpret_cast<uint2*>(dstp + kh * kPxRowBytes) = o;
        else if (o.x == 0x7fffffffu) p.err[0] = 2;
      }
      if (!(p.ablate & 32)) pv_fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        pv_mbar_arrive(&bar_full[slot]);
        if (p.use_tma) pv_mbar_arrive(&bar_rempty[rs]);
      }
      if (++rs == kRawRing) { rs = 0; rphase ^= 1u; }
      if (++slot == kPxRing) { slot = 0; phase ^= 1u; }
      if (p.dbg) { k_wr += k1 - k0; k_we += k3 - k2; k_work += (k2 - k1) + (clock64() - k3); }
    }
    if (p.dbg && threadIdx.x == 64) {
      long long* d = p.dbg + (long long)blockIdx.x * 8;
      d[4] = k_wr; d[5] = k_we; d[6] = k_work;
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const uint32_t m = (uint32_t)(quarter * 32 + lane);
    TileWalk t;
    t.init(p, blockIdx.x, gridDim.x);
    int buf = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t x = t.ct * kTileOut + m;
      const bool xvalid = (m < (uint32_t)kTileOut) && (x < (uint32_t)p.ow);
      const uint32_t oy0 = t.oy * kRows;
      const long long q0 = p.dbg ? clock64() : 0;
      if (p.ablate & 128) pv_mbar_wait(&bar_tfull[buf], aphase, p.err, 4);
      else pv_mbar_wait_backoff(&bar_tfull[buf], aphase, p.err, 4, 32);
      pv_tc_fence_after();
      if (p.dbg && threadIdx.x == 320) p.dbg[(long long)blockIdx.x * 8 + 7] += clock64() - q0;
      // two rows per pass: two TMEM loads in flight, scale / shift come from shared memory as 16-byte ld.shared
      const uint32_t sc_addr = pv_smem_u32(s_scale), sh_addr = pv_smem_u32(s_shift);
#pragma unroll 1
      for (int r = 0; r < kRows; r += 2) {
        uint32_t v[2][16];
        const uint32_t tcol = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((buf * kRows + r) * kN);
        if (!(p.ablate & 64)) {
          pv_tmem_ld16(tcol, v[0]);
          pv_tmem_ld16(tcol + kN, v[1]);
          pv_tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[0][j] = v[1][j] = (uint32_t)j;
        }
        if (r == kRows - 2) {
          pv_tc_fence_before();
          __syncwarp();
          if (lane == 0) pv_mbar_arrive(&bar_tempty[buf]);   // last accumulators are in registers: release the buffer
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (xvalid && oy0 + r + h < (uint32_t)p.oh && !(p.ablate & 2)) {
            const long long drow = row_of(p.dst, t.n, oy0 + r + h, x);
            uint32_t o[8];
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const uint4 sc = pv_lds128(sc_addr + g4 * 16), sh = pv_lds128(sh_addr + g4 * 16);
              float f0 = fmaf(__uint_as_float(v[h][4 * g4 + 0]), __uint_as_float(sc.x), __uint_as_float(sh.x));
              float f1 = fmaf(__uint_as_float(v[h][4 * g4 + 1]), __uint_as_float(sc.y), __uint_as_float(sh.y));
              float f2 = fmaf(__uint_as_float(v[h][4 * g4 + 2]), __uint_as_float(sc.z), __uint_as_float(sh.z));
              float f3 = fmaf(__uint_as_float(v[h][4 * g4 + 3]), __uint_as_float(sc.w), __uint_as_float(sh.w));
              if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
              o[2 * g4] = pv_pack_bf16x2(f0, f1);
              o[2 * g4 + 1] = pv_pack_bf16x2(f2, f3);
            }
            if (!(p.ablate & 1)) pv_stg256(p.out + drow * p.dst.cols, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
            else if (o[0] == 0x12345u) p.err[0] = 1;   // keep the math alive
          }
        }
      }
      if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
      t.next(p);
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, kAcc * kRows * kN);
  }
}

constexpr size_t kSmemBytes = 128 + kPxRing * kPxSlotBytes + kRawRing * kRawBytes + kWBytes +
                              (2 * kRawRing + 2 * kPxRing + 2 * kAcc) * 8 + 2 * kN * 4 + 64;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

long long* g_c1_dbg = nullptr;

}  // namespace

/* role timing of the last pv_conv1_fused launch when PV_C1_DEBUG is set: out8 (HOST) = cycles summed over CTAs
 * {MMA: wait accumulator, wait pixels, issue, tiles; converter: wait raw, wait slot, work; epilogue: wait} */
extern "C" int pv_conv1_debug(long long* out8) {
  PV_REQUIRE(out8 && g_c1_dbg, "pv_conv1_debug: PV_C1_DEBUG was not set");
  static long long h[8 * 1024];
  PV_CUDA_CHECK(cudaMemcpy(h, g_c1_dbg, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 8; ++k) out8[k] = 0;
  for (int i = 0; i < 1024; ++i)
    for (int k = 0; k < 8; ++k) out8[k] += h[8 * i + k];
  return PV_OK;
}

extern "C" int pv_conv1_fused(const void* plane_rgba, int B, int Hp, int Wp, const void* w_bf16, const float* scale,
                              const float* shift, int relu, void* out, const PvRowMap* dst, int oh, int ow,
                              const float* mean_host, int* err_flag, void* stream) {
  PV_REQUIRE(plane_rgba && w_bf16 && scale && shift && out && dst && mean_host && err_flag, "pv_conv1_fused: null argument");
  PV_REQUIRE(dst->cols >= kN && dst->cols % 16 == 0, "pv_conv1_fused: dst.cols=%d (rows are written with 32-byte stores)", dst->cols);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(out) & 31) == 0, "pv_conv1_fused: output must be 32-byte aligned");
  PV_REQUIRE(B > 0 && Hp >= kKH && Wp >= 5 && oh > 0 && ow > 0, "pv_conv1_fused: bad extent B=%d Hp=%d Wp=%d", B, Hp, Wp);
  PV_REQUIRE(2 * (oh - 1) + kKH <= Hp && 2 * (ow - 1) + 5 <= Wp, "pv_conv1_fused: output %dx%d does not fit plane %dx%d", oh, ow, Hp, Wp);
  PV_REQUIRE(Wp % 4 == 0, "pv_conv1_fused: plane width %d must be a multiple of 4 pixels (16-byte TMA row pitch)", Wp);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(plane_rgba) & 15) == 0, "pv_conv1_fused: plane must be 16-byte aligned");
  static EncodeTiledFn encode = nullptr;
  static int num_sms = 0;
  if (!encode) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PV_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qres));
    PV_REQUIRE(qres == cudaDriverEntryPointSuccess && fp, "pv_conv1_fused: cuTensorMapEncodeTiled unavailable");
    int dev = 0;
    PV_CUDA_CHECK(cudaGetDevice(&dev));
    PV_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    PV_CUDA_CHECK(cudaFuncSetAttribute(conv1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    encode = reinterpret_cast<EncodeTiledFn>(fp);
  }
  C1Params p;
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Wp, (cuuint64_t)B * (cuuint64_t)Hp};
    cuuint64_t gstride[1] = {(cuuint64_t)Wp * 4};
    cuuint32_t box[2] = {kRawW, kInRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.raw, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(plane_rgba), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pv_set_error("pv_conv1_fused: cuTensorMapEncodeTiled failed (CUresult %d, Wp=%d rows=%lld)", (int)r, Wp, (long long)B * Hp);
      return PV_ERR_CUDA;
    }
  }
  p.plane = static_cast<const uint32_t*>(plane_rgba);
  {
    const char* ut = getenv("PV_C1_TMA");
    p.use_tma = (ut && ut[0] == '0') ? 0 : 1;    // default: TMA raw-pixel ring; PV_C1_TMA=0 selects the direct-load variant
  }
  p.B = B;
  p.Hp = Hp;
  p.Wp = Wp;
  p.oh = oh;
  p.ow = ow;
  p.tiles_per_row = (ow + kTileOut - 1) / kTileOut;
  p.w = static_cast<const __nv_bfloat16*>(w_bf16);
  p.scale = scale;
  p.shift = shift;
  p.relu = relu;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.dst = *dst;
  p.c0 = -mean_host[0] * 0.00390625f;
  p.c1 = -mean_host[1] * 0.00390625f;
  p.c2 = -mean_host[2] * 0.00390625f;
  p.oh_tiles = (oh + kRows - 1) / kRows;
  const long long nt = (long long)B * p.oh_tiles * p.tiles_per_row;
  PV_REQUIRE(nt < (1ll << 31), "pv_conv1_fused: too many tiles");
  p.num_tiles = (int)nt;
  p.err = err_flag;
  static long long* d_dbg = nullptr;
  if (getenv("PV_C1_DEBUG") && !d_dbg) cudaMalloc(&d_dbg, sizeof(long long) * 8 * 1024);
  if (d_dbg) cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 8 * 1024, static_cast<cudaStream_t>(stream));
  p.dbg = d_dbg;
  {
    const char* ab = getenv("PV_C1_ABLATE");
    p.ablate = ab ? atoi(ab) : 0;
  }
  g_c1_dbg = d_dbg;
  int grid = num_sms * 2;
  if (grid > p.num_tiles) grid = p.num_tiles;
  conv1_fused_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
