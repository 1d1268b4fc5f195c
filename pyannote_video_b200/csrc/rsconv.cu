// rsconv.cu — "row-streaming" convolution on tcgen05 (sm_100a): the detector's conv layers 2..7.
//
// Replaces the `con` layers of dlib's CNN/MMOD face detector behind face_detector_(rgb, 1),
// pyannote/video/face/face.py:66.  Same contract as csrc/detconv.cu (plain NHWC bf16 activations, TMA
// out-of-bounds zero fill = the convolution's zero padding), different decomposition:
//
// With pixels as the M operand a 128 x 16 A tile (4 KB) is read from shared memory for every MMA, so an MMA
// with N <= 64 output channels is bound by the shared-memory port (32 cycles), not by the tensor core
// (N/2 cycles): the detector's layers (N = 16..48) cannot pass ~50 % of the tensor peak that way
// (profiles/README.md).  Here ONE input row feeds SEVERAL output rows in a single MMA:
//
//   work item   = 128 output columns x L output rows of one image (a vertical strip segment)
//   TMEM        = a ring of row slots, slot = the 128 x NC accumulator of one output row (NC = padded Cout)
//   input row y = one TMA box (128 + KW - 1 pixels) -> for every (kw, 16-channel chunk) one MMA
//                   D[slots of rows r_lo..r_hi] += A[row y shifted by kw] x [W_kh(r_lo) | ... | W_kh(r_hi)]
//                 with kh(r) = y + pad - stride * r: the B operand is the filter column kw of all KH (stride 1)
//                 or ceil(KH/2) (stride 2) filter rows side by side, N = up to 5 x 48 = 240.
//
// A 5x5 stride-1 layer therefore issues 15 MMAs of N = 240 per 128 outputs (tensor-bound, 120 cycles each)
// instead of 75 MMAs of N = 48 (shared-memory-bound, 44 cycles each), reads every input pixel once
// (132/128) and needs no im2col of any kind.  The first contribution to a row (kh = 0, kw = 0, chunk 0) is
// issued as its own MMA with accumulate = 0, so slots never have to be cleared.  An output row is complete
// when the input row of its last filter row has been issued: one tcgen05.commit per row hands the slot to the
// epilogue warps, which drain it (tcgen05.ld -> affine -> ReLU -> bf16 -> NHWC) while later rows accumulate.
//
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM owner), 2..5 = epilogue.  Weights stay resident in shared
// memory; stride 2 uses pixel-PAIR operand rows (2C channels), the tap's column parity selects the K offset.
#include <cuda.h>
#include <atomic>
#include <cstdlib>
#include <vector>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kTileW = 128;
constexpr int kMaxStages = 8;
constexpr int kMaxSlots = 16;
constexpr int kThreads = 192;

struct RsParams {
  CUtensorMap in[2];
  const uint8_t* w_img;
  uint32_t w_bytes;
  const float* scale;
  const float* shift;
  void* out;
  int B, OH, OW;
  int out_pitch, out_cs;
  int strips, segs, seg_rows;   // work items = B * segs * strips
  int num_items;
  int relu;
  int n_stages;
  int* err;
  long long* dbg;   // optional [grid][8] cycle counters (role timing), nullptr = off
  // embedder extensions: residual added before the ReLU (same geometry as `out`), and "gap" columns — images that hold
  // several faces side by side, every gap_period-th column (gap_pos) is the zero column between two faces and must
  // stay zero in the output (it is the padding of both neighbours)
  const __nv_bfloat16* resid;
  int gap_period, gap_pos;
};

__host__ __device__ constexpr int rs_align1k(int v) { return (v + 1023) & ~1023; }
__host__ __device__ constexpr uint32_t rs_layout_of_rowb(int rowb) { return rowb == 128 ? 2u : (rowb == 64 ? 4u : 6u); }
// hi word of a K-major swizzled A descriptor: SBO = 8 rows, version 1, swizzle mode
__host__ __device__ constexpr uint32_t rs_a_desc_hi(int rowb) {
  return (uint32_t)(((8 * rowb) >> 4) & 0x3FFF) | (1u << 14) | (rs_layout_of_rowb(rowb) << 29);
}

__device__ __forceinline__ void rs_tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(pv_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(pv_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void rs_bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   pv_smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(pv_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void rs_umma(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t accumulate) {
  const uint64_t da = ((uint64_t)a_hi << 32) | a_lo;
  const uint64_t db = ((uint64_t)b_hi << 32) | b_lo;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

template <int C, int NC, int KH, int KW, int S>
struct RsGeo {
  static constexpr int kRowEl0 = S == 2 ? 2 * C : (C >= 32 ? 32 : 16);
  static constexpr int kRowEl1 = S == 2 ? 0 : C - kRowEl0;
  static constexpr int kSegs = kRowEl1 > 0 ? 2 : 1;
  static constexpr int kRowB0 = kRowEl0 * 2, kRowB1 = kRowEl1 * 2;
  static constexpr int kAW = kTileW + (S == 2 ? (KW - 1) / 2 : KW - 1);   // operand rows per input row
  static constexpr int kSeg0Bytes = rs_align1k(kAW * kRowB0);
  static constexpr int kSeg1Bytes = kSegs > 1 ? rs_align1k(kAW * kRowB1) : 0;
  static constexpr int kStageBytes = kSeg0Bytes + kSeg1Bytes;
  static constexpr int kKch = C / 16;
  static constexpr int kPadY = S == 1 ? KH / 2 : 0, kPadX = S == 1 ? KW / 2 : 0;
  static constexpr int kNQ0 = (KH + S - 1) / S;            // filter rows of parity class 0 (kh = 0, S, 2S, ...)
  static constexpr int kNQ1 = S == 2 ? KH / 2 : 0;          // parity class 1 (kh = 1, 3, ...)
  static constexpr int kTile0Bytes = kNQ0 * NC * 32;        // one (kw, chunk) weight tile of class 0
  static constexpr int kTile1Bytes = kNQ1 * NC * 32;
  static constexpr int kQ1Base = KW * kKch * kTile0Bytes;
  static constexpr int kWBytes = KW * kKch * (kTile0Bytes + kTile1Bytes);
  // layers with <= 32 output channels run two CTAs per SM (each half of TMEM and of shared memory): their rows
  // carry few MMAs, so a second issuing thread hides the per-row hand-offs of the first — unless the resident weights
  // (the HOG detector's 10 x 10 x 32 filters: 100 KB) leave no room for two input rings
  static constexpr int kCtas = (NC <= 32 && kWBytes <= 64 * 1024) ? 2 : 1;
  static constexpr int kTmemBudget = 512 / kCtas;
  static constexpr int kSlots = (kTmemBudget / NC) < kMaxSlots ? (kTmemBudget / NC) : kMaxSlots;
  static constexpr uint32_t kTmemCols = kSlots * NC > 256 ? 512u : (kSlots * NC > 128 ? 256u : 128u);
};

// C: channels per input pixel in memory (16/32/48), NC: padded output channels, KH x KW filter, S stride
// (1: pad = K/2 on both axes, 2: pad 0), F32: fp32 output rows (last layer)
template <int C, int NC, int KH, int KW, int S, bool F32>
__global__ void __launch_bounds__(kThreads, (RsGeo<C, NC, KH, KW, S>::kCtas)) rsconv_kernel(const __grid_constant__ RsParams p) {
  using G = RsGeo<C, NC, KH, KW, S>;
  constexpr int NSLOT = G::kSlots;
  static_assert(NC % 16 == 0 && NC >= 16 && NC <= 64, "NC");
  static_assert(NSLOT >= G::kNQ0 + 2, "TMEM row ring too small");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + rs_align1k(G::kWBytes);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ring + (size_t)p.n_stages * G::kStageBytes);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_rfull = bar_empty + kMaxStages;
  uint64_t* bar_rempty = bar_rfull + kMaxSlots;
  uint64_t* bar_w = bar_rempty + kMaxSlots;
  float* s_scale = reinterpret_cast<float*>(bar_w + 1);
  float* s_shift = s_scale + NC;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_shift + NC);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < NC; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 0 && lane == 0) {
    pv_tma_prefetch_desc(&p.in[0]);
    if (G::kSegs > 1) pv_tma_prefetch_desc(&p.in[1]);
    for (int i = 0; i < p.n_stages; ++i) {
      pv_mbar_init(&bar_full[i], 1);
      pv_mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < NSLOT; ++i) {
      pv_mbar_init(&bar_rfull[i], 1);
      pv_mbar_init(&bar_rempty[i], 4);
    }
    pv_mbar_init(bar_w, 1);
    pv_fence_mbar_init();
  }
  if (warp == 1) pv_tmem_alloc(s_tmem, G::kTmemCols);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int items_per_img = p.segs * p.strips;

  if (warp == 0) {
    // ===================== TMA producer: one input row per stage =====================
    if (pv_elect_one()) {
      pv_mbar_arrive_expect_tx(bar_w, (uint32_t)G::kWBytes);
      for (uint32_t off = 0; off < (uint32_t)G::kWBytes; off += 32768u) {
        const uint32_t n = (uint32_t)G::kWBytes - off < 32768u ? (uint32_t)G::kWBytes - off : 32768u;
        rs_bulk_load_1d(wsm + off, p.w_img + off, n, bar_w);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int b = item / items_per_img;
        const int r = item - b * items_per_img;
        const int seg = r / p.strips;
        const int strip = r - seg * p.strips;
        const int ra = seg * p.seg_rows;
        const int rb = min(ra + p.seg_rows, p.OH);
        const int ix0 = strip * kTileW - G::kPadX;               // in operand rows (pixels, or pixel pairs for stride 2)
        const int y0 = S * ra - G::kPadY, y1 = S * (rb - 1) - G::kPadY + KH - 1;
        for (int y = y0; y <= y1; ++y) {
          pv_mbar_wait(&bar_empty[stage], phase ^ 1u, p.err, 1);
          uint8_t* dst = ring + (size_t)stage * G::kStageBytes;
          pv_mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(G::kAW * (G::kRowB0 + G::kRowB1)));
          rs_tma_load_4d(dst, &p.in[0], &bar_full[stage], 0, ix0, y, b);
          if (G::kSegs > 1) rs_tma_load_4d(dst + G::kSeg0Bytes, &p.in[1], &bar_full[stage], G::kRowEl0, ix0, y, b);
          if (++stage == p.n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (pv_elect_one()) {
      constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);   // + (N >> 3) << 17
      constexpr uint32_t AHI0 = rs_a_desc_hi(G::kRowB0);
      constexpr uint32_t AHI1 = rs_a_desc_hi(G::kRowB1 > 0 ? G::kRowB1 : 32);
      // B tiles are K-major rows of 32 bytes (16 channels) in the 32-byte swizzle (the un-swizzled core-matrix layout is
      // fetched at half the shared-memory rate: measured, profiles/README.md): SBO = 8 rows = 256 B, version 1
      constexpr uint32_t BHI = (uint32_t)(256 >> 4) | (1u << 14) | (6u << 29);
      constexpr uint32_t BLK16 = (uint32_t)(NC * 32 >> 4);          // one output row's block of a B tile: NC rows x 32 B
      constexpr int DC = (KH - 1) / S;                                // rows between the newest row and the one that completes
      pv_mbar_wait(bar_w, 0, p.err, 5);
      pv_tc_fence_after();
      const uint32_t w_addr16 = (pv_smem_u32(wsm) & 0x3FFFFu) >> 4;
      const uint32_t ring_lo = ((pv_smem_u32(ring) & 0x3FFFFu) >> 4) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      int slot_base = 0;                      // slot / ring parity of the item's first output row
      uint32_t par_base = 0;
      long long c_rempty = 0, c_full = 0, c_issue = 0, c_rows = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int r_img = item % items_per_img;
        const int seg = r_img / p.strips;
        const int ra = seg * p.seg_rows;
        const int rb = min(ra + p.seg_rows, p.OH);
        const int t0 = S * ra, t1 = S * (rb - 1) + KH - 1;       // t = y + pad
        int slot_top = slot_base;             // slot / parity of row r_top(t) (tracked past rb - 1 as if rows went on)
        uint32_t par_top = par_base;
        // barriers of the item's first input row (later rows are checked while the previous row's MMAs run)
        long long c0 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_rempty[slot_top], par_top ^ 1u, p.err, 2);
        long long c1 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_full[stage], phase, p.err, 3);
        long long c2 = p.dbg ? clock64() : 0;
        for (int t = t0; t <= t1; ++t) {
          pv_tc_fence_after();
          const int q = S == 2 ? (t & 1) : 0;
          const bool odd = S == 2 && q == 1;
          const int nq = odd ? G::kNQ1 : G::kNQ0;
          const int r_top = S == 2 ? (t >> 1) : t;               // row reading this input row with kh = q
          const int r_first = r_top - (nq - 1);                  // row reading it with the largest kh of the class
          const int r_hi = min(r_top, rb - 1);
          const int r_lo = max(r_first, ra);
          const bool new_row = (q == 0) && (r_top <= rb - 1);    // first contribution to row r_top (kh = 0)
          const int n_all = r_hi - r_lo + 1;                     // >= 1 inside the item's range of t
          int slot_lo = slot_top - (r_top - r_lo);
          if (slot_lo < 0) slot_lo += NSLOT;
          // segments of consecutive rows / slots: [r_lo .. r_lo+n1-1] at slot_lo, the rest from slot 0 (ring wrap)
          const int n1 = min(n_all, NSLOT - slot_lo);
          const int n2 = n_all - n1;
          const uint32_t d1 = tmem_base + (uint32_t)(slot_lo * NC);
          const uint32_t id1 = idesc0 | ((uint32_t)(n1 * NC >> 3) << 17);
          const uint32_t id2 = idesc0 | ((uint32_t)(n2 * NC >> 3) << 17);
          // weight tile of this parity class, first block = row r_lo
          const uint32_t tile16 = (uint32_t)((odd ? G::kTile1Bytes : G::kTile0Bytes) >> 4);
          const uint32_t bb1 = (w_addr16 + (odd ? (uint32_t)(G::kQ1Base >> 4) : 0u) + (uint32_t)(r_lo - r_first) * BLK16) | (1u << 16);
          const uint32_t bb2 = bb1 + (uint32_t)n1 * BLK16;
          const uint32_t a0 = ring_lo + (uint32_t)stage * (uint32_t)(G::kStageBytes >> 4);
          const uint32_t a1 = a0 + (uint32_t)(G::kSeg0Bytes >> 4);
          // next input row of this item: its barriers are checked in the middle of this row's MMAs
          const bool has_next = t < t1;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == p.n_stages) { nstage = 0; nphase ^= 1u; }
          const bool top_moves = S == 1 || q == 1;               // r_top(t+1) = r_top(t) + 1
          int nslot = slot_top;
          uint32_t npar = par_top;
          if (top_moves && ++nslot == NSLOT) { nslot = 0; npar ^= 1u; }
          const bool next_new = has_next && top_moves && (r_top + 1 <= rb - 1);
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
#pragma unroll
            for (int kc = 0; kc < G::kKch; ++kc) {
              uint32_t a_lo, a_hi;
              if (S == 2) {
                a_lo = a0 + (uint32_t)(((kw >> 1) * G::kRowB0 + (kw & 1) * C * 2 + kc * 32) >> 4);
                a_hi = AHI0;
              } else if (kc * 16 < G::kRowEl0) {
                a_lo = a0 + (uint32_t)((kw * G::kRowB0 + kc * 32) >> 4);
                a_hi = AHI0;
              } else {
                a_lo = a1 + (uint32_t)((kw * G::kRowB1 + (kc * 16 - G::kRowEl0) * 2) >> 4);
                a_hi = AHI1;
              }
              const uint32_t boff = (uint32_t)(kw * G::kKch + kc) * tile16;
              if (kw == 0 && kc == 0 && new_row) {
                // first contribution to the new row: older rows accumulate, the new row starts from zero
                const int o1 = min(n_all - 1, n1), o2 = n_all - 1 - o1;
                if (o1 > 0) rs_umma(d1, a_lo, a_hi, bb1, BHI, idesc0 | ((uint32_t)(o1 * NC >> 3) << 17), 1u);
                if (o2 > 0) rs_umma(tmem_base, a_lo, a_hi, bb1 + (uint32_t)o1 * BLK16, BHI, idesc0 | ((uint32_t)(o2 * NC >> 3) << 17), 1u);
                rs_umma(tmem_base + (uint32_t)(slot_top * NC), a_lo, a_hi, bb1 + (uint32_t)(n_all - 1) * BLK16, BHI,
                        idesc0 | ((uint32_t)(NC >> 3) << 17), 0u);
              } else {
                rs_umma(d1, a_lo, a_hi, bb1 + boff, BHI, id1, 1u);
                if (n2 > 0) rs_umma(tmem_base, a_lo, a_hi, bb2 + boff, BHI, id2, 1u);
              }
            }
            if (kw == (KW - 1) / 2 && has_next) {
              // the tensor pipe has this row's MMAs queued: look at the next row's barriers now
              c0 = p.dbg ? clock64() : 0;
              if (next_new) pv_mbar_wait(&bar_rempty[nslot], npar ^ 1u, p.err, 2);
              c1 = p.dbg ? clock64() : 0;
              pv_mbar_wait(&bar_full[nstage], nphase, p.err, 3);
              c2 = p.dbg ? clock64() : 0;
              if (p.dbg) { c_rempty += c1 - c0; c_full += c2 - c1; }
            }
          }
          pv_umma_commit(&bar_empty[stage]);
          // the row whose last filter row (kh = KH-1) this was is complete
          if (S == 1 || q == 0) {
            const int rc = r_top - DC;
            if (rc >= ra && rc < rb) {
              int sc = slot_top - DC;
              if (sc < 0) sc += NSLOT;
              pv_umma_commit(&bar_rfull[sc]);
            }
          }
          stage = nstage;
          phase = nphase;
          slot_top = nslot;
          par_top = npar;
          if (p.dbg) ++c_rows;
        }
        if (p.dbg) c_issue += clock64() - c2;
        const int adv = (slot_base + (rb - ra)) / NSLOT;
        slot_base = (slot_base + (rb - ra)) % NSLOT;
        par_base ^= (uint32_t)(adv & 1);
      }
      if (p.dbg) {
        long long* d = p.dbg + (long long)blockIdx.x * 8;
        d[0] = c_rempty; d[1] = c_full; d[2] = c_issue; d[3] = c_rows;
      }
    }
  } else {
    // ===================== epilogue: one output row (128 x NC) per slot =====================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    uint32_t cnt = 0;
    long long e_wait = 0, e_work = 0;
    const long long e_start = p.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int b = item / items_per_img;
      const int r_img = item - b * items_per_img;
      const int seg = r_img / p.strips;
      const int strip = r_img - seg * p.strips;
      const int ra = seg * p.seg_rows;
      const int rb = min(ra + p.seg_rows, p.OH);
      const int ox = strip * kTileW + m;
      const bool xvalid = ox < p.OW;
      const bool gap = p.gap_period > 0 && (ox % p.gap_period) == p.gap_pos;
      for (int r = ra; r < rb; ++r) {
        const uint32_t rel = cnt + (uint32_t)(r - ra);
        const uint32_t slot = rel % NSLOT;
        const long long pix = ((long long)b * p.OH + r) * p.out_pitch + ox;
        // the residual row does not depend on the accumulator: its loads are issued BEFORE the wait for the row's MMAs
        // (behind the wait they doubled the epilogue time of the embedder's second conv of every block)
        uint4 res[NC / 16][2];
        const bool has_res = !F32 && p.resid != nullptr && xvalid;
        if (has_res) {
#pragma unroll
          for (int j = 0; j < NC / 16; ++j) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.resid + pix * p.out_cs + j * 16);
            res[j][0] = __ldg(rp);
            res[j][1] = __ldg(rp + 1);
          }
        }
        const long long e0 = p.dbg ? clock64() : 0;
        pv_mbar_wait(&bar_rfull[slot], (rel / NSLOT) & 1u, p.err, 4);
        pv_tc_fence_after();
        const long long e1 = p.dbg ? clock64() : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * NC;
        uint32_t v[NC / 16][16];
#pragma unroll
        for (int j = 0; j < NC / 16; ++j) pv_tmem_ld16(taddr + j * 16, v[j]);
        pv_tmem_ld_wait();
        pv_tc_fence_before();
        __syncwarp();
        if (lane == 0) pv_mbar_arrive(&bar_rempty[slot]);      // accumulator row is in registers: release the slot
        if (xvalid) {
#pragma unroll
          for (int j = 0; j < NC / 16; ++j) {
            float f[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) f[k] = fmaf(__uint_as_float(v[j][k]), s_scale[j * 16 + k], s_shift[j * 16 + k]);
            if (has_res) {
              const uint32_t rw[8] = {res[j][0].x, res[j][0].y, res[j][0].z, res[j][0].w, res[j][1].x, res[j][1].y, res[j][1].z, res[j][1].w};
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                f[2 * k] += __uint_as_float(rw[k] << 16);
                f[2 * k + 1] += __uint_as_float(rw[k] & 0xFFFF0000u);
              }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (p.relu) f[k] = fmaxf(f[k], 0.f);
              if (gap) f[k] = 0.f;
            }
            if (F32) {
              float* dp = reinterpret_cast<float*>(p.out) + pix * p.out_cs + j * 16;
              pv_stg256(dp, __float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]),
                        __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
              pv_stg256(dp + 8, __float_as_uint(f[8]), __float_as_uint(f[9]), __float_as_uint(f[10]), __float_as_uint(f[11]),
                        __float_as_uint(f[12]), __float_as_uint(f[13]), __float_as_uint(f[14]), __float_as_uint(f[15]));
            } else {
              pv_stg256(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_cs + j * 16, pv_pack_bf16x2(f[0], f[1]),
                        pv_pack_bf16x2(f[2], f[3]), pv_pack_bf16x2(f[4], f[5]), pv_pack_bf16x2(f[6], f[7]), pv_pack_bf16x2(f[8], f[9]),
                        pv_pack_bf16x2(f[10], f[11]), pv_pack_bf16x2(f[12], f[13]), pv_pack_bf16x2(f[14], f[15]));
            }
          }
        }
        if (p.dbg) { e_wait += e1 - e0; e_work += clock64() - e1; }
      }
      cnt += (uint32_t)(rb - ra);
    }
    if (p.dbg && threadIdx.x == 64) {
      long long* d = p.dbg + (long long)blockIdx.x * 8;
      d[4] = e_wait; d[5] = e_work; d[6] = clock64() - e_start;
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, G::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn rs_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

struct RsPlan {
  RsParams p;
  int kind = -1;
  size_t smem_bytes = 0;
  int num_sms = 0;
  int Bmax = 0;
  int ctas = 1;
  int* d_err = nullptr;
  long long* d_dbg = nullptr;
};

struct RsInstance {
  int C, N, KH, KW, S, f32;
};
constexpr RsInstance kRsInstances[] = {
    {16, 32, 5, 5, 2, 0},   // conv2
    {32, 32, 5, 5, 2, 0},   // conv3
    {32, 48, 5, 5, 1, 0},   // conv4
    {48, 48, 5, 5, 1, 0},   // conv5, conv6
    {48, 16, 9, 1, 1, 1},   // conv7 as 9x1 (filter columns as output channels), fp32 rows
    {32, 32, 3, 3, 1, 0},   // embedder level 4 (35 x 35, faces side by side)
    {32, 64, 3, 3, 2, 0},   // embedder level 3 entry (stride 2)
    {64, 64, 3, 3, 1, 0},   // embedder level 3 (17 x 17)
    {32, 16, 10, 10, 1, 1}, // HOG detector: up to 16 linear filters of 10 x 10 cells x 31(+1) features, fp32 scores
};
constexpr int kNumRsInstances = sizeof(kRsInstances) / sizeof(kRsInstances[0]);

template <int C, int N, int KH, int KW, int S, bool F32>
cudaError_t rs_launch(const RsPlan* plan, const RsParams& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (pv_attr_needed(&attr)) {
    cudaError_t e = cudaFuncSetAttribute(rsconv_kernel<C, N, KH, KW, S, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
  }
  rsconv_kernel<C, N, KH, KW, S, F32><<<grid, kThreads, plan->smem_bytes, st>>>(p);
  return cudaGetLastError();
}

template <int C, int N, int KH, int KW, int S>
void rs_geometry(int* aw, int* rowel0, int* rowel1, int* stage_bytes, int* w_bytes, int* ctas) {
  using G = RsGeo<C, N, KH, KW, S>;
  *ctas = G::kCtas;
  *aw = G::kAW;
  *rowel0 = G::kRowEl0;
  *rowel1 = G::kRowEl1;
  *stage_bytes = G::kStageBytes;
  *w_bytes = G::kWBytes;
}

}  // namespace

extern "C" int pv_rsconv_create(const PvDetconvDesc* d, void** out_handle) {
  PV_REQUIRE(d && out_handle, "pv_rsconv_create: null argument");
  PV_REQUIRE(d->x && d->w_img && d->scale && d->shift && d->out, "pv_rsconv_create: null operand");
  int kind = -1;
  for (int i = 0; i < kNumRsInstances; ++i) {
    const RsInstance& in = kRsInstances[i];
    if (in.C == d->c_in && in.N == d->n_out && in.KH == d->kh && in.KW == d->kw && in.S == d->stride && in.f32 == d->out_f32)
      kind = i;
  }
  PV_REQUIRE(kind >= 0, "pv_rsconv_create: no kernel instance for C=%d N=%d %dx%d stride %d f32=%d", d->c_in, d->n_out,
             d->kh, d->kw, d->stride, d->out_f32);
  const RsInstance& in = kRsInstances[kind];
  PV_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->pitch >= d->W && d->pitch % 2 == 0, "pv_rsconv_create: bad input extent");
  const int pad_y = in.S == 1 ? in.KH / 2 : 0, pad_x = in.S == 1 ? in.KW / 2 : 0;
  const int OH = (d->H + 2 * pad_y - in.KH) / in.S + 1, OW = (d->W + 2 * pad_x - in.KW) / in.S + 1;
  PV_REQUIRE(OH > 0 && OW > 0, "pv_rsconv_create: empty output");
  PV_REQUIRE(d->out_pitch >= OW && d->out_cs >= in.N && d->out_cs % 16 == 0, "pv_rsconv_create: output pitch %d / channel stride %d (rows are written with 32-byte stores)",
             d->out_pitch, d->out_cs);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w_img) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->out) & 31) == 0, "pv_rsconv_create: operands must be 16-byte (output: 32-byte) aligned");
  int aw, rowel0, rowel1, stage_bytes, w_bytes, ctas;
  switch (kind) {
    case 0: rs_geometry<16, 32, 5, 5, 2>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 1: rs_geometry<32, 32, 5, 5, 2>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 2: rs_geometry<32, 48, 5, 5, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 3: rs_geometry<48, 48, 5, 5, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 4: rs_geometry<48, 16, 9, 1, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 5: rs_geometry<32, 32, 3, 3, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 6: rs_geometry<32, 64, 3, 3, 2>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    case 7: rs_geometry<64, 64, 3, 3, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
    default: rs_geometry<32, 16, 10, 10, 1>(&aw, &rowel0, &rowel1, &stage_bytes, &w_bytes, &ctas); break;
  }
  PV_REQUIRE(d->w_bytes == (int64_t)w_bytes, "pv_rsconv_create: weight image is %lld bytes, expected %d", (long long)d->w_bytes, w_bytes);
  const size_t fixed = (2 * kMaxStages + 2 * kMaxSlots + 1) * sizeof(uint64_t) + 2 * in.N * sizeof(float) + 64;
  const long long budget = (227 * 1024) / ctas - 1024 - 1024 - (long long)fixed - rs_align1k(w_bytes);
  int n_stages = (int)(budget / stage_bytes);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  PV_REQUIRE(n_stages >= 3, "pv_rsconv_create: input row of %d bytes (+%d weights) does not fit a 3-deep ring", stage_bytes, w_bytes);

  EncodeTiledFn enc = rs_encode_fn();
  if (!enc) {
    pv_set_error("pv_rsconv_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return PV_ERR_CUDA;
  }
  RsPlan* plan = new RsPlan();
  memset(&plan->p, 0, sizeof(RsParams));
  RsParams& p = plan->p;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    pv_set_error("pv_rsconv_create: no CUDA device: %s", cudaGetErrorString(e));
    delete plan;
    return PV_ERR_CUDA;
  }
  for (int s = 0; s < (rowel1 > 0 ? 2 : 1); ++s) {
    const int segw = s == 0 ? rowel0 : rowel1;
    const cuuint64_t row_el = in.S == 2 ? 2 * (cuuint64_t)in.C : (cuuint64_t)in.C;
    const cuuint64_t npx = in.S == 2 ? (cuuint64_t)d->pitch / 2 : (cuuint64_t)d->pitch;
    cuuint64_t gdim[4] = {row_el, npx, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t gstr[3] = {row_el * 2, (cuuint64_t)d->pitch * in.C * 2, (cuuint64_t)d->H * d->pitch * in.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)segw, (cuuint32_t)aw, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const int rowb = segw * 2;
    CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(&p.in[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pv_set_error("pv_rsconv_create: cuTensorMapEncodeTiled failed: CUresult %d (seg %d, C=%d pitch=%d H=%d B=%d box %dx%d)", (int)r, s,
                   in.C, d->pitch, d->H, d->B, segw, aw);
      delete plan;
      return PV_ERR_CUDA;
    }
  }
  if (cudaMalloc(&plan->d_err, sizeof(int)) != cudaSuccess) {
    pv_set_error("pv_rsconv_create: cudaMalloc failed");
    delete plan;
    return PV_ERR_CUDA;
  }
  cudaMemset(plan->d_err, 0, sizeof(int));
  if (getenv("PV_RS_DEBUG")) {
    cudaMalloc(&plan->d_dbg, sizeof(long long) * 8 * plan->num_sms * 2);
    cudaMemset(plan->d_dbg, 0, sizeof(long long) * 8 * plan->num_sms * 2);
  }
  p.dbg = plan->d_dbg;
  p.w_img = static_cast<const uint8_t*>(d->w_img);
  p.w_bytes = (uint32_t)w_bytes;
  p.scale = d->scale;
  p.shift = d->shift;
  p.out = d->out;
  p.B = d->B;
  p.OH = OH;
  p.OW = OW;
  p.out_pitch = d->out_pitch;
  p.out_cs = d->out_cs;
  p.strips = (OW + kTileW - 1) / kTileW;
  // rows per work item: items are dealt round-robin to the persistent CTAs, so the launch lasts as long as the CTA
  // with the most items; pick the segment height that minimises  ceil(items / CTAs) x (input rows per item)
  // among heights whose halo (KH - 1 extra input rows) stays below ~25 % of the item
  {
    const int grid_max = plan->num_sms * ctas;
    const int halo = in.KH - 1;
    const int min_rows = 4 * (halo / in.S) > 8 ? 4 * (halo / in.S) : 8;
    long long best_cost = -1;
    int best_rows = OH;
    for (int segs = 1; segs <= OH; ++segs) {
      const int rows = (OH + segs - 1) / segs;
      if (rows < min_rows && segs > 1) break;
      const int nseg = (OH + rows - 1) / rows;
      const long long items = (long long)d->B * nseg * p.strips;
      const long long waves = (items + grid_max - 1) / grid_max;
      const long long cost = waves * ((long long)rows * in.S + halo);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_rows = rows;
      }
    }
    p.seg_rows = best_rows;
    p.segs = (OH + best_rows - 1) / best_rows;
  }
  p.relu = d->relu;
  p.resid = static_cast<const __nv_bfloat16*>(d->resid);
  p.gap_period = d->gap_period;
  p.gap_pos = d->gap_pos;
  PV_REQUIRE(!d->resid || (reinterpret_cast<uintptr_t>(d->resid) & 31) == 0, "pv_rsconv_create: residual must be 32-byte aligned");
  p.n_stages = n_stages;
  p.err = plan->d_err;
  plan->kind = kind;
  plan->ctas = ctas;
  plan->Bmax = d->B;
  plan->smem_bytes = (size_t)rs_align1k(w_bytes) + (size_t)n_stages * stage_bytes + fixed + 1024;
  *out_handle = plan;
  return PV_OK;
}

extern "C" int pv_rsconv_run(void* handle, int B, void* stream) {
  PV_REQUIRE(handle, "pv_rsconv_run: null handle");
  RsPlan* plan = static_cast<RsPlan*>(handle);
  PV_REQUIRE(B > 0 && B <= plan->Bmax, "pv_rsconv_run: B=%d outside [1,%d]", B, plan->Bmax);
  RsParams p = plan->p;
  p.B = B;
  const long long ni = (long long)B * p.segs * p.strips;
  PV_REQUIRE(ni < (1ll << 31), "pv_rsconv_run: too many work items");
  p.num_items = (int)ni;
  const int max_grid = plan->num_sms * plan->ctas;
  const int grid = p.num_items < max_grid ? p.num_items : max_grid;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (plan->kind) {
    case 0: e = rs_launch<16, 32, 5, 5, 2, false>(plan, p, grid, st); break;
    case 1: e = rs_launch<32, 32, 5, 5, 2, false>(plan, p, grid, st); break;
    case 2: e = rs_launch<32, 48, 5, 5, 1, false>(plan, p, grid, st); break;
    case 3: e = rs_launch<48, 48, 5, 5, 1, false>(plan, p, grid, st); break;
    case 4: e = rs_launch<48, 16, 9, 1, 1, true>(plan, p, grid, st); break;
    case 5: e = rs_launch<32, 32, 3, 3, 1, false>(plan, p, grid, st); break;
    case 6: e = rs_launch<32, 64, 3, 3, 2, false>(plan, p, grid, st); break;
    case 7: e = rs_launch<64, 64, 3, 3, 1, false>(plan, p, grid, st); break;
    default: e = rs_launch<32, 16, 10, 10, 1, true>(plan, p, grid, st); break;
  }
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(e);
  return PV_OK;
}

extern "C" int pv_rsconv_info(void* handle, int* n_stages, int* smem_bytes, int* strips, int* segs, int* seg_rows) {
  PV_REQUIRE(handle, "pv_rsconv_info: null handle");
  RsPlan* plan = static_cast<RsPlan*>(handle);
  if (n_stages) *n_stages = plan->p.n_stages;
  if (smem_bytes) *smem_bytes = (int)plan->smem_bytes;
  if (strips) *strips = plan->p.strips;
  if (segs) *segs = plan->p.segs;
  if (seg_rows) *seg_rows = plan->p.seg_rows;
  return PV_OK;
}

/* role timing of the last launch (PV_RS_DEBUG=1 at create time): out[8] = sums over CTAs of
 * {mma: wait slot, wait data, issue, rows; epilogue: wait row, work, total; -} in cycles */
extern "C" int pv_rsconv_debug(void* handle, long long* out8) {
  PV_REQUIRE(handle && out8, "pv_rsconv_debug: null argument");
  RsPlan* plan = static_cast<RsPlan*>(handle);
  PV_REQUIRE(plan->d_dbg, "pv_rsconv_debug: plan was created without PV_RS_DEBUG");
  std::vector<long long> h(8 * plan->num_sms * 2);
  PV_CUDA_CHECK(cudaMemcpy(h.data(), plan->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 8; ++k) out8[k] = 0;
  for (int i = 0; i < plan->num_sms * 2; ++i)
    for (int k = 0; k < 8; ++k) out8[k] += h[8 * i + k];
  return PV_OK;
}

extern "C" int pv_rsconv_check(void* handle, void* stream) {
  PV_REQUIRE(handle, "pv_rsconv_check: null handle");
  RsPlan* plan = static_cast<RsPlan*>(handle);
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&flag, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    pv_set_error("pv_rsconv_check: %s", cudaGetErrorString(e));
    return PV_ERR_CUDA;
  }
  if (flag != 0) {
    pv_set_error("rsconv: device-side pipeline timeout (role code %d)", flag);
    cudaMemset(plan->d_err, 0, sizeof(int));
    return PV_ERR_DEVICE_TIMEOUT;
  }
  return PV_OK;
}

extern "C" int pv_rsconv_destroy(void* handle) {
  if (!handle) return PV_OK;
  RsPlan* plan = static_cast<RsPlan*>(handle);
  cudaFree(plan->d_err);
  if (plan->d_dbg) cudaFree(plan->d_dbg);
  delete plan;
  return PV_OK;
}
