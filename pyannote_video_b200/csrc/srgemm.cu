// srgemm.cu — shifted-row GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces the convolution arithmetic inside dlib's CNN face detector and
// face_recognition_model_v1 that the reference reaches through
// pyannote/video/face/face.py:66 and :74-75.  See include/pv_b200.h for the
// contract and DESIGN.md §"srgemm" for the data layout.
//
// One persistent CTA per SM, 6 warps:
//   warp 0   TMA producer: activation slabs (+ weights, unless resident) -> ring of smem slots
//   warp 1   tcgen05.mma issuer (one elected lane), accumulators in a ring of TMEM buffers
//   warps 2-5 epilogue: tcgen05.ld -> affine (+residual) (+ReLU) -> bf16/f32 -> global
// A ring slot holds several table entries (slab + taps) so one mbarrier round trip covers up to
// ~48 KB of operands; weights of small layers stay resident in shared memory for the whole launch.
#include <cuda.h>
#include <atomic>
#include <vector>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kThreadsBase = 160;   // producer warp + 4 epilogue warps; + 32 per MMA-issuing warp
constexpr int kMaxRing = 8;
constexpr int kMaxAcc = 8;
constexpr int kTileM = 128;

// One record per tcgen05.mma of a tile, expanded on the host from the entry table so that the
// single issuing lane does no address arithmetic beyond two adds:
//   x = A operand offset inside the ring slot (16-byte units), y = B operand offset (inside the slot,
//   or inside the resident weight region), z = flags:
//   bit0 wait for the slot's TMA data before issuing, bit1 release the slot after issuing,
//   bit2 segment class, bit3 accumulate, bits 8.. index of the TMEM accumulator of this tile.
struct MmaRec {
  uint32_t a_rel16, b_rel16, flags, pad;
};
constexpr int kMaxMma = 160;   // records live in the kernel parameter block (constant bank -> uniform loads)

__host__ __device__ inline uint32_t desc_hi(int width) {
  const uint32_t rowb = (uint32_t)width * 2u;
  const uint32_t layout = width == 64 ? 2u : (width == 32 ? 4u : 6u);
  return ((8u * rowb) >> 4) | (1u << 14) | (layout << 29);   // SBO, version = 1, swizzle mode
}

struct SrParams {
  CUtensorMap a_main[2];
  CUtensorMap a_tail[2];
  CUtensorMap b[2];
  const PvSrEntry* entries;
  int n_mma;
  const float* scale;
  const float* shift;
  void* out;
  const __nv_bfloat16* resid;
  PvRowMap dst;
  PvRowMap res;
  long long q_rows;
  int num_tiles;
  int n_out, n_entries, n_ring, slot_bytes, tail_rows;
  int cls_width[2];
  int resident;          // weights live in smem for the whole launch
  int res_bytes;         // bytes of the resident region (0 if streamed)
  int res_cls_off[2];    // byte offset of each class inside the resident region
  int res_tiles[2];      // number of N-row weight tiles per class
  int n_acc;             // TMEM accumulator buffers (tiles in flight)
  int acc_split;         // accumulators per tile: consecutive MMAs rotate over them (independent chains)
  int acc_cols;          // TMEM columns per tile = acc_split * n_out
  int hq, wq, oh, ow, relu, out_mode, has_resid;
  uint32_t tmem_cols;
  int* err;
  int n_mma_w[4];        // records per MMA warp
  uint4 mma[kMaxMma];    // per-tile MMA records, warp w's list starts at w * (kMaxMma / MMAW)
                         // (x,y operand offsets, z flags | tmem offset << 16, w descriptor hi)
};

__device__ __forceinline__ long long row_of(const PvRowMap& m, uint32_t n, uint32_t y, uint32_t x) {
  const uint32_t Y = y + m.py, X = x + m.px;
  if (m.kind == 0) return (long long)n * m.img + (long long)Y * m.w + X;
  const uint32_t plane = ((Y & 1u) << 1) | (X & 1u);
  return (long long)plane * m.plane_rows + (long long)n * m.img + (long long)(Y >> 1) * m.w + (X >> 1);
}

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

template <int MMAW>
__global__ void __launch_bounds__(kThreadsBase + 32 * MMAW, (MMAW == 1 ? 4 : 2)) srgemm_kernel(const __grid_constant__ SrParams p) {
  constexpr int kThreads = kThreadsBase + 32 * MMAW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint8_t* resw = smem;                                  // resident weights (may be empty)
  uint8_t* ring = smem + p.res_bytes;                    // res_bytes is a multiple of 1024
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ring + (size_t)p.n_ring * p.slot_bytes);
  uint64_t* bar_empty = bar_full + kMaxRing;
  uint64_t* bar_tfull = bar_empty + kMaxRing;
  uint64_t* bar_tempty = bar_tfull + kMaxAcc;
  uint64_t* bar_res = bar_tempty + kMaxAcc;
  float* s_scale = reinterpret_cast<float*>(bar_res + 1);
  float* s_shift = s_scale + p.n_out;
  PvSrEntry* s_entry = reinterpret_cast<PvSrEntry*>(s_shift + p.n_out);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_entry + p.n_entries);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = p.n_out;

  // ---- one-time setup -------------------------------------------------------
  for (int i = threadIdx.x; i < p.n_entries * (int)(sizeof(PvSrEntry) / 4); i += kThreads)
    reinterpret_cast<uint32_t*>(s_entry)[i] = reinterpret_cast<const uint32_t*>(p.entries)[i];
  for (int i = threadIdx.x; i < N; i += kThreads) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  if (warp == 0 && lane == 0) {
    pv_tma_prefetch_desc(&p.a_main[0]);
    pv_tma_prefetch_desc(&p.b[0]);
    for (int i = 0; i < p.n_ring; ++i) {
      pv_mbar_init(&bar_full[i], 1);
      pv_mbar_init(&bar_empty[i], MMAW);     // every MMA warp commits its own MMAs of the slot
    }
    for (int i = 0; i < p.n_acc; ++i) {
      pv_mbar_init(&bar_tfull[i], MMAW);
      pv_mbar_init(&bar_tempty[i], 4);
    }
    pv_mbar_init(bar_res, 1);
    pv_fence_mbar_init();
  }
  if (warp == 1) pv_tmem_alloc(s_tmem, p.tmem_cols);   // warp 1 owns the TMEM allocation
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // the whole warp runs the (uniform) loop, one elected lane issues
    const bool leader = pv_elect_one();
    if (p.resident && leader) {
      pv_mbar_arrive_expect_tx(bar_res, (uint32_t)(p.res_tiles[0] * N * p.cls_width[0] * 2 +
                                                   p.res_tiles[1] * N * p.cls_width[1] * 2));
      for (int c = 0; c < 2; ++c) {
        const int rowb = p.cls_width[c] * 2;
        for (int t = 0; t < p.res_tiles[c]; ++t)
          pv_tma_load_2d(resw + p.res_cls_off[c] + (size_t)t * N * rowb, &p.b[c], bar_res, 0, t * N);
      }
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long q0 = (long long)tile * kTileM;
      for (int e = 0; e < p.n_entries; ++e) {
        const PvSrEntry& st = s_entry[e];
        const int cls = st.cls;
        const int rowb = p.cls_width[cls] * 2;
        if (st.flags & 1) pv_mbar_wait(&bar_empty[slot], phase ^ 1u, p.err, 1);
        if (leader) {
          uint8_t* base = ring + (size_t)slot * p.slot_bytes;
          if (st.flags & 1) pv_mbar_arrive_expect_tx(&bar_full[slot], st.slot_tx_bytes);
          uint8_t* a_dst = base + st.a_smem_off;
          const int32_t r0 = (int32_t)(q0 + st.a_row_off);
          pv_tma_load_2d(a_dst, &p.a_main[cls], &bar_full[slot], st.a_col, r0);
          if (st.use_tail)
            pv_tma_load_2d(a_dst + kTileM * rowb, &p.a_tail[cls], &bar_full[slot], st.a_col, r0 + kTileM);
          if (!p.resident) {
            uint8_t* b_dst = base + st.b_smem_off;
            for (int t = 0; t < st.n_taps; ++t)
              pv_tma_load_2d(b_dst + (size_t)t * N * rowb, &p.b[cls], &bar_full[slot], 0, st.b_row + t * N);
          }
        }
        __syncwarp();
        if (st.flags & 2) {
          if (++slot == p.n_ring) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp <= MMAW) {
    // ===================== MMA issuer(s) =====================
    // With MMAW > 1 the MMAs of a tile are dealt round-robin to MMAW warps, each accumulating into its
    // own TMEM accumulator (the epilogue adds them): MMAW independent single-lane issue streams.
    const int my = warp - 1;
    const uint32_t lead = pv_elect_one() ? 1u : 0u;
    // Warp-uniform loop (so descriptors live in uniform registers), one elected lane issues.
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, M=128, N
    const bool leader = pv_elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(kTileM >> 4) << 24);
    if (p.resident) {
      pv_mbar_wait(bar_res, 0, p.err, 5);
      pv_tc_fence_after();
    }
    const uint32_t res_lo = desc_lo(pv_smem_u32(resw));
    const uint32_t ring_base = pv_smem_u32(ring);
    const int n_mine = p.n_mma_w[my];
    const int tab0 = my * (kMaxMma / MMAW);
    const bool resident = p.resident != 0;
    const uint32_t slot_bytes = (uint32_t)p.slot_bytes;
    const int n_ring = p.n_ring, n_acc = p.n_acc, acc_cols = p.acc_cols;
    int slot = 0;
    uint32_t phase = 0;
    int buf = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      pv_mbar_wait(&bar_tempty[buf], aphase ^ 1u, p.err, 2);
      pv_tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * acc_cols);
      uint32_t slot_lo = 0, b_base = res_lo;
      for (int i = 0; i < n_mine; ++i) {
        const uint4 cur = p.mma[tab0 + i];   // constant-bank load
        if (cur.z & 1u) {
          pv_mbar_wait(&bar_full[slot], phase, p.err, 3);
          pv_tc_fence_after();
          slot_lo = desc_lo(ring_base + (uint32_t)slot * slot_bytes);
          b_base = resident ? res_lo : slot_lo;
        }
        // record: x/y = operand offsets (16-byte units), w = descriptor hi word, z = flags | TMEM offset << 16
        const uint32_t alo = slot_lo + cur.x;
        const uint32_t blo = b_base + cur.y;
        pv_umma_bf16_pred(tmem_d + (cur.z >> 16), ((uint64_t)cur.w << 32) | alo, ((uint64_t)cur.w << 32) | blo, idesc,
                          cur.z & 8u, lead);
        if (cur.z & 2u) {
          pv_umma_commit_pred(&bar_empty[slot], lead);  // frees the smem slot once these MMAs retire
          if (++slot == n_ring) { slot = 0; phase ^= 1u; }
        }
      }
      pv_umma_commit_pred(&bar_tfull[buf], lead);  // accumulator complete
      if (++buf == n_acc) { buf = 0; aphase ^= 1u; }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
    const int m = quarter * 32 + lane;
    const uint32_t img = (uint32_t)p.hq * (uint32_t)p.wq;
    int buf = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t q = (uint32_t)tile * kTileM + (uint32_t)m;   // q_rows < 2^31
      bool valid = (long long)q < p.q_rows;
      const uint32_t n = q / img;
      const uint32_t rem = q - n * img;
      const uint32_t y = rem / (uint32_t)p.wq;
      const uint32_t x = rem - y * (uint32_t)p.wq;
      valid = valid && (y < (uint32_t)p.oh) && (x < (uint32_t)p.ow);
      const long long drow = row_of(p.dst, n, y, x);
      const long long rrow = p.has_resid ? row_of(p.res, n, y, x) : 0;

      pv_mbar_wait(&bar_tfull[buf], aphase, p.err, 4);
      pv_tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.acc_cols);

      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[2][16];
        const bool two = (c0 + 16) < N;
        pv_tmem_ld16(taddr + c0, v[0]);
        if (two) pv_tmem_ld16(taddr + c0 + 16, v[1]);
        pv_tmem_ld_wait();
        for (int sp = 1; sp < p.acc_split; ++sp) {   // add the other partial accumulators of this tile
          uint32_t w[2][16];
          pv_tmem_ld16(taddr + sp * N + c0, w[0]);
          if (two) pv_tmem_ld16(taddr + sp * N + c0 + 16, w[1]);
          pv_tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[0][j] = __float_as_uint(__uint_as_float(v[0][j]) + __uint_as_float(w[0][j]));
            if (two) v[1][j] = __float_as_uint(__uint_as_float(v[1][j]) + __uint_as_float(w[1][j]));
          }
        }
        if (valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int cb = c0 + 16 * h;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaf(__uint_as_float(v[h][j]), s_scale[cb + j], s_shift[cb + j]);
            if (p.has_resid) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.resid + rrow * p.res.cols + cb);
              uint4 r0 = rp[0], r1 = rp[1];
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
                f[2 * j] += __bfloat162float(hh.x);
                f[2 * j + 1] += __bfloat162float(hh.y);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.out_mode == 0) {
              uint4 o0, o1;
              o0.x = pv_pack_bf16x2(f[0], f[1]);
              o0.y = pv_pack_bf16x2(f[2], f[3]);
              o0.z = pv_pack_bf16x2(f[4], f[5]);
              o0.w = pv_pack_bf16x2(f[6], f[7]);
              o1.x = pv_pack_bf16x2(f[8], f[9]);
              o1.y = pv_pack_bf16x2(f[10], f[11]);
              o1.z = pv_pack_bf16x2(f[12], f[13]);
              o1.w = pv_pack_bf16x2(f[14], f[15]);
              uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + drow * p.dst.cols + cb);
              dp[0] = o0;
              dp[1] = o1;
            } else if (p.out_mode == 3) {
              float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + drow * p.dst.cols + cb);
              dp[0] = make_float4(f[0], f[1], f[2], f[3]);
              dp[1] = make_float4(f[4], f[5], f[6], f[7]);
              dp[2] = make_float4(f[8], f[9], f[10], f[11]);
              dp[3] = make_float4(f[12], f[13], f[14], f[15]);
            } else if (cb == 0) {
              reinterpret_cast<float*>(p.out)[drow] = f[0];
            }
          }
        }
      }
      pv_tc_fence_before();
      __syncwarp();
      if (lane == 0) pv_mbar_arrive(&bar_tempty[buf]);
      if (++buf == p.n_acc) { buf = 0; aphase ^= 1u; }
    }
  }

  // ---- teardown ----------------------------------------------------------------
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D bf16 map: `cols` elements per row, rows `row_stride_bytes` apart (may be smaller than a row:
// overlapping rows are how a first conv can read pixel runs straight out of an NHWC plane).
int encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
              uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    pv_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return PV_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols == 64   ? CU_TENSOR_MAP_SWIZZLE_128B
                          : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                           : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pv_set_error("cuTensorMapEncodeTiled failed: CUresult %d (cols=%llu rows=%llu stride=%llu box=%ux%u)", (int)r,
                 (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_cols,
                 box_rows);
    return PV_ERR_CUDA;
  }
  return PV_OK;
}

struct SrPlan {
  SrParams p;
  PvSrEntry* d_entries = nullptr;
  int* d_err = nullptr;
  size_t smem_bytes = 0;
  int num_sms = 0;
  int max_ctas = 0;
  int mmaw = 1;
};

inline int align1k(int v) { return (v + 1023) & ~1023; }

}  // namespace

extern "C" int pv_srgemm_create(const PvSrgemmDesc* d, void** out_handle) {
  PV_REQUIRE(d && out_handle, "pv_srgemm_create: null argument");
  PV_REQUIRE(d->n_out >= 16 && d->n_out <= 256 && d->n_out % 16 == 0, "srgemm: n_out=%d must be a multiple of 16 in [16,256]", d->n_out);
  PV_REQUIRE(d->n_classes == 1 || d->n_classes == 2, "srgemm: n_classes=%d", d->n_classes);
  PV_REQUIRE(d->n_entries >= 1 && d->n_entries <= PV_SR_MAX_ENTRIES, "srgemm: n_entries=%d out of range", d->n_entries);
  PV_REQUIRE(d->tail_rows >= 0 && d->tail_rows <= 128 && d->tail_rows % 8 == 0, "srgemm: tail_rows=%d", d->tail_rows);
  PV_REQUIRE(d->x_cols % 8 == 0 && d->x_cols >= 16, "srgemm: x_cols=%d must be a multiple of 8", d->x_cols);
  PV_REQUIRE(d->x_rows > 0 && d->x_rows < (1ll << 31), "srgemm: x_rows=%lld out of range", (long long)d->x_rows);
  PV_REQUIRE(d->hq > 0 && d->wq > 0 && d->oh > 0 && d->ow > 0, "srgemm: bad grid");
  PV_REQUIRE(d->out_mode == 0 || d->out_mode == 2 || d->out_mode == 3, "srgemm: out_mode=%d", d->out_mode);
  PV_REQUIRE(d->x && d->out && d->scale && d->shift && d->entries, "srgemm: null operand");
  const long long row_stride = d->x_row_stride_bytes > 0 ? d->x_row_stride_bytes : (long long)d->x_cols * 2;
  PV_REQUIRE(row_stride % 16 == 0, "srgemm: x_row_stride_bytes=%lld must be a multiple of 16", row_stride);
  for (int c = 0; c < d->n_classes; ++c) {
    const int w = d->class_width[c];
    PV_REQUIRE(w == 16 || w == 32 || w == 64, "srgemm: class_width[%d]=%d", c, w);
    PV_REQUIRE(d->w_packed[c] && d->w_rows[c] > 0 && d->w_rows[c] % d->n_out == 0, "srgemm: bad weights for class %d", c);
  }
  if (d->out_mode == 0 || d->out_mode == 3) {
    PV_REQUIRE(d->dst.cols % 8 == 0 && d->dst.cols >= d->n_out, "srgemm: dst.cols=%d < n_out=%d", d->dst.cols, d->n_out);
  }
  if (d->resid) PV_REQUIRE(d->res.cols % 8 == 0 && d->res.cols >= d->n_out, "srgemm: res.cols=%d", d->res.cols);

  // ---- resident weights? ----
  int res_bytes = 0, res_cls_off[2] = {0, 0}, res_tiles[2] = {0, 0};
  long long wbytes = 0;
  for (int c = 0; c < d->n_classes; ++c) wbytes += d->w_rows[c] * d->class_width[c] * 2;
  const bool resident = d->weights_resident_max_bytes > 0 && wbytes <= d->weights_resident_max_bytes;
  if (resident) {
    int off = 0;
    for (int c = 0; c < d->n_classes; ++c) {
      res_cls_off[c] = off;
      res_tiles[c] = (int)(d->w_rows[c] / d->n_out);
      off = align1k(off + (int)(d->w_rows[c] * d->class_width[c] * 2));
    }
    res_bytes = off;
  }

  // ---- validate entries, lay out slots ----
  std::vector<PvSrEntry> ent(d->entries, d->entries + d->n_entries);
  int slot_bytes = 0;
  {
    int cur = 0;
    uint32_t tx = 0;
    int first = -1;
    for (int e = 0; e < d->n_entries; ++e) {
      PvSrEntry& st = ent[e];
      PV_REQUIRE(st.cls >= 0 && st.cls < d->n_classes, "srgemm: entry %d class %d", e, st.cls);
      PV_REQUIRE(st.n_taps >= 1 && st.n_taps <= PV_SR_MAX_TAPS, "srgemm: entry %d n_taps %d", e, st.n_taps);
      const int w = d->class_width[st.cls];
      PV_REQUIRE(st.a_col >= 0 && st.a_col + w <= d->x_cols, "srgemm: entry %d column segment out of range", e);
      for (int t = 0; t < st.n_taps; ++t) {
        const int rel = st.tap_rel[t];
        PV_REQUIRE(rel >= 0 && rel <= (st.use_tail ? d->tail_rows : 0), "srgemm: entry %d tap %d rel %d outside slab", e, t, rel);
      }
      PV_REQUIRE(st.b_row >= 0 && st.b_row % d->n_out == 0 &&
                     (long long)st.b_row + (long long)st.n_taps * d->n_out <= d->w_rows[st.cls],
                 "srgemm: entry %d weights out of range", e);
      if (e == 0) PV_REQUIRE(st.flags & 1, "srgemm: first entry must open a slot");
      if (st.flags & 1) { cur = 0; tx = 0; first = e; }
      const int slab_rows = kTileM + (st.use_tail ? d->tail_rows : 0);
      st.a_smem_off = cur;
      cur = align1k(cur + slab_rows * w * 2);
      tx += (uint32_t)(slab_rows * w * 2);
      if (!resident) {
        st.b_smem_off = cur;
        cur = align1k(cur + st.n_taps * d->n_out * w * 2);
        tx += (uint32_t)(st.n_taps * d->n_out * w * 2);
      } else {
        st.b_smem_off = 0;
      }
      ent[first].slot_tx_bytes = tx;
      if (cur > slot_bytes) slot_bytes = cur;
      if (e == d->n_entries - 1) PV_REQUIRE(st.flags & 2, "srgemm: last entry must close its slot");
      if ((st.flags & 2) && e + 1 < d->n_entries) PV_REQUIRE(ent[e + 1].flags & 1, "srgemm: entry %d must open a slot", e + 1);
      if (!(st.flags & 2)) PV_REQUIRE(e + 1 < d->n_entries && !(ent[e + 1].flags & 1), "srgemm: slot flags inconsistent at entry %d", e);
    }
  }
  int mma_per_tile = 0;
  for (int k = 0; k < d->n_entries; ++k) mma_per_tile += ent[k].n_taps * (d->class_width[ent[k].cls] / 16);
  PV_REQUIRE(mma_per_tile <= kMaxMma, "srgemm: %d MMAs per tile exceed %d", mma_per_tile, kMaxMma);
  const int cps = d->ctas_per_sm > 1 ? d->ctas_per_sm : 1;
  PV_REQUIRE(cps == 1 || cps == 2 || cps == 4, "srgemm: ctas_per_sm=%d (1, 2 or 4)", cps);
  const size_t fixed = sizeof(PvSrEntry) * d->n_entries + 2 * d->n_out * sizeof(float) +
                       (2 * kMaxRing + 2 * kMaxAcc + 1) * sizeof(uint64_t) + 64;
  // each resident CTA costs 1 KB of reserved shared memory on top of its dynamic allocation
  const long long budget = (227 * 1024) / cps - 1024 - 1024 - (long long)fixed - res_bytes;
  int n_ring = (int)(budget / slot_bytes);
  if (n_ring > kMaxRing) n_ring = kMaxRing;
  PV_REQUIRE(n_ring >= 2, "srgemm: slot of %d bytes (+%d resident) does not fit a 2-deep ring", slot_bytes, res_bytes);

  SrPlan* plan = new SrPlan();
  memset(&plan->p, 0, sizeof(SrParams));
  SrParams& p = plan->p;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    pv_set_error("pv_srgemm_create: no CUDA device: %s", cudaGetErrorString(e));
    delete plan;
    return PV_ERR_CUDA;
  }
  plan->max_ctas = (d->max_ctas > 0 ? d->max_ctas : plan->num_sms) * (d->ctas_per_sm > 1 ? d->ctas_per_sm : 1);

  int rc = PV_OK;
  for (int c = 0; c < d->n_classes && rc == PV_OK; ++c) {
    const int w = d->class_width[c];
    rc = encode_2d(&p.a_main[c], d->x, d->x_cols, d->x_rows, row_stride, w, kTileM);
    if (rc == PV_OK && d->tail_rows > 0) rc = encode_2d(&p.a_tail[c], d->x, d->x_cols, d->x_rows, row_stride, w, d->tail_rows);
    if (rc == PV_OK) rc = encode_2d(&p.b[c], d->w_packed[c], w, d->w_rows[c], (uint64_t)w * 2, w, d->n_out);
  }
  if (rc != PV_OK) {
    delete plan;
    return rc;
  }
  if (cudaMalloc(&plan->d_entries, sizeof(PvSrEntry) * d->n_entries) != cudaSuccess ||
      cudaMalloc(&plan->d_err, sizeof(int)) != cudaSuccess) {
    pv_set_error("pv_srgemm_create: cudaMalloc failed");
    delete plan;
    return PV_ERR_CUDA;
  }
  cudaMemcpy(plan->d_entries, ent.data(), sizeof(PvSrEntry) * d->n_entries, cudaMemcpyHostToDevice);
  cudaMemset(plan->d_err, 0, sizeof(int));

  p.entries = plan->d_entries;
  p.scale = d->scale;
  p.shift = d->shift;
  p.out = d->out;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(d->resid);
  p.dst = d->dst;
  p.res = d->res;
  p.n_out = d->n_out;
  p.n_entries = d->n_entries;
  p.n_ring = n_ring;
  p.slot_bytes = slot_bytes;
  p.tail_rows = d->tail_rows;
  p.cls_width[0] = d->class_width[0];
  p.cls_width[1] = d->n_classes > 1 ? d->class_width[1] : d->class_width[0];
  p.resident = resident ? 1 : 0;
  p.res_bytes = res_bytes;
  p.res_cls_off[0] = res_cls_off[0];
  p.res_cls_off[1] = res_cls_off[1];
  p.res_tiles[0] = res_tiles[0];
  p.res_tiles[1] = res_tiles[1];
  p.hq = d->hq;
  p.wq = d->wq;
  p.oh = d->oh;
  p.ow = d->ow;
  p.relu = d->relu;
  p.out_mode = d->out_mode;
  p.has_resid = d->resid != nullptr;
  const int mmaw = (d->mma_warps == 2 || d->mma_warps == 4) ? d->mma_warps : 1;
  int split = mmaw > 1 ? mmaw : d->acc_split;
  if (split <= 0) split = 1;   // measured (gpurun #5): rotating accumulators does not help — the issue loop, not the
                               // accumulate dependency, bounds small-N layers; kept as an explicit option
  const int tmem_cap = 512 / cps;
  if (mmaw == 1)
    while (split > 1 && (split > mma_per_tile || split * d->n_out * 2 > tmem_cap)) split >>= 1;
  PV_REQUIRE(mmaw == 1 || (mma_per_tile >= mmaw && cps <= 2), "srgemm: mma_warps=%d needs >= that many MMAs per tile and ctas_per_sm <= 2", mmaw);
  const int acc_cols = split * d->n_out;
  PV_REQUIRE(acc_cols <= tmem_cap, "srgemm: n_out=%d does not fit %d TMEM columns (ctas_per_sm=%d)", d->n_out, tmem_cap, cps);
  int n_acc = tmem_cap / acc_cols;
  if (n_acc > kMaxAcc) n_acc = kMaxAcc;
  uint32_t cols = 32;
  while (cols < (uint32_t)(n_acc * acc_cols)) cols <<= 1;
  plan->mmaw = mmaw;
  p.n_acc = n_acc;
  p.acc_split = split;
  p.acc_cols = acc_cols;
  // ---- expand entries into per-MMA records, dealt round-robin to the MMA warps ----
  {
    std::vector<MmaRec> lists[4];
    std::vector<int> slot_of[4];
    uint32_t idx = 0;
    int slot_id = -1;
    for (int k = 0; k < d->n_entries; ++k) {
      const PvSrEntry& st = ent[k];
      if (st.flags & 1) ++slot_id;
      const int w = d->class_width[st.cls];
      const int rowb = w * 2;
      const long long b0 = resident ? (long long)res_cls_off[st.cls] + (long long)st.b_row * rowb : (long long)st.b_smem_off;
      for (int t = 0; t < st.n_taps; ++t)
        for (int ks = 0; ks < w / 16; ++ks) {
          MmaRec r;
          r.a_rel16 = (uint32_t)((st.a_smem_off + st.tap_rel[t] * rowb + ks * 32) >> 4);
          r.b_rel16 = (uint32_t)((b0 + (long long)t * d->n_out * rowb + ks * 32) >> 4);
          r.flags = ((idx >= (uint32_t)split ? 1u : 0u) << 3) | (((idx % (uint32_t)split) * (uint32_t)d->n_out) << 16);
          r.pad = desc_hi(w);
          const int wp = mmaw > 1 ? (int)(idx % (uint32_t)mmaw) : 0;
          lists[wp].push_back(r);
          slot_of[wp].push_back(slot_id);
          ++idx;
        }
    }
    const int n_slots = slot_id + 1;
    const int cap = kMaxMma / mmaw;
    bool ok = true;
    for (int wp = 0; wp < mmaw && ok; ++wp) {
      if ((int)lists[wp].size() > cap) ok = false;
      // bit0 on the warp's first record of every slot (wait for the data), bit1 on its last (release the slot)
      int seen = 0;
      for (size_t k = 0; k < lists[wp].size(); ++k) {
        if (k == 0 || slot_of[wp][k] != slot_of[wp][k - 1]) { lists[wp][k].flags |= 1u; ++seen; }
        if (k + 1 == lists[wp].size() || slot_of[wp][k + 1] != slot_of[wp][k]) lists[wp][k].flags |= 2u;
      }
      if (seen != n_slots) ok = false;   // every warp must touch every slot (it owes the slot one commit)
      p.n_mma_w[wp] = (int)lists[wp].size();
      for (size_t k = 0; k < lists[wp].size(); ++k)
        p.mma[wp * cap + k] = make_uint4(lists[wp][k].a_rel16, lists[wp][k].b_rel16, lists[wp][k].flags, lists[wp][k].pad);
    }
    if (!ok) {
      pv_set_error("srgemm: MMA table does not fit (mma_warps=%d, %d MMAs/tile, %d slots)", mmaw, mma_per_tile, n_slots);
      cudaFree(plan->d_entries);
      cudaFree(plan->d_err);
      delete plan;
      return PV_ERR_INVALID;
    }
  }
  p.n_mma = mma_per_tile;

  p.tmem_cols = cols;
  p.err = plan->d_err;
  plan->smem_bytes = (size_t)res_bytes + (size_t)n_ring * slot_bytes + fixed + 1024;

  e = cudaFuncSetAttribute(srgemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);  // function-wide: always the maximum
  if (e == cudaSuccess) e = cudaFuncSetAttribute(srgemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(srgemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    pv_set_error("pv_srgemm_create: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    cudaFree(plan->d_entries);
    cudaFree(plan->d_err);
    delete plan;
    return PV_ERR_CUDA;
  }
  *out_handle = plan;
  return PV_OK;
}

extern "C" int pv_srgemm_run(void* handle, int64_t q_rows, void* stream) {
  PV_REQUIRE(handle, "pv_srgemm_run: null handle");
  SrPlan* plan = static_cast<SrPlan*>(handle);
  PV_REQUIRE(q_rows > 0 && q_rows < (1ll << 31), "pv_srgemm_run: q_rows=%lld", (long long)q_rows);
  SrParams p = plan->p;
  p.q_rows = q_rows;
  p.num_tiles = (int)((q_rows + kTileM - 1) / kTileM);
  int grid = p.num_tiles < plan->max_ctas ? p.num_tiles : plan->max_ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plan->mmaw == 4)
    srgemm_kernel<4><<<grid, kThreadsBase + 128, plan->smem_bytes, st>>>(p);
  else if (plan->mmaw == 2)
    srgemm_kernel<2><<<grid, kThreadsBase + 64, plan->smem_bytes, st>>>(p);
  else
    srgemm_kernel<1><<<grid, kThreadsBase + 32, plan->smem_bytes, st>>>(p);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_srgemm_info(void* handle, int* n_ring, int* slot_bytes, int* resident, int* n_acc, int* acc_split) {
  PV_REQUIRE(handle, "pv_srgemm_info: null handle");
  SrPlan* plan = static_cast<SrPlan*>(handle);
  if (n_ring) *n_ring = plan->p.n_ring;
  if (slot_bytes) *slot_bytes = plan->p.slot_bytes;
  if (resident) *resident = plan->p.resident;
  if (n_acc) *n_acc = plan->p.n_acc;
  if (acc_split) *acc_split = plan->p.acc_split;
  return PV_OK;
}

extern "C" int pv_srgemm_check(void* handle, void* stream) {
  PV_REQUIRE(handle, "pv_srgemm_check: null handle");
  SrPlan* plan = static_cast<SrPlan*>(handle);
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&flag, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    pv_set_error("pv_srgemm_check: %s", cudaGetErrorString(e));
    return PV_ERR_CUDA;
  }
  if (flag != 0) {
    pv_set_error("srgemm: device-side pipeline timeout (role code %d)", flag);
    cudaMemset(plan->d_err, 0, sizeof(int));
    return PV_ERR_DEVICE_TIMEOUT;
  }
  return PV_OK;
}

extern "C" int pv_srgemm_destroy(void* handle) {
  if (!handle) return PV_OK;
  SrPlan* plan = static_cast<SrPlan*>(handle);
  cudaFree(plan->d_entries);
  cudaFree(plan->d_err);
  delete plan;
  return PV_OK;
}
