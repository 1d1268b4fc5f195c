// detconv.cu — the detector's 5x5 / 9x1 convolutions as 2-D tiled implicit GEMMs on tcgen05 (sm_100a).
//
// Replaces the `con` layers 2..7 of dlib's CNN/MMOD face detector that the reference reaches through
// face_detector_(rgb, 1), pyannote/video/face/face.py:66.  Activations are plain NHWC bf16 tensors
// [B, H, Ws, C] (Ws = row pitch in pixels, even); nothing is padded in HBM: TMA's out-of-bounds zero
// fill IS the convolution's zero padding.
//
// One output tile = 8 x 16 output pixels (x, y) = the 128 rows of one TMEM accumulator, row m <-> pixel
// (tx = m & 7, ty = m >> 3).  Its input patch ((8+KW-1) x (16+KH-1) pixels for stride 1) is fetched ONCE
// by a 4-D TMA box into swizzled shared memory, one patch pixel per operand row.  Every filter tap
// (kh, kw) then reads its A operand straight out of that patch: the 8-row core-matrix groups of the
// UMMA descriptor are the 8 output pixels of one output row, so
//     start address = patch + ((kh * PW + kw) * row_bytes),   SBO (group stride) = PW * row_bytes
// walks the tile's 16 output rows — an im2col that exists only in the descriptor.  (The swizzle XOR is a
// function of absolute shared-memory address bits for both TMA and the tensor core, so descriptors may
// start at any row; measured on B200)  For stride 2 an operand row is a PAIR of pixels
// (2C channels): neighbouring outputs are then exactly one row apart, the tap's column parity selects
// the K offset inside the row, and SBO skips two patch rows.
//
// Compared with the 1-D "shifted row" srgemm this reads (8+4)(16+4)/128 = 1.9 input pixels per output
// instead of 5.3 and its MMA issue loop is unrolled at compile time (two adds per tcgen05.mma).
//
// Warps: 0 = TMA producer, 1 = MMA issuer (+ TMEM owner), 2..5 = epilogue (tcgen05.ld -> affine -> ReLU
// -> bf16 -> NHWC global).  Weights stay resident in shared memory for the whole launch.
#include <cuda.h>
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kTW = 8, kTH = 16;
constexpr int kMaxStages = 6;
constexpr int kAcc = 4;
constexpr int kThreads = 192;

struct DcParams {
  CUtensorMap in[2];
  const uint8_t* w_img;      // shared-memory image of the packed weights
  uint32_t w_bytes;
  const float* scale;
  const float* shift;
  void* out;
  int B, OH, OW;
  int out_pitch;             // pixels per output row in memory
  int out_cs;                // channels per output pixel in memory
  int tiles_x, tiles_y, num_tiles;
  int relu;
  int n_stages;
  int* err;
};

template <int C, int S>
struct Geo {
  // operand row: one pixel (stride 1) or one pixel pair (stride 2)
  static constexpr int kRowEl0 = S == 2 ? 2 * C : (C >= 32 ? 32 : 16);
  static constexpr int kRowEl1 = S == 2 ? 0 : C - kRowEl0;
  static constexpr int kSegs = kRowEl1 > 0 ? 2 : 1;
};

__host__ __device__ constexpr uint32_t layout_of_rowb(int rowb) { return rowb == 128 ? 2u : (rowb == 64 ? 4u : 6u); }

__host__ __device__ constexpr int align1k(int v) { return (v + 1023) & ~1023; }

// hi word of a K-major swizzled A descriptor: SBO, version 1, swizzle mode
__host__ __device__ constexpr uint32_t a_desc_hi(int sbo_bytes, int rowb) {
  return (uint32_t)((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_of_rowb(rowb) << 29);
}
// hi word of the un-swizzled K-major B descriptor: SBO = 128 (8 rows of 16 bytes)
__host__ __device__ constexpr uint32_t b_desc_hi() { return (uint32_t)(128 >> 4) | (1u << 14); }

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(pv_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(pv_smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   pv_smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(pv_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
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

// C: input channels per pixel in memory (16/32/48), N: output channels (multiple of 16, <= 64),
// KH x KW filter, S stride (1: pad = K/2 on both axes; 2: pad 0), F32: fp32 output rows (last layer)
template <int C, int N, int KH, int KW, int S, bool F32>
__global__ void __launch_bounds__(kThreads, 1) detconv_kernel(const __grid_constant__ DcParams p) {
  using G = Geo<C, S>;
  constexpr int PW = S == 2 ? ((kTW - 1) * 2 + KW + 1) / 2 : kTW + KW - 1;   // operand rows per patch row
  constexpr int PH = (kTH - 1) * S + KH;
  constexpr int ROWS = PW * PH;
  constexpr int ROWB0 = G::kRowEl0 * 2, ROWB1 = G::kRowEl1 * 2;
  constexpr int SEG0_BYTES = align1k(ROWS * ROWB0);
  constexpr int SEG1_BYTES = G::kSegs > 1 ? align1k(ROWS * ROWB1) : 0;
  constexpr int STAGE_BYTES = SEG0_BYTES + SEG1_BYTES;
  constexpr uint32_t TX_BYTES = (uint32_t)(ROWS * (ROWB0 + ROWB1));
  constexpr int KCH = C / 16;                 // 16-channel K chunks per tap
  constexpr int BTILE = N * 32;               // bytes of one N x 16 weight tile
  constexpr int PADY = S == 1 ? KH / 2 : 0, PADX = S == 1 ? KW / 2 : 0;
  constexpr uint32_t TMEM_COLS = (kAcc * N <= 64) ? 64u : ((kAcc * N <= 128) ? 128u : 256u);
  static_assert(kAcc * N <= 256, "accumulator ring does not fit");
  static_assert(N % 16 == 0 && N >= 16 && N <= 64, "N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* ring = smem + align1k((int)p.w_bytes);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ring + (size_t)p.n_stages * STAGE_BYTES);
  uint64_t* bar_empty = bar_full + kMaxStages;
  uint64_t* bar_tfull = bar_empty + kMaxStages;
  uint64_t* bar_tempty = bar_tfull + kAcc;
  uint64_t* bar_w = bar_tempty + kAcc;
  float* s_scale = reinterpret_cast<float*>(bar_w + 1);
  float* s_shift = s_scale + N;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_shift + N);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < N; i += kThreads) {
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
    for (int i = 0; i < kAcc; ++i) {
      pv_mbar_init(&bar_tfull[i], 1);
      pv_mbar_init(&bar_tempty[i], 4);
    }
    pv_mbar_init(bar_w, 1);
    pv_fence_mbar_init();
  }
  if (warp == 1) pv_tmem_alloc(s_tmem, TMEM_COLS);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (pv_elect_one()) {
      pv_mbar_arrive_expect_tx(bar_w, p.w_bytes);
      for (uint32_t off = 0; off < p.w_bytes; off += 32768u) {
        const uint32_t n = p.w_bytes - off < 32768u ? p.w_bytes - off : 32768u;
        bulk_load_1d(wsm + off, p.w_img + off, n, bar_w);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        const int tyi = r / p.tiles_x;
        const int txi = r - tyi * p.tiles_x;
        const int ix0 = S == 2 ? txi * kTW : txi * kTW - PADX;        // in operand rows (pixels or pixel pairs)
        const int iy0 = tyi * kTH * S - PADY;
        pv_mbar_wait(&bar_empty[stage], phase ^ 1u, p.err, 1);
        uint8_t* dst = ring + (size_t)stage * STAGE_BYTES;
        pv_mbar_arrive_expect_tx(&bar_full[stage], TX_BYTES);
        tma_load_4d(dst, &p.in[0], &bar_full[stage], 0, ix0, iy0, b);
        if (G::kSegs > 1) tma_load_4d(dst + SEG0_BYTES, &p.in[1], &bar_full[stage], G::kRowEl0, ix0, iy0, b);
        if (++stage == p.n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (pv_elect_one()) {
      // kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr int SBO0 = (S == 2 ? 2 : 1) * PW * ROWB0;
      constexpr int SBO1 = PW * ROWB1;
      constexpr uint32_t AHI0 = a_desc_hi(SBO0, ROWB0);
      constexpr uint32_t AHI1 = a_desc_hi(SBO1, ROWB1 > 0 ? ROWB1 : 32);
      constexpr uint32_t BHI = b_desc_hi();
      constexpr uint32_t B_LBO = (uint32_t)((N * 16) >> 4) << 16;   // K-direction core-matrix stride, in the lo word
      pv_mbar_wait(bar_w, 0, p.err, 5);
      pv_tc_fence_after();
      const uint32_t w_lo = ((pv_smem_u32(wsm) & 0x3FFFFu) >> 4) | B_LBO;
      const uint32_t ring_lo = (pv_smem_u32(ring) & 0x3FFFFu) >> 4;
      int stage = 0, buf = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        pv_mbar_wait(&bar_tempty[buf], aphase ^ 1u, p.err, 2);
        pv_mbar_wait(&bar_full[stage], phase, p.err, 3);
        pv_tc_fence_after();
        const uint32_t a0 = (ring_lo + (uint32_t)stage * (uint32_t)(STAGE_BYTES >> 4)) | (1u << 16);
        const uint32_t a1 = a0 + (uint32_t)(SEG0_BYTES >> 4);
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * N);
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
#pragma unroll
          for (int kw = 0; kw < KW; ++kw) {
            const int tap = kh * KW + kw;
#pragma unroll
            for (int kc = 0; kc < KCH; ++kc) {
              uint32_t a_lo, a_hi;
              if (S == 2) {
                const int off = (kh * PW + (kw >> 1)) * ROWB0 + (kw & 1) * C * 2 + kc * 32;
                a_lo = a0 + (uint32_t)(off >> 4);
                a_hi = AHI0;
              } else if (kc * 16 < G::kRowEl0) {
                const int off = (kh * PW + kw) * ROWB0 + kc * 32;
                a_lo = a0 + (uint32_t)(off >> 4);
                a_hi = AHI0;
              } else {
                const int off = (kh * PW + kw) * ROWB1 + (kc * 16 - G::kRowEl0) * 2;
                a_lo = a1 + (uint32_t)(off >> 4);
                a_hi = AHI1;
              }
              const uint32_t b_lo = w_lo + (uint32_t)(((tap * KCH + kc) * BTILE) >> 4);
              umma_bf16(tmem_d, a_lo, a_hi, b_lo, BHI, idesc, (kh | kw | kc) != 0 ? 1u : 0u);
            }
          }
        }
        pv_umma_commit(&bar_empty[stage]);
        pv_umma_commit(&bar_tfull[buf]);
        if (++stage == p.n_stages) { stage = 0; phase ^= 1u; }
        if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int tx = m & 7, ty = m >> 3;
    int buf = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      const int tyi = r / p.tiles_x;
      const int txi = r - tyi * p.tiles_x;
      const int oy = tyi * kTH + ty, ox = txi * kTW + tx;
      const bool valid = oy < p.OH && ox < p.OW;
      const long long pix = ((long long)b * p.OH + oy) * p.out_pitch + ox;
      pv_mbar_wait(&bar_tfull[buf], aphase, p.err, 4);
      pv_tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * N);
      uint32_t v[N / 16][16];
#pragma unroll
      for (int j = 0; j < N / 16; ++j) pv_tmem_ld16(taddr + j * 16, v[j]);
      pv_tmem_ld_wait();
      pv_tc_fence_before();
      __syncwarp();
      if (lane == 0) pv_mbar_arrive(&bar_tempty[buf]);     // accumulator is in registers: release it early
      if (valid) {
#pragma unroll
        for (int j = 0; j < N / 16; ++j) {
          float f[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            f[k] = fmaf(__uint_as_float(v[j][k]), s_scale[j * 16 + k], s_shift[j * 16 + k]);
            if (p.relu) f[k] = fmaxf(f[k], 0.f);
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
      if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

struct DcPlan {
  DcParams p;
  int kind = -1;
  size_t smem_bytes = 0;
  int num_sms = 0;
  int Bmax = 0;
  int* d_err = nullptr;
};

struct Instance {
  int C, N, KH, KW, S, f32;
};
constexpr Instance kInstances[] = {
    {16, 32, 5, 5, 2, 0},   // conv2
    {32, 32, 5, 5, 2, 0},   // conv3
    {32, 48, 5, 5, 1, 0},   // conv4
    {48, 48, 5, 5, 1, 0},   // conv5, conv6
    {48, 16, 9, 1, 1, 1},   // conv7 as 9x1 (filter columns as output channels), fp32 rows
};
constexpr int kNumInstances = sizeof(kInstances) / sizeof(kInstances[0]);

template <int C, int N, int KH, int KW, int S, bool F32>
cudaError_t launch(const DcPlan* plan, const DcParams& p, int grid, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (pv_attr_needed(&attr)) {
    cudaError_t e = cudaFuncSetAttribute(detconv_kernel<C, N, KH, KW, S, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return e;
  }
  detconv_kernel<C, N, KH, KW, S, F32><<<grid, kThreads, plan->smem_bytes, st>>>(p);
  return cudaGetLastError();
}

void patch_geometry(const Instance& in, int* pw, int* ph, int* rowel0, int* rowel1) {
  *pw = in.S == 2 ? ((kTW - 1) * 2 + in.KW + 1) / 2 : kTW + in.KW - 1;
  *ph = (kTH - 1) * in.S + in.KH;
  *rowel0 = in.S == 2 ? 2 * in.C : (in.C >= 32 ? 32 : 16);
  *rowel1 = in.S == 2 ? 0 : in.C - *rowel0;
}

}  // namespace

extern "C" int pv_detconv_create(const PvDetconvDesc* d, void** out_handle) {
  PV_REQUIRE(d && out_handle, "pv_detconv_create: null argument");
  PV_REQUIRE(d->x && d->w_img && d->scale && d->shift && d->out, "pv_detconv_create: null operand");
  int kind = -1;
  for (int i = 0; i < kNumInstances; ++i) {
    const Instance& in = kInstances[i];
    if (in.C == d->c_in && in.N == d->n_out && in.KH == d->kh && in.KW == d->kw && in.S == d->stride && in.f32 == d->out_f32)
      kind = i;
  }
  PV_REQUIRE(kind >= 0, "pv_detconv_create: no kernel instance for C=%d N=%d %dx%d stride %d f32=%d", d->c_in, d->n_out,
             d->kh, d->kw, d->stride, d->out_f32);
  const Instance& in = kInstances[kind];
  PV_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->pitch >= d->W && d->pitch % 2 == 0, "pv_detconv_create: bad input extent");
  const int pad_y = in.S == 1 ? in.KH / 2 : 0, pad_x = in.S == 1 ? in.KW / 2 : 0;
  const int OH = (d->H + 2 * pad_y - in.KH) / in.S + 1, OW = (d->W + 2 * pad_x - in.KW) / in.S + 1;
  PV_REQUIRE(OH > 0 && OW > 0, "pv_detconv_create: empty output");
  PV_REQUIRE(d->out_pitch >= OW && d->out_cs >= in.N && d->out_cs % 16 == 0, "pv_detconv_create: output pitch %d / channel stride %d (rows are written with 32-byte stores)",
             d->out_pitch, d->out_cs);
  PV_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w_img) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->out) & 31) == 0, "pv_detconv_create: operands must be 16-byte (output: 32-byte) aligned");
  const int kch = in.C / 16;
  const uint32_t w_bytes = (uint32_t)(in.KH * in.KW * kch * in.N * 32);
  PV_REQUIRE(d->w_bytes == (int64_t)w_bytes, "pv_detconv_create: weight image is %lld bytes, expected %u", (long long)d->w_bytes, w_bytes);

  int pw, ph, rowel0, rowel1;
  patch_geometry(in, &pw, &ph, &rowel0, &rowel1);
  const int rows = pw * ph;
  const int stage_bytes = align1k(rows * rowel0 * 2) + (rowel1 > 0 ? align1k(rows * rowel1 * 2) : 0);
  const size_t fixed = (2 * kMaxStages + 2 * kAcc + 1) * sizeof(uint64_t) + 2 * in.N * sizeof(float) + 64;
  const long long budget = 227 * 1024 - 1024 - 1024 - (long long)fixed - align1k((int)w_bytes);
  int n_stages = (int)(budget / stage_bytes);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  PV_REQUIRE(n_stages >= 2, "pv_detconv_create: patch of %d bytes (+%u weights) does not fit a 2-deep ring", stage_bytes, w_bytes);

  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    pv_set_error("pv_detconv_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return PV_ERR_CUDA;
  }
  DcPlan* plan = new DcPlan();
  memset(&plan->p, 0, sizeof(DcParams));
  DcParams& p = plan->p;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    pv_set_error("pv_detconv_create: no CUDA device: %s", cudaGetErrorString(e));
    delete plan;
    return PV_ERR_CUDA;
  }
  for (int s = 0; s < (rowel1 > 0 ? 2 : 1); ++s) {
    const int segw = s == 0 ? rowel0 : rowel1;
    // stride 1: (C, pitch, H, B) pixels; stride 2: (2C, pitch/2, H, B) pixel pairs
    const cuuint64_t row_el = in.S == 2 ? 2 * (cuuint64_t)in.C : (cuuint64_t)in.C;
    const cuuint64_t npx = in.S == 2 ? (cuuint64_t)d->pitch / 2 : (cuuint64_t)d->pitch;
    // the valid width bounds the box reads: pixels >= W up to the pitch are zero in memory anyway
    cuuint64_t gdim[4] = {row_el, npx, (cuuint64_t)d->H, (cuuint64_t)d->B};
    cuuint64_t gstr[3] = {row_el * 2, (cuuint64_t)d->pitch * in.C * 2, (cuuint64_t)d->H * d->pitch * in.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)segw, (cuuint32_t)pw, (cuuint32_t)ph, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const int rowb = segw * 2;
    CUtensorMapSwizzle sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = enc(&p.in[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      pv_set_error("pv_detconv_create: cuTensorMapEncodeTiled failed: CUresult %d (seg %d, C=%d pitch=%d H=%d B=%d box %dx%dx%d)", (int)r,
                   s, in.C, d->pitch, d->H, d->B, segw, pw, ph);
      delete plan;
      return PV_ERR_CUDA;
    }
  }
  if (cudaMalloc(&plan->d_err, sizeof(int)) != cudaSuccess) {
    pv_set_error("pv_detconv_create: cudaMalloc failed");
    delete plan;
    return PV_ERR_CUDA;
  }
  cudaMemset(plan->d_err, 0, sizeof(int));
  p.w_img = static_cast<const uint8_t*>(d->w_img);
  p.w_bytes = w_bytes;
  p.scale = d->scale;
  p.shift = d->shift;
  p.out = d->out;
  p.B = d->B;
  p.OH = OH;
  p.OW = OW;
  p.out_pitch = d->out_pitch;
  p.out_cs = d->out_cs;
  p.tiles_x = (OW + kTW - 1) / kTW;
  p.tiles_y = (OH + kTH - 1) / kTH;
  p.relu = d->relu;
  p.n_stages = n_stages;
  p.err = plan->d_err;
  plan->kind = kind;
  plan->Bmax = d->B;
  plan->smem_bytes = (size_t)align1k((int)w_bytes) + (size_t)n_stages * stage_bytes + fixed + 1024;
  *out_handle = plan;
  return PV_OK;
}

extern "C" int pv_detconv_run(void* handle, int B, void* stream) {
  PV_REQUIRE(handle, "pv_detconv_run: null handle");
  DcPlan* plan = static_cast<DcPlan*>(handle);
  PV_REQUIRE(B > 0 && B <= plan->Bmax, "pv_detconv_run: B=%d outside [1,%d]", B, plan->Bmax);
  DcParams p = plan->p;
  p.B = B;
  const long long nt = (long long)B * p.tiles_x * p.tiles_y;
  PV_REQUIRE(nt < (1ll << 31), "pv_detconv_run: too many tiles");
  p.num_tiles = (int)nt;
  const int grid = p.num_tiles < plan->num_sms ? p.num_tiles : plan->num_sms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (plan->kind) {
    case 0: e = launch<16, 32, 5, 5, 2, false>(plan, p, grid, st); break;
    case 1: e = launch<32, 32, 5, 5, 2, false>(plan, p, grid, st); break;
    case 2: e = launch<32, 48, 5, 5, 1, false>(plan, p, grid, st); break;
    case 3: e = launch<48, 48, 5, 5, 1, false>(plan, p, grid, st); break;
    default: e = launch<48, 16, 9, 1, 1, true>(plan, p, grid, st); break;
  }
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(e);
  return PV_OK;
}

extern "C" int pv_detconv_info(void* handle, int* n_stages, int* smem_bytes, int* tiles_x, int* tiles_y) {
  PV_REQUIRE(handle, "pv_detconv_info: null handle");
  DcPlan* plan = static_cast<DcPlan*>(handle);
  if (n_stages) *n_stages = plan->p.n_stages;
  if (smem_bytes) *smem_bytes = (int)plan->smem_bytes;
  if (tiles_x) *tiles_x = plan->p.tiles_x;
  if (tiles_y) *tiles_y = plan->p.tiles_y;
  return PV_OK;
}

extern "C" int pv_detconv_check(void* handle, void* stream) {
  PV_REQUIRE(handle, "pv_detconv_check: null handle");
  DcPlan* plan = static_cast<DcPlan*>(handle);
  cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  int flag = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&flag, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    pv_set_error("pv_detconv_check: %s", cudaGetErrorString(e));
    return PV_ERR_CUDA;
  }
  if (flag != 0) {
    pv_set_error("detconv: device-side pipeline timeout (role code %d)", flag);
    cudaMemset(plan->d_err, 0, sizeof(int));
    return PV_ERR_DEVICE_TIMEOUT;
  }
  return PV_OK;
}

extern "C" int pv_detconv_destroy(void* handle) {
  if (!handle) return PV_OK;
  DcPlan* plan = static_cast<DcPlan*>(handle);
  cudaFree(plan->d_err);
  delete plan;
  return PV_OK;
}
