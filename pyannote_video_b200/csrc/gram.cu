// gram.cu — pairwise embedding distances on tcgen05 (sm_100a): the Gram matrix behind
// `-squareform(pdist(X, 'euclidean'))`, pyannote/video/face/clustering.py:101.
//
//   D[i][j] = sqrt(max(0, |x_i|^2 + |x_j|^2 - 2 <x_i, x_j>))         (metric 0, scipy 'euclidean')
//   D[i][j] = 1 - <x_i, x_j> / (|x_i| |x_j|)                           (metric 1, cosine distance)
//
// <x_i, x_j> comes from the tensor cores.  The reference computes in float64 and the parity tests hold the
// distances to 2e-5 of scipy's, which one bf16 product (8 mantissa bits) cannot give, so every float32 embedding is
// split into three bf16 parts x = hi + mid + lo (24 mantissa bits) and the six products whose weight is >= 2^-16
// are accumulated in fp32 TMEM:  hi.hi + hi.mid + mid.hi + hi.lo + lo.hi + mid.mid   (error ~ 2^-24 of the product).
// That is a GEMM with K = 6 x 128 on operands that stay in L2 (100 k embeddings x 3 parts x 256 B = 77 MB).
//
// One persistent CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..9 = epilogue (two per
// TMEM lane quarter, 64 columns each).  Work item = one 128-row block of i (its three parts, 96 KB, stay resident in
// shared memory) x a run of 128-column blocks of j streamed through a 3-stage ring, one part (32 KB) per stage.
// Accumulators: two tiles in flight, each two 128 x 128 fp32 accumulators in TMEM (all 512 columns), so the epilogue of
// one tile (tcgen05.ld -> norms -> sqrt -> 256-bit row stores) overlaps the 48 MMAs of the next one.  The output is the
// roofline: 4 N^2 bytes are written once.
//
// Accumulation order matters: the tensor core adds into the fp32 accumulator with truncation, i.e. every accumulating
// MMA costs up to an ulp OF THE ACCUMULATOR'S MAGNITUDE (measured: 48 adds into a sum of 16 gave |dG| ~ 6e-5, 1.3e-4 on D).
// So the five small products go first, while the accumulator is ~2^-8 of the final value, hi.hi goes last, and its
// second K half goes into a second accumulator that the epilogue adds in fp32: 4 + 4 adds at half magnitude.
#include <cuda.h>
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kDim = 128;
constexpr int kBlk = 128;            // rows per i block = columns per j block
constexpr int kStages = 3;
constexpr int kAcc = 2;              // tiles in flight in TMEM; a tile owns TWO 128-column accumulators (see the MMA order below)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kPartBytes = kBlk * kDim * 2;      // 32 KB: one part of one block (two 64-column swizzle-128B halves)
constexpr int kHalfBytes = kPartBytes / 2;

struct GramParams {
  CUtensorMap xs;          // bf16 [3 * npad rows][128], box [128 rows][64 columns], 128-byte swizzle
  const float* norms;      // [npad] squared norms
  float* D;                // [n][n]
  long long n;
  int npad;                // n rounded up to 128
  int n_blk;               // npad / 128
  int jrun;                // j blocks per work item
  int n_items;
  int metric;
  int* err;
};

__global__ void split3_kernel(const float* __restrict__ X, long long n, int npad, __nv_bfloat16* __restrict__ xs,
                              float* __restrict__ norms) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= npad) return;
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = lane + 32 * q;
    const float v = row < n ? X[(long long)row * kDim + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
    xs[((long long)0 * npad + row) * kDim + c] = hi;
    xs[((long long)1 * npad + row) * kDim + c] = mid;
    xs[((long long)2 * npad + row) * kDim + c] = lo;
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) norms[row] = s;
}

__global__ void __launch_bounds__(kThreads, 1) gram_kernel(const __grid_constant__ GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_sm = smem;                                   // 3 parts x 32 KB
  uint8_t* b_sm = smem + 3 * kPartBytes;                  // kStages x 32 KB
  uint64_t* bar_a_full = reinterpret_cast<uint64_t*>(b_sm + kStages * kPartBytes);
  uint64_t* bar_a_empty = bar_a_full + 1;
  uint64_t* bar_b_full = bar_a_empty + 1;
  uint64_t* bar_b_empty = bar_b_full + kStages;
  uint64_t* bar_acc_full = bar_b_empty + kStages;
  uint64_t* bar_acc_empty = bar_acc_full + kAcc;
  float* s_nj = reinterpret_cast<float*>(bar_acc_empty + kAcc);       // [kAcc][128]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_nj + kAcc * kBlk);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    pv_tma_prefetch_desc(&p.xs);
    pv_mbar_init(bar_a_full, 1);
    pv_mbar_init(bar_a_empty, 1);
    for (int i = 0; i < kStages; ++i) { pv_mbar_init(&bar_b_full[i], 1); pv_mbar_init(&bar_b_empty[i], 1); }
    for (int i = 0; i < kAcc; ++i) { pv_mbar_init(&bar_acc_full[i], 1); pv_mbar_init(&bar_acc_empty[i], kEpiWarps); }
    pv_fence_mbar_init();
  }
  if (warp == 1) pv_tmem_alloc(s_tmem, 512);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int runs_per_i = (p.n_blk + p.jrun - 1) / p.jrun;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (pv_elect_one()) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ib = item / runs_per_i, run = item - ib * runs_per_i;
        const int jb0 = run * p.jrun, jb1 = min(jb0 + p.jrun, p.n_blk);
        pv_mbar_wait(bar_a_empty, a_phase ^ 1u, p.err, 1);
        pv_mbar_arrive_expect_tx(bar_a_full, 3u * kPartBytes);
        for (int part = 0; part < 3; ++part)
          for (int h = 0; h < 2; ++h)
            pv_tma_load_2d(a_sm + part * kPartBytes + h * kHalfBytes, &p.xs, bar_a_full, h * 64, part * p.npad + ib * kBlk);
        a_phase ^= 1u;
        for (int jb = jb0; jb < jb1; ++jb)
          for (int part = 2; part >= 0; --part) {                 // lo, mid, hi: small products first
            pv_mbar_wait(&bar_b_empty[stage], phase ^ 1u, p.err, 2);
            pv_mbar_arrive_expect_tx(&bar_b_full[stage], (uint32_t)kPartBytes);
            for (int h = 0; h < 2; ++h)
              pv_tma_load_2d(b_sm + stage * kPartBytes + h * kHalfBytes, &p.xs, &bar_b_full[stage], h * 64, part * p.npad + jb * kBlk);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (pv_elect_one()) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBlk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_addr = pv_smem_u32(a_sm), b_addr = pv_smem_u32(b_sm);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0, acc_cnt = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int ib = item / runs_per_i, run = item - ib * runs_per_i;
        const int jb0 = run * p.jrun, jb1 = min(jb0 + p.jrun, p.n_blk);
        pv_mbar_wait(bar_a_full, a_phase, p.err, 3);
        a_phase ^= 1u;
        for (int jb = jb0; jb < jb1; ++jb) {
          const uint32_t slot = acc_cnt % kAcc;
          pv_mbar_wait(&bar_acc_empty[slot], ((acc_cnt / kAcc) & 1u) ^ 1u, p.err, 4);
          const uint32_t d0 = tmem_base + slot * (2 * kBlk), d1 = d0 + kBlk;
          uint32_t first = 1;
          for (int bp = 2; bp >= 0; --bp) {                     // B part: lo, mid, hi
            pv_mbar_wait(&bar_b_full[stage], phase, p.err, 5);
            pv_tc_fence_after();
            const int n_ap = 3 - bp;                            // A parts whose product with this B part is kept
            for (int ap = n_ap - 1; ap >= 0; --ap)              // smallest product first; hi.hi is the very last one
#pragma unroll
              for (int k = 0; k < 8; ++k) {                     // K = 128 in steps of 16: half k >> 2, 32-byte step inside the atom
                const uint32_t off = (uint32_t)((k >> 2) * kHalfBytes + (k & 3) * 32);
                const uint64_t da = pv_umma_desc(a_addr + ap * kPartBytes + off, 1024, 2, 0);
                const uint64_t db = pv_umma_desc(b_addr + stage * kPartBytes + off, 1024, 2, 0);
                if (bp == 0 && ap == 0 && k >= 4) {
                  pv_umma_bf16(d1, da, db, idesc, k == 4 ? 0u : 1u);   // second K half of hi.hi: its own accumulator
                } else {
                  pv_umma_bf16(d0, da, db, idesc, first ? 0u : 1u);
                  first = 0;
                }
              }
            pv_umma_commit(&bar_b_empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          pv_umma_commit(&bar_acc_full[slot]);
          ++acc_cnt;
        }
        pv_umma_commit(bar_a_empty);                            // the item's MMAs have read A
      }
    }
  } else {
    // ===================== epilogue: row i = TMEM lane; warps 2..5 take columns 0..63, warps 6..9 columns 64..127 =====
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                            // 0..255 among the epilogue threads
    const int grp = (warp - 2) >> 2;                            // column half
    const bool vec_ok = (p.n % 8) == 0;
    uint32_t acc_cnt = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int ib = item / runs_per_i, run = item - ib * runs_per_i;
      const int jb0 = run * p.jrun, jb1 = min(jb0 + p.jrun, p.n_blk);
      const long long i = (long long)ib * kBlk + m;
      const float ni = p.norms[ib * kBlk + m];
      const float sni = sqrtf(ni);
      for (int jb = jb0; jb < jb1; ++jb) {
        const uint32_t slot = acc_cnt % kAcc;
        float* nj = s_nj + slot * kBlk;
        if (et < kBlk) nj[et] = p.norms[jb * kBlk + et];
        pv_mbar_wait(&bar_acc_full[slot], (acc_cnt / kAcc) & 1u, p.err, 6);
        pv_tc_fence_after();
        asm volatile("bar.sync 1, 256;" ::: "memory");           // nj[] of this slot is complete
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * (2 * kBlk) + grp * 64;
        const long long j0 = (long long)jb * kBlk + grp * 64;
        float* drow = p.D + i * p.n + j0;
        const int diag = (ib == jb) ? m - grp * 64 : -1;         // column of this thread's diagonal element, if in range
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v0[2][16], v1[2][16];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            pv_tmem_ld16(taddr + half * 32 + c * 16, v0[c]);
            pv_tmem_ld16(taddr + kBlk + half * 32 + c * 16, v1[c]);
          }
          pv_tmem_ld_wait();
          if (half == 1) {
            pv_tc_fence_before();
            __syncwarp();
            if (lane == 0) pv_mbar_arrive(&bar_acc_empty[slot]);   // this warp's part of both accumulators is in registers
          }
          if (i < p.n) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int col = half * 32 + c * 16;
              float f[16];
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float g = __uint_as_float(v0[c][k]) + __uint_as_float(v1[c][k]);
                const float njk = nj[grp * 64 + col + k];
                float dv;
                if (p.metric == 0) {
                  const float d2 = fmaxf(ni + njk - 2.0f * g, 0.f);
                  asm("sqrt.approx.f32 %0, %1;" : "=f"(dv) : "f"(d2));
                } else {
                  dv = 1.f - g / fmaxf(sni * sqrtf(njk), 1e-30f);
                }
                if (col + k == diag) dv = 0.f;
                f[k] = dv;
              }
              if (vec_ok && j0 + col + 16 <= p.n) {
                pv_stg256(drow + col, __float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]),
                          __float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
                pv_stg256(drow + col + 8, __float_as_uint(f[8]), __float_as_uint(f[9]), __float_as_uint(f[10]),
                          __float_as_uint(f[11]), __float_as_uint(f[12]), __float_as_uint(f[13]), __float_as_uint(f[14]),
                          __float_as_uint(f[15]));
              } else {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  if (j0 + col + k < p.n) drow[col + k] = f[k];
              }
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");           // nj[] of this slot may be overwritten two tiles later
        ++acc_cnt;
      }
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn gram_encode_fn() {
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

constexpr size_t kGramSmem = 1024 + 3 * kPartBytes + kStages * kPartBytes + 16 * 8 + kAcc * kBlk * 4 + 64;

}  // namespace

/* workspace sizes for pv_gram_dist: xs = 3 * npad * 128 bf16, norms = npad floats, npad = n rounded up to 128 */
extern "C" int64_t pv_gram_npad(int64_t n) { return (n + kBlk - 1) / kBlk * kBlk; }

/* D[i][j] (float32 [n][n]) = metric(x_i, x_j) via tcgen05; X float32 [n][128]; xs_ws / norms_ws: device workspaces */
extern "C" int pv_gram_dist(const float* X, int64_t n, int dim, int metric, float* D, void* xs_ws, float* norms_ws, int* err_flag,
                            void* stream) {
  PV_REQUIRE(X && D && xs_ws && norms_ws, "pv_gram_dist: null argument");
  PV_REQUIRE(dim == kDim, "pv_gram_dist: dim=%d (must be 128)", dim);
  PV_REQUIRE(metric == 0 || metric == 1, "pv_gram_dist: metric=%d", metric);
  if (n == 0) return PV_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long npad = pv_gram_npad(n);
  PV_REQUIRE(3 * npad < (1ll << 31), "pv_gram_dist: n=%lld too large", (long long)n);
  split3_kernel<<<(unsigned)((npad + 7) / 8), 256, 0, st>>>(X, n, (int)npad, static_cast<__nv_bfloat16*>(xs_ws), norms_ws);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  EncodeTiledFn enc = gram_encode_fn();
  PV_REQUIRE(enc != nullptr, "pv_gram_dist: cuTensorMapEncodeTiled unavailable");
  GramParams p;
  const cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)(3 * npad)};
  const cuuint64_t strides[1] = {(cuuint64_t)(kDim * 2)};
  const cuuint32_t box[2] = {64, (cuuint32_t)kBlk};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&p.xs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs_ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PV_REQUIRE(r == CUDA_SUCCESS, "pv_gram_dist: cuTensorMapEncodeTiled failed (%d)", (int)r);
  p.norms = norms_ws;
  p.D = D;
  p.n = n;
  p.npad = (int)npad;
  p.n_blk = (int)(npad / kBlk);
  p.jrun = p.n_blk < 16 ? p.n_blk : 16;
  p.n_items = p.n_blk * ((p.n_blk + p.jrun - 1) / p.jrun);
  p.metric = metric;
  p.err = err_flag;
  int dev = 0, sms = 0;
  PV_CUDA_CHECK(cudaGetDevice(&dev));
  PV_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static unsigned long long attr = 0;
  if (pv_attr_needed(&attr))
    PV_CUDA_CHECK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramSmem));
  const int grid = p.n_items < sms ? p.n_items : sms;
  gram_kernel<<<grid, kThreads, kGramSmem, st>>>(p);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
