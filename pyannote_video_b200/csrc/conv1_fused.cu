// conv1_fused.cu — first detector convolution (dlib con<16,5,5,2,2> on the tiled RGB pyramid, the
// input stage of face_detector_(rgb, 1), pyannote/video/face/face.py:66) reading the RGBA u8 plane
// directly.
//
// The generic path first writes a "gathered" bf16 matrix (16 B per plane pixel) and reads it back
// through TMA; with ~32 M plane pixels per 1080p frame that round trip is the largest HBM stream of
// the whole detector.  Here four producer warps build the 128 x 16 A tile of every filter row in
// shared memory themselves — normalise (v - mean)/256, convert to bf16, store in the 32-byte-swizzled
// K-major layout tcgen05.mma expects — so the plane is read once (4 B/pixel) and nothing is written.
// MMA issue, TMEM accumulators and the fused affine + ReLU epilogue are as in srgemm.cu.
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kTileM = 128;
constexpr int kKH = 5, kKW = 5;
constexpr int kN = 16;
constexpr int kRing = 4;
constexpr int kAcc = 2;
constexpr int kSlabBytes = kTileM * 32;           // 128 rows x 16 bf16
constexpr int kSlotBytes = kKH * kSlabBytes;      // 20 KB
constexpr int kWBytes = kKH * kN * 32;            // 2.5 KB
constexpr int kProdWarps = 8;                     // rows x {filter rows 0-2 | 3-4}
constexpr int kThreads = 32 * (kProdWarps + 1 + 4);  // producer warps, 1 MMA warp, 4 epilogue warps

struct C1Params {
  const uchar4* plane;   // [B, Hp, Wp] RGBA
  int B, Hp, Wp;
  int hq, wq, oh, ow;    // output grid (ceil(Hp/2), ceil(Wp/2)) and valid extent
  const __nv_bfloat16* w;  // [5][16][16] (k = kw*3 + c, k = 15 is zero)
  const float* scale;
  const float* shift;
  int relu;
  __nv_bfloat16* out;
  PvRowMap dst;
  float m0, m1, m2;
  long long q_rows;
  int num_tiles;
  int* err;
};

__device__ __forceinline__ long long row_of(const PvRowMap& m, uint32_t n, uint32_t y, uint32_t x) {
  const uint32_t Y = y + m.py, X = x + m.px;
  if (m.kind == 0) return (long long)n * m.img + (long long)Y * m.w + X;
  const uint32_t plane = ((Y & 1u) << 1) | (X & 1u);
  return (long long)plane * m.plane_rows + (long long)n * m.img + (long long)(Y >> 1) * m.w + (X >> 1);
}

// byte offset of 16-byte chunk `c` (0/1) of row `r` in a 32-byte-swizzled K-major tile whose base is
// 256-byte aligned: Swizzle<1,4,3> = address bit 4 ^= address bit 7
__device__ __forceinline__ uint32_t sw32(uint32_t r, uint32_t c) { return r * 32u + ((c ^ ((r >> 2) & 1u)) << 4); }

// filter rows [KH0, KH1) of tile row m: pixel loads are issued up front with clamped coordinates (no
// branches around them) so they are all in flight together; validity is applied afterwards
template <int KH0, int KH1>
__device__ __forceinline__ void produce_rows(const C1Params& p, bool in_range, uint32_t n, uint32_t oy, uint32_t ox,
                                             uint32_t m, uint8_t* ring, uint64_t* bar_empty, uint32_t phase, int slot) {
  uchar4 px[KH1 - KH0][kKW];
  const int x0 = 2 * (int)ox;
#pragma unroll
  for (int kh = KH0; kh < KH1; ++kh) {
    const int y = min(2 * (int)oy + kh, p.Hp - 1);
    const uchar4* src = p.plane + ((long long)min(n, (uint32_t)(p.B - 1)) * p.Hp + y) * p.Wp;
#pragma unroll
    for (int kw = 0; kw < kKW; ++kw) px[kh - KH0][kw] = __ldg(src + min(x0 + kw, p.Wp - 1));
  }
  pv_mbar_wait(bar_empty, phase ^ 1u, p.err, 1);
  uint8_t* base = ring + slot * kSlotBytes;
#pragma unroll
  for (int kh = KH0; kh < KH1; ++kh) {
    const bool row_ok = in_range && (2 * (int)oy + kh < p.Hp);
    float f[16];
    f[15] = 0.f;
#pragma unroll
    for (int kw = 0; kw < kKW; ++kw) {
      const uchar4 q4 = px[kh - KH0][kw];
      const bool ok = row_ok && (x0 + kw < p.Wp) && (q4.w != 0);
      f[kw * 3 + 0] = ok ? __fmul_rn(__fsub_rn((float)q4.x, p.m0), 0.00390625f) : 0.f;
      f[kw * 3 + 1] = ok ? __fmul_rn(__fsub_rn((float)q4.y, p.m1), 0.00390625f) : 0.f;
      f[kw * 3 + 2] = ok ? __fmul_rn(__fsub_rn((float)q4.z, p.m2), 0.00390625f) : 0.f;
    }
    uint4 c0, c1;
    c0.x = pv_pack_bf16x2(f[0], f[1]);
    c0.y = pv_pack_bf16x2(f[2], f[3]);
    c0.z = pv_pack_bf16x2(f[4], f[5]);
    c0.w = pv_pack_bf16x2(f[6], f[7]);
    c1.x = pv_pack_bf16x2(f[8], f[9]);
    c1.y = pv_pack_bf16x2(f[10], f[11]);
    c1.z = pv_pack_bf16x2(f[12], f[13]);
    c1.w = pv_pack_bf16x2(f[14], f[15]);
    uint8_t* slab = base + kh * kSlabBytes;
    *reinterpret_cast<uint4*>(slab + sw32(m, 0)) = c0;
    *reinterpret_cast<uint4*>(slab + sw32(m, 1)) = c1;
  }
}

__global__ void __launch_bounds__(kThreads, 2) conv1_fused_kernel(const __grid_constant__ C1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                               // kRing slots of 5 slabs
  uint8_t* wsm = ring + kRing * kSlotBytes;           // swizzled weights (1024-aligned: 20 KB slots)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(wsm + ((kWBytes + 1023) & ~1023));
  uint64_t* bar_empty = bar_full + kRing;
  uint64_t* bar_tfull = bar_empty + kRing;
  uint64_t* bar_tempty = bar_tfull + kAcc;
  float* s_scale = reinterpret_cast<float*>(bar_tempty + kAcc);
  float* s_shift = s_scale + kN;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_shift + kN);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- setup ----
  for (int i = threadIdx.x; i < kKH * kN * 2; i += kThreads) {   // one 16-byte chunk per iteration
    const int kh = i / (kN * 2), rem = i - kh * kN * 2;
    const int n = rem >> 1, c = rem & 1;
    const uint4 v = *reinterpret_cast<const uint4*>(p.w + ((kh * kN + n) * 16 + c * 8));
    *reinterpret_cast<uint4*>(wsm + kh * (kN * 32) + sw32((uint32_t)n, (uint32_t)c)) = v;
  }
  if (threadIdx.x < kN) {
    s_scale[threadIdx.x] = p.scale[threadIdx.x];
    s_shift[threadIdx.x] = p.shift[threadIdx.x];
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kRing; ++i) {
      pv_mbar_init(&bar_full[i], kProdWarps);     // one arrival per producer warp
      pv_mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      pv_mbar_init(&bar_tfull[i], 1);
      pv_mbar_init(&bar_tempty[i], 4);
    }
    pv_fence_mbar_init();
  }
  pv_fence_proxy_async();                  // weights written through the generic proxy, read by the tensor core
  if (warp == kProdWarps) pv_tmem_alloc(s_tmem, 32);
  pv_tc_fence_before();
  __syncthreads();
  pv_tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t img = (uint32_t)p.hq * (uint32_t)p.wq;

  if (warp < kProdWarps) {
    // ===================== producers: build the A slabs =====================
    const uint32_t m = (uint32_t)threadIdx.x & 127u;   // row of the tile
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t q = (uint32_t)tile * kTileM + m;
      const bool in_range = (long long)q < p.q_rows;
      const uint32_t n = q / img;
      const uint32_t rem = q - n * img;
      const uint32_t oy = rem / (uint32_t)p.wq;
      const uint32_t ox = rem - oy * (uint32_t)p.wq;
      if (threadIdx.x >> 7)
        produce_rows<3, 5>(p, in_range, n, oy, ox, m, ring, &bar_empty[slot], phase, slot);
      else
        produce_rows<0, 3>(p, in_range, n, oy, ox, m, ring, &bar_empty[slot], phase, slot);
      pv_fence_proxy_async();        // make this thread's smem writes visible to the async (tensor-core) proxy
      __syncwarp();
      if (lane == 0) pv_mbar_arrive(&bar_full[slot]);
      if (++slot == kRing) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == kProdWarps) {
    // ===================== MMA issuer =====================
    const uint32_t lead = pv_elect_one() ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    const uint32_t hi = ((256u >> 4)) | (1u << 14) | (6u << 29);   // SBO = 256 B, version 1, SWIZZLE_32B
    const uint32_t w_lo = ((pv_smem_u32(wsm) & 0x3FFFFu) >> 4) | (1u << 16);
    int slot = 0, buf = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      pv_mbar_wait(&bar_tempty[buf], aphase ^ 1u, p.err, 2);
      pv_mbar_wait(&bar_full[slot], phase, p.err, 3);
      pv_tc_fence_after();
      const uint32_t a_lo = ((pv_smem_u32(ring + slot * kSlotBytes) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kN);
#pragma unroll
      for (int kh = 0; kh < kKH; ++kh)
        pv_umma_bf16_pred(tmem_d, ((uint64_t)hi << 32) | (a_lo + kh * (kSlabBytes >> 4)),
                          ((uint64_t)hi << 32) | (w_lo + kh * ((kN * 32) >> 4)), idesc, kh > 0 ? 1u : 0u, lead);
      pv_umma_commit_pred(&bar_empty[slot], lead);
      pv_umma_commit_pred(&bar_tfull[buf], lead);
      if (++slot == kRing) { slot = 0; phase ^= 1u; }
      if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const uint32_t m = (uint32_t)(quarter * 32 + lane);
    int buf = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const uint32_t q = (uint32_t)tile * kTileM + m;
      bool valid = (long long)q < p.q_rows;
      const uint32_t n = q / img;
      const uint32_t rem = q - n * img;
      const uint32_t y = rem / (uint32_t)p.wq;
      const uint32_t x = rem - y * (uint32_t)p.wq;
      valid = valid && (y < (uint32_t)p.oh) && (x < (uint32_t)p.ow);
      const long long drow = row_of(p.dst, n, y, x);
      pv_mbar_wait(&bar_tfull[buf], aphase, p.err, 4);
      pv_tc_fence_after();
      uint32_t v[16];
      pv_tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * kN), v);
      pv_tmem_ld_wait();
      if (valid) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f[j] = fmaf(__uint_as_float(v[j]), s_scale[j], s_shift[j]);
          if (p.relu) f[j] = fmaxf(f[j], 0.f);
        }
        uint4 o0, o1;
        o0.x = pv_pack_bf16x2(f[0], f[1]);
        o0.y = pv_pack_bf16x2(f[2], f[3]);
        o0.z = pv_pack_bf16x2(f[4], f[5]);
        o0.w = pv_pack_bf16x2(f[6], f[7]);
        o1.x = pv_pack_bf16x2(f[8], f[9]);
        o1.y = pv_pack_bf16x2(f[10], f[11]);
        o1.z = pv_pack_bf16x2(f[12], f[13]);
        o1.w = pv_pack_bf16x2(f[14], f[15]);
        uint4* dp = reinterpret_cast<uint4*>(p.out + drow * p.dst.cols);
        dp[0] = o0;
        dp[1] = o1;
      }
      pv_tc_fence_before();
      __syncwarp();
      if (lane == 0) pv_mbar_arrive(&bar_tempty[buf]);
      if (++buf == kAcc) { buf = 0; aphase ^= 1u; }
    }
  }
  pv_tc_fence_before();
  __syncthreads();
  if (warp == kProdWarps) {
    __syncwarp();
    pv_tmem_dealloc(tmem_base, 32);
  }
}

constexpr size_t kSmemBytes = 1024 + kRing * kSlotBytes + ((kWBytes + 1023) & ~1023) + (2 * kRing + 2 * kAcc) * 8 + 2 * kN * 4 + 64;

}  // namespace

extern "C" int pv_conv1_fused(const void* plane_rgba, int B, int Hp, int Wp, const void* w_bf16, const float* scale,
                              const float* shift, int relu, void* out, const PvRowMap* dst, int oh, int ow,
                              const float* mean_host, int* err_flag, void* stream) {
  PV_REQUIRE(plane_rgba && w_bf16 && scale && shift && out && dst && mean_host && err_flag, "pv_conv1_fused: null argument");
  PV_REQUIRE(dst->cols >= kN && dst->cols % 8 == 0, "pv_conv1_fused: dst.cols=%d", dst->cols);
  C1Params p;
  p.plane = static_cast<const uchar4*>(plane_rgba);
  p.B = B;
  p.Hp = Hp;
  p.Wp = Wp;
  p.hq = (Hp + 1) / 2;
  p.wq = (Wp + 1) / 2;
  p.oh = oh;
  p.ow = ow;
  p.w = static_cast<const __nv_bfloat16*>(w_bf16);
  p.scale = scale;
  p.shift = shift;
  p.relu = relu;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.dst = *dst;
  p.m0 = mean_host[0];
  p.m1 = mean_host[1];
  p.m2 = mean_host[2];
  p.q_rows = (long long)B * p.hq * p.wq;
  PV_REQUIRE(p.q_rows < (1ll << 31), "pv_conv1_fused: too many rows");
  p.num_tiles = (int)((p.q_rows + kTileM - 1) / kTileM);
  p.err = err_flag;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    PV_CUDA_CHECK(cudaGetDevice(&dev));
    PV_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    PV_CUDA_CHECK(cudaFuncSetAttribute(conv1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  }
  int grid = num_sms * 2;
  if (grid > p.num_tiles) grid = p.num_tiles;
  conv1_fused_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(p);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
