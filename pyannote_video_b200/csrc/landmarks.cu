// landmarks.cu — 68-point ERT shape predictor and face-chip extraction.
//   ert_forward   replaces dlib.shape_predictor.__call__            (pyannote/video/face/face.py:70)
//   chip_extract  replaces get_face_chip_details + extract_image_chip inside
//                 face_recognition_model_v1.compute_face_descriptor (pyannote/video/face/face.py:74-75)
// Both are gather-bound: one CTA per face, the 65 MB leaf table streams from L2/HBM with coalesced
// 544-byte rows, the per-stage similarity fit is done by one thread in the oracle's summation order
// (float32, unfused) so results are bit-identical to oracle/landmarks.py.
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kPts = 68;
constexpr int kMaxPool = 512;
constexpr int kMaxTrees = 512;
constexpr int kChunk = 64;            // trees staged per shared-memory chunk
constexpr int kRow4 = 2 * kPts / 4;    // float4 per leaf row (136 floats)

struct Sim {
  float m00, m01, m10, m11, tx, ty;
};

// closed-form 2-D similarity b ~ M a + t over the points listed in idx (or all if idx == nullptr);
// sequential float32 sums, every operation explicitly rounded (no FMA contraction).
__device__ Sim similarity_fit(const float* a, const float* b, const int* idx, int n) {
  float sax = 0.f, say = 0.f, sbx = 0.f, sby = 0.f;
  for (int k = 0; k < n; ++k) {
    const int i = idx ? idx[k] : k;
    sax = __fadd_rn(sax, a[2 * i]);
    say = __fadd_rn(say, a[2 * i + 1]);
    sbx = __fadd_rn(sbx, b[2 * i]);
    sby = __fadd_rn(sby, b[2 * i + 1]);
  }
  const float fn = (float)n;
  const float max_ = __fdiv_rn(sax, fn), may = __fdiv_rn(say, fn), mbx = __fdiv_rn(sbx, fn), mby = __fdiv_rn(sby, fn);
  float A = 0.f, Bc = 0.f, den = 0.f;
  for (int k = 0; k < n; ++k) {
    const int i = idx ? idx[k] : k;
    const float acx = __fsub_rn(a[2 * i], max_), acy = __fsub_rn(a[2 * i + 1], may);
    const float bcx = __fsub_rn(b[2 * i], mbx), bcy = __fsub_rn(b[2 * i + 1], mby);
    A = __fadd_rn(A, __fadd_rn(__fmul_rn(acx, bcx), __fmul_rn(acy, bcy)));
    Bc = __fadd_rn(Bc, __fsub_rn(__fmul_rn(acx, bcy), __fmul_rn(acy, bcx)));
    den = __fadd_rn(den, __fadd_rn(__fmul_rn(acx, acx), __fmul_rn(acy, acy)));
  }
  Sim s;
  s.m00 = __fdiv_rn(A, den);
  s.m10 = __fdiv_rn(Bc, den);
  s.m01 = -s.m10;
  s.m11 = s.m00;
  s.tx = __fsub_rn(mbx, __fadd_rn(__fmul_rn(s.m00, max_), __fmul_rn(s.m01, may)));
  s.ty = __fsub_rn(mby, __fadd_rn(__fmul_rn(s.m10, max_), __fmul_rn(s.m11, may)));
  return s;
}

struct ErtModel {
  const float* initial_shape;  // [136]
  const int* anchor_idx;       // [S,P]
  const float* deltas;         // [S,P,2]
  const int* split_idx1;       // [S,T,15]
  const int* split_idx2;       // [S,T,15]
  const float* split_thresh;   // [S,T,15]
  const float* leaf_values;    // [S,T,16,136]
  int stages, trees, pool;
};

// frames: uint8 [F,H,W,3]; rects int32 [M,4] (l,t,r,b); frame_idx int32 [M]; out int32 [M,68,2]
__global__ void __launch_bounds__(256) ert_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                  const int* __restrict__ rects, const int* __restrict__ frame_idx,
                                                  ErtModel m, int* __restrict__ out) {
  __shared__ float s_init[2 * kPts];
  __shared__ float s_cur[2 * kPts];
  __shared__ float s_feat[kMaxPool];
  __shared__ uint8_t s_leaf[kMaxTrees];
  __shared__ float4 s_rows[kChunk * kRow4];
  __shared__ Sim s_sim;
  const int face = blockIdx.x;
  const int tid = threadIdx.x;
  const uint8_t* img = frames + (long long)frame_idx[face] * H * W * 3;
  const int rl = rects[4 * face], rt = rects[4 * face + 1], rr = rects[4 * face + 2], rb = rects[4 * face + 3];
  const float l = (float)rl, t = (float)rt, wr = (float)(rr - rl), hr = (float)(rb - rt);
  for (int i = tid; i < 2 * kPts; i += blockDim.x) {
    s_init[i] = m.initial_shape[i];
    s_cur[i] = m.initial_shape[i];
  }
  __syncthreads();
  for (int s = 0; s < m.stages; ++s) {
    if (tid == 0) s_sim = similarity_fit(s_init, s_cur, nullptr, kPts);
    __syncthreads();
    const Sim sm = s_sim;
    for (int i = tid; i < m.pool; i += blockDim.x) {
      const int an = m.anchor_idx[s * m.pool + i];
      const float d0 = m.deltas[(s * m.pool + i) * 2], d1 = m.deltas[(s * m.pool + i) * 2 + 1];
      const float dx = __fadd_rn(__fadd_rn(__fmul_rn(sm.m00, d0), __fmul_rn(sm.m01, d1)), s_cur[2 * an]);
      const float dy = __fadd_rn(__fadd_rn(__fmul_rn(sm.m10, d0), __fmul_rn(sm.m11, d1)), s_cur[2 * an + 1]);
      const float px = __fadd_rn(l, __fmul_rn(dx, wr));
      const float py = __fadd_rn(t, __fmul_rn(dy, hr));
      const int ix = (int)floorf(__fadd_rn(px, 0.5f)), iy = (int)floorf(__fadd_rn(py, 0.5f));
      float f = 0.f;
      if (ix >= 0 && ix < W && iy >= 0 && iy < H) {
        const uint8_t* p = img + ((long long)iy * W + ix) * 3;
        f = (float)(((unsigned)p[0] + (unsigned)p[1] + (unsigned)p[2]) / 3u);
      }
      s_feat[i] = f;
    }
    __syncthreads();
    for (int tr = tid; tr < m.trees; tr += blockDim.x) {
      const long long base = ((long long)s * m.trees + tr) * 15;
      int node = 0;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const float diff = __fsub_rn(s_feat[m.split_idx1[base + node]], s_feat[m.split_idx2[base + node]]);
        node = (diff > m.split_thresh[base + node]) ? 2 * node + 1 : 2 * node + 2;
      }
      s_leaf[tr] = (uint8_t)(node - 15);
    }
    __syncthreads();
    // leaf-vector accumulation: rows are staged through shared memory in chunks of kChunk trees
    // (coalesced 16-byte loads, many in flight), then each coordinate adds its column in tree
    // order — the oracle's summation order, so the float32 result is bit-identical.
    {
      const float4* lv4 = reinterpret_cast<const float4*>(m.leaf_values + (long long)s * m.trees * 16 * (2 * kPts));
      float acc = (tid < 2 * kPts) ? s_cur[tid] : 0.f;
      for (int t0 = 0; t0 < m.trees; t0 += kChunk) {
        const int nt = min(kChunk, m.trees - t0);
        for (int e = tid; e < nt * kRow4; e += blockDim.x) {
          const int tr = e / kRow4, c4 = e - tr * kRow4;
          s_rows[e] = lv4[((long long)(t0 + tr) * 16 + s_leaf[t0 + tr]) * kRow4 + c4];
        }
        __syncthreads();
        if (tid < 2 * kPts) {
          const float* rows = reinterpret_cast<const float*>(s_rows);
#pragma unroll 4
          for (int tr = 0; tr < nt; ++tr) acc = __fadd_rn(acc, rows[tr * (2 * kPts) + tid]);
        }
        __syncthreads();
      }
      if (tid < 2 * kPts) s_cur[tid] = acc;
    }
    __syncthreads();
  }
  if (tid < kPts) {
    const float x = __fadd_rn(l, __fmul_rn(s_cur[2 * tid], wr));
    const float y = __fadd_rn(t, __fmul_rn(s_cur[2 * tid + 1], hr));
    out[((long long)face * kPts + tid) * 2] = (int)floorf(__fadd_rn(x, 0.5f));
    out[((long long)face * kPts + tid) * 2 + 1] = (int)floorf(__fadd_rn(y, 0.5f));
  }
}

// chip extraction: frames uint8 [F,H,W,3], parts int32 [M,68,2] -> RGBA u8 [M,size,size,4] (A=255)
__global__ void __launch_bounds__(256) chip_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                   const int* __restrict__ parts, const int* __restrict__ frame_idx,
                                                   const float* __restrict__ from_pts,  // [68,2] chip-space targets
                                                   const int* __restrict__ pt_idx, int n_idx, int size,
                                                   uchar4* __restrict__ out) {
  __shared__ float s_from[2 * kPts];
  __shared__ float s_to[2 * kPts];
  __shared__ int s_idx[kPts];
  __shared__ Sim s_sim;
  const int face = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * kPts; i += blockDim.x) {
    s_from[i] = from_pts[i];
    s_to[i] = (float)parts[(long long)face * 2 * kPts + i];
  }
  for (int i = tid; i < n_idx; i += blockDim.x) s_idx[i] = pt_idx[i];
  __syncthreads();
  if (tid == 0) s_sim = similarity_fit(s_from, s_to, s_idx, n_idx);
  __syncthreads();
  const Sim sm = s_sim;
  const uint8_t* img = frames + (long long)frame_idx[face] * H * W * 3;
  for (int p = tid; p < size * size; p += blockDim.x) {
    const int r = p / size, c = p - r * size;
    const float fc = (float)c, fr = (float)r;
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(sm.m00, fc), __fmul_rn(sm.m01, fr)), sm.tx);
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(sm.m10, fc), __fmul_rn(sm.m11, fr)), sm.ty);
    const int left = (int)floorf(x), top = (int)floorf(y);
    uchar4 o = make_uchar4(0, 0, 0, 255);
    // left >= 0 && left + 1 < W on the floats: a degenerate similarity fit (inf / NaN / beyond int range) must read nothing
    if (x >= 0.f && y >= 0.f && x < (float)(W - 1) && y < (float)(H - 1)) {
      const float lr = __fsub_rn(x, (float)left), tb = __fsub_rn(y, (float)top);
      const float omlr = __fsub_rn(1.0f, lr), omtb = __fsub_rn(1.0f, tb);
      const uint8_t* ptl = img + ((long long)top * W + left) * 3;
      const uint8_t* pbl = ptl + (long long)W * 3;
      uint8_t res[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float a = __fadd_rn(__fmul_rn(omlr, (float)ptl[ch]), __fmul_rn(lr, (float)ptl[3 + ch]));
        const float b = __fadd_rn(__fmul_rn(omlr, (float)pbl[ch]), __fmul_rn(lr, (float)pbl[3 + ch]));
        float v = __fadd_rn(__fmul_rn(omtb, a), __fmul_rn(tb, b));
        v = fminf(fmaxf(floorf(__fadd_rn(v, 0.5f)), 0.f), 255.f);
        res[ch] = (uint8_t)v;
      }
      o.x = res[0];
      o.y = res[1];
      o.z = res[2];
    }
    out[(long long)face * size * size + p] = o;
  }
}

struct ErtHandle {
  ErtModel m;
};

}  // namespace

extern "C" int pv_ert_create(const float* initial_shape, const int* anchor_idx, const float* deltas,
                             const int* split_idx1, const int* split_idx2, const float* split_thresh,
                             const float* leaf_values, int stages, int trees, int pool, void** out_handle) {
  PV_REQUIRE(initial_shape && anchor_idx && deltas && split_idx1 && split_idx2 && split_thresh && leaf_values &&
                 out_handle,
             "pv_ert_create: null argument");
  PV_REQUIRE(trees <= kMaxTrees && pool <= kMaxPool && stages > 0, "pv_ert_create: trees=%d pool=%d", trees, pool);
  ErtHandle* h = new ErtHandle();
  h->m.initial_shape = initial_shape;
  h->m.anchor_idx = anchor_idx;
  h->m.deltas = deltas;
  h->m.split_idx1 = split_idx1;
  h->m.split_idx2 = split_idx2;
  h->m.split_thresh = split_thresh;
  h->m.leaf_values = leaf_values;
  h->m.stages = stages;
  h->m.trees = trees;
  h->m.pool = pool;
  *out_handle = h;
  return PV_OK;
}

extern "C" int pv_ert_destroy(void* handle) {
  delete static_cast<ErtHandle*>(handle);
  return PV_OK;
}

extern "C" int pv_ert_forward(void* handle, const void* frames, int H, int W, const int* rects, const int* frame_idx,
                              int M, int* out_parts, void* stream) {
  PV_REQUIRE(handle && frames && rects && frame_idx && out_parts, "pv_ert_forward: null argument");
  if (M == 0) return PV_OK;
  ErtHandle* h = static_cast<ErtHandle*>(handle);
  ert_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(frames), H, W, rects,
                                                                frame_idx, h->m, out_parts);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_chip_extract(const void* frames, int H, int W, const int* parts, const int* frame_idx, int M,
                               const float* from_pts, const int* pt_idx, int n_idx, int size, void* out_rgba,
                               void* stream) {
  PV_REQUIRE(frames && parts && frame_idx && from_pts && pt_idx && out_rgba, "pv_chip_extract: null argument");
  PV_REQUIRE(n_idx > 1 && n_idx <= kPts, "pv_chip_extract: n_idx=%d", n_idx);
  if (M == 0) return PV_OK;
  chip_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint8_t*>(frames), H, W, parts,
                                                                 frame_idx, from_pts, pt_idx, n_idx, size,
                                                                 static_cast<uchar4*>(out_rgba));
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
