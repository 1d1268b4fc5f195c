// cluster.cu — pairwise distances + average-linkage agglomeration over face-track embeddings.
// Replaces _Model.compute_similarity_matrix / compute_similarity (scipy pdist + per-pair np.mean
// + pyannote.algorithms greedy HAC), pyannote/video/face/clustering.py:92-119,138-148.
//
// The cluster-pair quantity kept on the device is the SUM of embedding distances between two
// clusters, S[A][C] = sum_{i in A, j in C} d(i,j); average linkage = S / (|A||C|) and a merge is
// S[A u B][C] = S[A][C] + S[B][C] (exact Lance-Williams for UPGMA).  One agglomeration round =
//   row_argmin (nearest cluster of every cluster)  ->  host picks reciprocal pairs under the
//   threshold  ->  pool_rows + pool_cols contract the matrix  (S' = P S P^T).
// Average linkage is reducible, so merging all reciprocal-nearest pairs per round yields the same
// partition at the threshold as the reference's one-pair-at-a-time greedy loop (absent ties).
#include <atomic>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

extern std::atomic<long long> g_pv_launches;

namespace {

constexpr int kDim = 128;
constexpr int kTile = 64;

// D[i][j] = metric(x_i, x_j); 64x64 output tile per CTA (256 threads, 4x4 outputs each),
// operands staged through shared memory in fp32.  metric 0: Euclidean, 1: cosine distance.
__global__ void __launch_bounds__(256) pdist_kernel(const float* __restrict__ X, long long n, float* __restrict__ D,
                                                    int metric) {
  constexpr int kHalf = 64;
  __shared__ float sa[kTile][kHalf + 1];
  __shared__ float sb[kTile][kHalf + 1];
  const long long i0 = (long long)blockIdx.y * kTile, j0 = (long long)blockIdx.x * kTile;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4], na[4], nb[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    na[a] = nb[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  }
  for (int k0 = 0; k0 < kDim; k0 += kHalf) {
    __syncthreads();
    for (int e = tid; e < kTile * kHalf; e += 256) {
      const int r = e / kHalf, c = e - r * kHalf;
      sa[r][c] = (i0 + r < n) ? X[(i0 + r) * kDim + k0 + c] : 0.f;
      sb[r][c] = (j0 + r < n) ? X[(j0 + r) * kDim + k0 + c] : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < kHalf; ++k) {
      float va[4], vb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        va[a] = sa[ty + 16 * a][k];
        vb[a] = sb[tx + 16 * a][k];
      }
      if (metric == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float d = va[a] - vb[b];
            acc[a][b] = fmaf(d, d, acc[a][b]);
          }
      } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          na[a] = fmaf(va[a], va[a], na[a]);
          nb[a] = fmaf(vb[a], vb[a], nb[a]);
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(va[a], vb[b], acc[a][b]);
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < n && j < n) {
        float v;
        if (metric == 0)
          v = sqrtf(acc[a][b]);
        else
          v = 1.f - acc[a][b] / fmaxf(sqrtf(na[a]) * sqrtf(nb[b]), 1e-30f);
        if (i == j) v = 0.f;
        D[i * n + j] = v;
      }
    }
}

// R[A][c] = sum_{a in members(A)} S[a][c]        (CSR: offs[A]..offs[A+1] into memb)
__global__ void pool_rows_kernel(const float* __restrict__ S, long long tin, const int* __restrict__ offs,
                                 const int* __restrict__ memb, float* __restrict__ R, long long tout) {
  const long long A = blockIdx.y;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= tout || c >= tin) return;
  float s = 0.f;
  for (int k = offs[A]; k < offs[A + 1]; ++k) s += S[(long long)memb[k] * tin + c];
  R[A * tin + c] = s;
}

// S'[A][C] = sum_{c in members(C)} R[A][c]
__global__ void pool_cols_kernel(const float* __restrict__ R, long long tin, const int* __restrict__ offs,
                                 const int* __restrict__ memb, float* __restrict__ Sout, long long tout) {
  const long long A = blockIdx.y;
  const long long Cc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= tout || Cc >= tout) return;
  const float* row = R + A * tin;
  float s = 0.f;
  for (int k = offs[Cc]; k < offs[Cc + 1]; ++k) s += row[memb[k]];
  Sout[A * tout + Cc] = s;
}

// nearest cluster of every cluster under average linkage: argmin_{C != A} S[A][C] / (size_A size_C);
// one CTA per row, ties -> smallest index.
__global__ void __launch_bounds__(256) row_argmin_kernel(const float* __restrict__ S, long long t,
                                                         const float* __restrict__ sizes, int* __restrict__ nn,
                                                         float* __restrict__ nnd) {
  const long long A = blockIdx.x;
  const float sa = sizes[A];
  float best = INFINITY;
  int bi = -1;
  for (long long c = threadIdx.x; c < t; c += blockDim.x) {
    if (c == A) continue;
    const float v = S[A * t + c] / (sa * sizes[c]);
    if (v < best || (v == best && (int)c < bi)) {
      best = v;
      bi = (int)c;
    }
  }
  __shared__ float sv[256];
  __shared__ int si[256];
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      const float v = sv[threadIdx.x + s];
      const int i = si[threadIdx.x + s];
      if (i >= 0 && (si[threadIdx.x] < 0 || v < sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x]))) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = i;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    nn[A] = si[0];
    nnd[A] = sv[0];
  }
}

// ---- one agglomeration round entirely on the device -------------------------------------------------------------
// plan: cluster a merges with b = nn[a] iff they are reciprocal nearest neighbours and their average distance is under the
// threshold; the pair is represented by its smaller index.  keep[a] = 1 for clusters that survive as a representative,
// partner[a] = the absorbed cluster (or -1).
__global__ void hac_plan_kernel(const int* __restrict__ nn, const float* __restrict__ nnd, int t, float thr, int strict,
                                int* __restrict__ keep, int* __restrict__ partner) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= t) return;
  const int b = nn[a];
  bool absorbed = false, leads = false;
  if (b >= 0 && nn[b] == a) {
    const int lo = a < b ? a : b;
    const bool ok = strict ? (nnd[lo] < thr) : (nnd[lo] <= thr);    // judged from the representative's row
    leads = ok && a < b;
    absorbed = ok && b < a;
  }
  keep[a] = absorbed ? 0 : 1;
  partner[a] = leads ? b : -1;
}

// members of the new clusters (A = newidx[a] for kept a), their sizes, and the old -> new index map
__global__ void hac_members_kernel(const int* __restrict__ keep, const int* __restrict__ partner, const int* __restrict__ newidx,
                                   const int* __restrict__ nn, const float* __restrict__ sizes, int t, int* __restrict__ m0,
                                   int* __restrict__ m1, float* __restrict__ sizes2, int* __restrict__ map) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= t) return;
  if (keep[a]) {
    const int A = newidx[a];
    const int b = partner[a];
    m0[A] = a;
    m1[A] = b;
    sizes2[A] = sizes[a] + (b >= 0 ? sizes[b] : 0.f);
    map[a] = A;
  } else {
    map[a] = newidx[nn[a]];
  }
}

// S2[A][B] = sum over the (at most 2 x 2) old clusters: rows pooled first, then columns (same order as pool_rows + pool_cols)
__global__ void hac_contract_kernel(const float* __restrict__ S, long long t, const int* __restrict__ m0, const int* __restrict__ m1,
                                    float* __restrict__ S2, long long tout, long long row0) {
  const long long A = row0 + blockIdx.y;
  const long long B = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (A >= tout || B >= tout) return;
  const int a0 = m0[A], a1 = m1[A], b0 = m0[B], b1 = m1[B];
  const float* r0 = S + (long long)a0 * t;
  float v = r0[b0];
  if (a1 >= 0) v += S[(long long)a1 * t + b0];
  if (b1 >= 0) {
    float w = r0[b1];
    if (a1 >= 0) w += S[(long long)a1 * t + b1];
    v += w;
  }
  S2[A * tout + B] = v;
}

__global__ void hac_relabel_kernel(int* __restrict__ cl, long long n, const int* __restrict__ map) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cl[i] = map[cl[i]];
}

}  // namespace

extern "C" int pv_hac_plan(const int* nn, const float* nnd, int64_t t, float threshold, int strict, int* keep, int* partner,
                           void* stream) {
  PV_REQUIRE(nn && nnd && keep && partner, "pv_hac_plan: null argument");
  if (t == 0) return PV_OK;
  hac_plan_kernel<<<(unsigned)((t + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(nn, nnd, (int)t, threshold, strict,
                                                                                            keep, partner);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

/* newidx = exclusive prefix sum of keep (caller); writes members m0/m1 [tout], sizes2 [tout], map [t] */
extern "C" int pv_hac_members(const int* keep, const int* partner, const int* newidx, const int* nn, const float* sizes,
                              int64_t t, int* m0, int* m1, float* sizes2, int* map, void* stream) {
  PV_REQUIRE(keep && partner && newidx && nn && sizes && m0 && m1 && sizes2 && map, "pv_hac_members: null argument");
  if (t == 0) return PV_OK;
  hac_members_kernel<<<(unsigned)((t + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, partner, newidx, nn, sizes,
                                                                                               (int)t, m0, m1, sizes2, map);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_hac_contract(const float* S, int64_t t, const int* m0, const int* m1, float* S2, int64_t tout, void* stream) {
  PV_REQUIRE(S && m0 && m1 && S2, "pv_hac_contract: null argument");
  if (tout == 0) return PV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long maxy = 65535;
  for (long long a0 = 0; a0 < tout; a0 += maxy) {
    const long long na = (tout - a0 < maxy) ? tout - a0 : maxy;
    hac_contract_kernel<<<dim3((unsigned)((tout + 255) / 256), (unsigned)na), 256, 0, s>>>(S, t, m0, m1, S2, tout, a0);
    g_pv_launches.fetch_add(1);
  }
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_hac_relabel(int* cl, int64_t n, const int* map, void* stream) {
  PV_REQUIRE(cl && map, "pv_hac_relabel: null argument");
  if (n == 0) return PV_OK;
  hac_relabel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cl, n, map);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

// ---- whole-stage entry points (what a non-Python consumer binds) ---------------------------------------------------
namespace {
__global__ void hac_iota_kernel(int* a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int)i;
}
__global__ void hac_count_scan_kernel(const int* __restrict__ keep, int* __restrict__ newidx, long long t, int* __restrict__ total) {
  // single-CTA exclusive scan of keep[0..t) (t <= a few 10^5: one pass of 1024 threads with a running carry)
  __shared__ int s_part[1024];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < t; base += 1024) {
    const long long i = base + threadIdx.x;
    const int v = i < t ? keep[i] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < t) newidx[i] = s_carry + s_part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void hac_first_kernel(const int* __restrict__ cl, long long n, int* __restrict__ first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMin(&first[cl[i]], (int)i);
}
__global__ void hac_label_kernel(const int* __restrict__ cl, const int* __restrict__ first, long long n, int* __restrict__ labels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) labels[i] = first[cl[i]];
}
__global__ void hac_fill_kernel(int* a, long long n, int v) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
__global__ void hac_ones_kernel(float* a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = 1.0f;
}
}  // namespace

extern "C" int pv_gram_dist(const float* X, int64_t n, int dim, int metric, float* D, void* xs_ws, float* norms_ws, int* err_flag,
                            void* stream);
extern "C" int64_t pv_gram_npad(int64_t n);

/* FaceClustering as one call for embeddings that are one-per-track (pyannote/video/face/clustering.py:92-148 with
 * singleton tracks): X f32 [n][128] (device) -> labels i32 [n] (device): label = smallest row index of the row's cluster.
 * Average-linkage agglomeration stopped at `threshold` (strict: stop at >=).  Allocates its work space (2 n^2 floats + the
 * Gram operands) with cudaMalloc and synchronises the stream once per round (it reads how many clusters are left).
 * rounds_out (HOST, may be NULL) receives the number of rounds. */
extern "C" int pv_hac_threshold(const float* X, int64_t n, int dim, int metric, float threshold, int strict, int* labels,
                                int* rounds_out, void* stream) {
  PV_REQUIRE(X && labels, "pv_hac_threshold: null argument");
  PV_REQUIRE(dim == kDim, "pv_hac_threshold: dim=%d (must be 128)", dim);
  if (rounds_out) *rounds_out = 0;
  if (n == 0) return PV_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long npad = pv_gram_npad(n);
  float *bufA = nullptr, *bufB = nullptr, *norms = nullptr, *sz = nullptr, *sz2 = nullptr, *nnd = nullptr;
  void* xs = nullptr;
  int *nn = nullptr, *keep = nullptr, *partner = nullptr, *newidx = nullptr, *m0 = nullptr, *m1 = nullptr, *map = nullptr, *cl = nullptr,
      *first = nullptr, *d_total = nullptr;
  auto cleanup = [&]() {
    cudaFree(bufA); cudaFree(bufB); cudaFree(norms); cudaFree(sz); cudaFree(sz2); cudaFree(nnd); cudaFree(xs); cudaFree(nn);
    cudaFree(keep); cudaFree(partner); cudaFree(newidx); cudaFree(m0); cudaFree(m1); cudaFree(map); cudaFree(cl); cudaFree(first);
    cudaFree(d_total);
  };
#define PV_TRY(expr)                                                          \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) {                                                  \
      pv_set_error("pv_hac_threshold: %s -> %s", #expr, cudaGetErrorString(_e)); \
      cleanup();                                                              \
      return PV_ERR_CUDA;                                                     \
    }                                                                         \
  } while (0)
  PV_TRY(cudaMalloc(&bufA, sizeof(float) * (size_t)n * n));
  PV_TRY(cudaMalloc(&xs, (size_t)3 * npad * kDim * 2));
  PV_TRY(cudaMalloc(&norms, sizeof(float) * npad));
  PV_TRY(cudaMalloc(&sz, sizeof(float) * n));
  PV_TRY(cudaMalloc(&sz2, sizeof(float) * n));
  PV_TRY(cudaMalloc(&nnd, sizeof(float) * n));
  PV_TRY(cudaMalloc(&nn, sizeof(int) * n));
  PV_TRY(cudaMalloc(&keep, sizeof(int) * n));
  PV_TRY(cudaMalloc(&partner, sizeof(int) * n));
  PV_TRY(cudaMalloc(&newidx, sizeof(int) * n));
  PV_TRY(cudaMalloc(&m0, sizeof(int) * n));
  PV_TRY(cudaMalloc(&m1, sizeof(int) * n));
  PV_TRY(cudaMalloc(&map, sizeof(int) * n));
  PV_TRY(cudaMalloc(&cl, sizeof(int) * n));
  PV_TRY(cudaMalloc(&first, sizeof(int) * n));
  PV_TRY(cudaMalloc(&d_total, sizeof(int)));
  const unsigned gb = (unsigned)((n + 255) / 256);
  int rc = pv_gram_dist(X, n, dim, metric, bufA, xs, norms, nullptr, stream);
  if (rc != PV_OK) { cleanup(); return rc; }
  hac_iota_kernel<<<gb, 256, 0, st>>>(cl, n);
  hac_ones_kernel<<<gb, 256, 0, st>>>(sz, n);
  float* S = bufA;
  float* other = nullptr;      // allocated at the first contraction (its size is known then)
  long long t = n;
  int rounds = 0;
  while (t > 1) {
    rc = pv_row_argmin(S, t, sz, nn, nnd, stream);
    if (rc == PV_OK) rc = pv_hac_plan(nn, nnd, t, threshold, strict, keep, partner, stream);
    if (rc != PV_OK) { cleanup(); return rc; }
    hac_count_scan_kernel<<<1, 1024, 0, st>>>(keep, newidx, t, d_total);
    int tout = 0;
    PV_TRY(cudaMemcpyAsync(&tout, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
    PV_TRY(cudaStreamSynchronize(st));
    if (tout == t) break;
    if (!other) {
      PV_TRY(cudaMalloc(&bufB, sizeof(float) * (size_t)tout * tout));
      other = bufB;
    }
    rc = pv_hac_members(keep, partner, newidx, nn, sz, t, m0, m1, sz2, map, stream);
    if (rc == PV_OK) rc = pv_hac_contract(S, t, m0, m1, other, tout, stream);
    if (rc == PV_OK) rc = pv_hac_relabel(cl, n, map, stream);
    if (rc != PV_OK) { cleanup(); return rc; }
    float* tmp = S; S = other; other = tmp;
    float* ts = sz; sz = sz2; sz2 = ts;
    t = tout;
    ++rounds;
  }
  hac_fill_kernel<<<gb, 256, 0, st>>>(first, n, 0x7fffffff);
  hac_first_kernel<<<gb, 256, 0, st>>>(cl, n, first);
  hac_label_kernel<<<gb, 256, 0, st>>>(cl, first, n, labels);
  PV_TRY(cudaStreamSynchronize(st));
  PV_TRY(cudaGetLastError());
#undef PV_TRY
  if (rounds_out) *rounds_out = rounds;
  cleanup();
  return PV_OK;
}

/* TrackingByDetection._match (pyannote/video/tracking.py:129-134): area of the intersection of two (l,t,r,b) rectangles in
 * dlib drectangle arithmetic (width = r - l), or 0 unless it covers at least `ratio` of BOTH rectangles.  Host function. */
extern "C" double pv_rect_overlap(const double* a, const double* b, double ratio) {
  const double l = a[0] > b[0] ? a[0] : b[0], t = a[1] > b[1] ? a[1] : b[1];
  const double r = a[2] < b[2] ? a[2] : b[2], bo = a[3] < b[3] ? a[3] : b[3];
  if (r <= l || bo <= t) return 0.0;
  const double inter = (r - l) * (bo - t);
  const double a1 = (a[2] - a[0]) * (a[3] - a[1]), a2 = (b[2] - b[0]) * (b[3] - b[1]);
  if (inter >= ratio * a1 && inter >= ratio * a2) return inter;
  return 0.0;
}

extern "C" int pv_pdist(const float* X, int64_t n, int dim, int metric, float* D, void* stream) {
  PV_REQUIRE(X && D, "pv_pdist: null argument");
  PV_REQUIRE(dim == kDim, "pv_pdist: dim=%d (must be 128)", dim);
  PV_REQUIRE(metric == 0 || metric == 1, "pv_pdist: metric=%d", metric);
  if (n == 0) return PV_OK;
  const unsigned g = (unsigned)((n + kTile - 1) / kTile);
  pdist_kernel<<<dim3(g, g), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, n, D, metric);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_pool_rows(const float* S, int64_t tin, const int* offs, const int* memb, float* R, int64_t tout,
                            void* stream) {
  PV_REQUIRE(S && offs && memb && R, "pv_pool_rows: null argument");
  if (tout == 0 || tin == 0) return PV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long maxy = 65535;
  for (long long a0 = 0; a0 < tout; a0 += maxy) {  // gridDim.y limit
    const long long na = (tout - a0 < maxy) ? tout - a0 : maxy;
    pool_rows_kernel<<<dim3((unsigned)((tin + 255) / 256), (unsigned)na), 256, 0, s>>>(S, tin, offs + a0, memb,
                                                                                      R + a0 * tin, na);
    g_pv_launches.fetch_add(1);
  }
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_pool_cols(const float* R, int64_t tin, const int* offs, const int* memb, float* Sout, int64_t tout,
                            void* stream) {
  PV_REQUIRE(R && offs && memb && Sout, "pv_pool_cols: null argument");
  if (tout == 0 || tin == 0) return PV_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long maxy = 65535;
  for (long long a0 = 0; a0 < tout; a0 += maxy) {
    const long long na = (tout - a0 < maxy) ? tout - a0 : maxy;
    // rows a0..a0+na of the output; the column CSR is the full one
    pool_cols_kernel<<<dim3((unsigned)((tout + 255) / 256), (unsigned)na), 256, 0, s>>>(R + a0 * tin, tin, offs, memb,
                                                                                       Sout + a0 * tout, tout);
    g_pv_launches.fetch_add(1);
  }
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}

extern "C" int pv_row_argmin(const float* S, int64_t t, const float* sizes, int* nn, float* nnd, void* stream) {
  PV_REQUIRE(S && sizes && nn && nnd, "pv_row_argmin: null argument");
  if (t == 0) return PV_OK;
  row_argmin_kernel<<<(unsigned)t, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, t, sizes, nn, nnd);
  g_pv_launches.fetch_add(1);
  PV_CUDA_CHECK(cudaGetLastError());
  return PV_OK;
}
