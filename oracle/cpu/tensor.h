// ORACLE (test infrastructure) — NHWC float32 activation tensor with a zero halo (see conv.cpp).
#ifndef PVC_TENSOR_H
#define PVC_TENSOR_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <memory>

namespace pvc {

typedef float v16 __attribute__((vector_size(64)));
typedef float v16u __attribute__((vector_size(64), aligned(4)));   // unaligned loads

static inline v16 v16_zero() { return v16{} ; }
static inline v16 v16_max0(v16 v) {
  const v16 z = v16{};
  return v > z ? v : z;
}

struct Tensor {
  int N = 0, H = 0, W = 0, C = 0, Cp = 0, halo = 0;
  long pitch_px = 0;   // pixels per stored row (W + 2 halo + slack)
  long rows = 0;       // stored rows per image (H + 2 halo + slack)
  std::shared_ptr<char> raw_store;
  float* base = nullptr;   // &(n=0, y=-halo, x=-halo, c=0)
  long c_off = 0;          // channel offset of a view

  Tensor() {}
  Tensor(int n, int h, int w, int c, int halo_, int cp = 0) : N(n), H(h), W(w), C(c), halo(halo_) {
    Cp = cp ? cp : (c + 15) / 16 * 16;
    pitch_px = (long)W + 2 * halo + 48;     // slack: register tiles may read past the last output's window
    rows = (long)H + 2 * halo + 2;
    const size_t n_float = (size_t)N * rows * pitch_px * Cp + 64;
    // calloc: large blocks come from fresh zero pages (no memset pass; first touch happens in the worker threads)
    char* raw = static_cast<char*>(calloc(n_float * sizeof(float) + 64, 1));
    raw_store = std::shared_ptr<char>(raw, free);
    base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 63) & ~(uintptr_t)63);
  }
  inline long row_stride() const { return pitch_px * Cp; }
  inline float* at(int n, int y, int x) const {
    return base + (((long)n * rows + (y + halo)) * pitch_px + (x + halo)) * Cp + c_off;
  }
  Tensor channel_view(int c0, int c) const {
    Tensor v = *this;
    v.c_off = c_off + c0;
    v.C = c;
    return v;   // NB: Cp stays the pixel stride of the parent; writers store 16-float blocks below v.C
  }
};

}  // namespace pvc
#endif
