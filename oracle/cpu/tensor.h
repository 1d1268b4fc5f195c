// ORACLE (test infrastructure) — NHWC float32 activation tensor with a zero halo (see conv.cpp).
#ifndef PVC_TENSOR_H
#define PVC_TENSOR_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace pvc {

typedef float v16 __attribute__((vector_size(64)));
typedef float v16u __attribute__((vector_size(64), aligned(4)));   // unaligned loads

static inline v16 v16_zero() { return v16{} ; }
static inline v16 v16_max0(v16 v) {
  const v16 z = v16{};
  return v > z ? v : z;
}

// free-list of activation blocks keyed by size: a frame's tensors are reused by the next frame instead of being
// returned to the OS (fresh zero pages cost a page fault per 4 KB: half of the run time at 16 threads)
struct BlockPool {
  std::mutex mu;
  std::multimap<size_t, char*> free_blocks;
  size_t held = 0;
  static BlockPool& get() {
    static BlockPool p;
    return p;
  }
  char* take(size_t bytes, bool* fresh) {
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.find(bytes);
      if (it != free_blocks.end()) {
        char* p = it->second;
        free_blocks.erase(it);
        held -= bytes;
        *fresh = false;
        return p;
      }
    }
    *fresh = true;
    return static_cast<char*>(calloc(bytes, 1));
  }
  void give(size_t bytes, char* p) {
    std::lock_guard<std::mutex> g(mu);
    if (held + bytes > ((size_t)48 << 30)) {   // keep at most 48 GiB parked (16 concurrent 1080p frames hold ~24 GiB)
      free(p);
      return;
    }
    free_blocks.emplace(bytes, p);
    held += bytes;
  }
};

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
    // pooled blocks; a fresh block is calloc'ed (zero pages, first touch in the worker threads).  A recycled block is
    // cleared only when the tensor has a halo that must read as zero: without a halo every element a kept output
    // depends on is written by the producer, and the slack only feeds register-tile lanes that are discarded.
    const size_t bytes = n_float * sizeof(float) + 64;
    bool fresh = false;
    char* raw = BlockPool::get().take(bytes, &fresh);
    raw_store = std::shared_ptr<char>(raw, [bytes](char* p) { BlockPool::get().give(bytes, p); });
    if (!fresh && halo > 0) memset(raw, 0, bytes);
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
