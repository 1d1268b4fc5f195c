/* pv_b200.h — C ABI of libpvb200.so: the B200-native (sm_100a) kernels behind the
 * pyannote.video face path (detect -> landmarks -> embed -> track -> cluster).
 *
 * The reference (pyannote-video 1.6.3) is pure Python that binds dlib 19.12 through dlib's
 * Python module; each entry point below names the dlib call site it replaces
 * (paths are relative to the reference checkout):
 *
 *   pyramid / srgemm / detect_decode  <- dlib.get_frontal_face_detector()(rgb, 1)
 *                                        pyannote/video/face/face.py:54,66
 *                                        (as dlib's CNN/MMOD detector, per BASELINE.json north_star)
 *   ert_forward                       <- dlib.shape_predictor.__call__   pyannote/video/face/face.py:58,70
 *   chip_extract / srgemm / embed_*   <- dlib.face_recognition_model_v1.compute_face_descriptor
 *                                        pyannote/video/face/face.py:62,74-75
 *   tracker_*                         <- dlib.correlation_tracker.start_track/update/get_position
 *                                        pyannote/video/tracking.py:203,231,250-251
 *   rect_overlap                      <- TrackingByDetection._match      pyannote/video/tracking.py:129-134
 *   pdist_* / hac_*                   <- scipy pdist + pyannote.algorithms HAC
 *                                        pyannote/video/face/clustering.py:92-119,138-148
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = PV_ERR_*; pv_last_error() returns a message.
 *   - every data pointer is a DEVICE pointer owned by the caller (a torch.Tensor.data_ptr())
 *     unless the parameter name says host; outputs are pre-allocated by the caller.
 *   - the library owns only opaque handles from pv_*_create / pv_*_destroy.
 *   - the last argument of every launch is a cudaStream_t passed as void*; launches never
 *     synchronise the host.
 *   - no function has a CPU fallback: without a CUDA device they return PV_ERR_CUDA.
 */
#ifndef PV_B200_H
#define PV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_OK 0
#define PV_ERR_INVALID (-1)
#define PV_ERR_CUDA (-2)
#define PV_ERR_UNSUPPORTED (-3)
#define PV_ERR_DEVICE_TIMEOUT (-4)

const char* pv_last_error(void);
int pv_version(void);
/* number of kernels this library has launched since load (bench.py: gpu_launches) */
int64_t pv_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * srgemm: "shifted-row GEMM", the one tensor-core kernel behind every convolution.
 *
 *   D[q, 0:N] = sum_t  X[q + off_t, col_t : col_t + w_t] * W_t[w_t, N]        q in [0, Q)
 *   Y[dst(q)] = act( scale * D[q] + shift (+ R[res(q)]) )                      if valid(q)
 *
 * X is a row-major bf16 matrix (an activation tensor in one of the row layouts of DESIGN.md);
 * a convolution is a list of taps (row offset, column segment).  128 consecutive rows q form
 * one tile = one 128xN fp32 accumulator in TMEM; operands are staged by TMA into 32/64/128-byte
 * swizzled shared memory and consumed by tcgen05.mma (cta_group::1, kind::f16, bf16 x bf16).
 * ------------------------------------------------------------------------------------------ */
#define PV_SR_MAX_TAPS 9
#define PV_SR_MAX_STAGES 192

typedef struct PvSrStage {
  int32_t a_row_off;   /* first row of the A slab, relative to the tile's first row q0        */
  int32_t b_row;       /* first row of this stage's weights in the packed matrix of class cls */
  int16_t a_col;       /* first column (element) of the segment inside X                      */
  int16_t cls;         /* segment-width class (0 or 1)                                        */
  int16_t n_taps;      /* taps sharing this slab (1..PV_SR_MAX_TAPS)                          */
  int16_t use_tail;    /* 1: slab = 128 + tail_rows rows, 0: 128 rows                         */
  int16_t tap_rel[PV_SR_MAX_TAPS + 1]; /* row offset of each tap inside the slab              */
} PvSrStage;

/* maps an output grid position (n, y, x) to a row of a destination / residual matrix */
typedef struct PvRowMap {
  int32_t kind;        /* 0: padded NHWC  row = n*img + (y+py)*w + (x+px)
                          1: parity split row = plane*plane_rows + n*img + ((y+py)>>1)*w + ((x+px)>>1),
                             plane = ((y+py)&1)*2 + ((x+px)&1)                               */
  int32_t cols;        /* row length in elements                                              */
  int32_t w;           /* grid width of the destination                                       */
  int32_t py, px;
  int64_t img;         /* rows per image                                                      */
  int64_t plane_rows;  /* rows per parity plane (kind 1)                                      */
} PvRowMap;

typedef struct PvSrgemmDesc {
  /* operands (device pointers, bound for the lifetime of the plan) */
  const void* x;        /* bf16 [x_rows, x_cols]                                               */
  int64_t x_rows;
  int32_t x_cols;
  int32_t n_out;        /* N: multiple of 16, 16..256                                          */
  int32_t n_classes;    /* 1 or 2                                                              */
  int32_t class_width[2]; /* 16, 32 or 64 elements                                             */
  const void* w_packed[2]; /* bf16 [w_rows[c], class_width[c]]                                 */
  int64_t w_rows[2];
  int32_t tail_rows;    /* multiple of 8, 0..128                                               */
  int32_t n_stages;
  const PvSrStage* stages; /* HOST pointer, copied                                             */
  const float* scale;   /* [N] */
  const float* shift;   /* [N] */
  /* output grid */
  int32_t hq, wq;       /* grid rows / cols per image; q = (n*hq + y)*wq + x                   */
  int32_t oh, ow;       /* valid output extent: y < oh && x < ow                               */
  int32_t relu;
  int32_t out_mode;     /* 0: bf16 rows via dst map; 2: fp32 channel 0 only, row = dst map     */
  void* out;
  PvRowMap dst;
  const void* resid;    /* bf16 or NULL */
  PvRowMap res;
  int32_t desc_mode;    /* 0: UMMA descriptor base_offset = 0; 1: base_offset = (addr>>7)&7    */
  int32_t max_ctas;     /* 0 = number of SMs                                                   */
} PvSrgemmDesc;

int pv_srgemm_create(const PvSrgemmDesc* desc, void** out_handle);
int pv_srgemm_run(void* handle, int64_t q_rows, void* stream);
int pv_srgemm_destroy(void* handle);
/* reads (and clears) the device-side error flag of a plan; 0 = none. Synchronises the stream. */
int pv_srgemm_check(void* handle, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PV_B200_H */
