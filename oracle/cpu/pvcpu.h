/* ORACLE (test infrastructure) — C ABI of the C++ CPU restatement of the face path.
 *
 * This is the "restated dlib-style CPU baseline" of BASELINE.md §2 (B-cpu-1 with one thread, B-cpu-N with
 * OpenMP over all host cores): the same algorithm as oracle/NAME.py (which cite the reference lines they follow:
 * pyannote/video/face/face.py:64-76, pyannote/video/tracking.py:203,231,250-251), written as plain C++ so
 * that it can be timed as a CPU implementation rather than as Python.  Byte / integer stages (pyramid plane,
 * decode + NMS, ERT landmarks, chips) are float32 with unfused operations in the numpy oracle's order and are
 * bit-exact with it (tests/test_cpu_ref_cpu.py); the convolutions are fp32 (FMA, blocked direct convolution)
 * and agree within float rounding.  Only tests/, __graft_entry__ and bench.py's cpu legs load this library.
 * Not dlib: dlib 19.12 and its weights are absent from the build environment (parity unpinned).
 */
#ifndef PVCPU_H
#define PVCPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int pvc_set_threads(int n);            /* n <= 0: all cores; returns the thread count in effect */
int pvc_get_threads(void);

/* ---- image pyramid plane (dlib input_rgb_image_pyramid; placement = [L][4] x0,y0,w,h, validated by the caller) */
int pvc_build_plane(const uint8_t* rgb, int H, int W, int upsample, const int* rects, int n_levels, int Hp, int Wp,
                    uint8_t* plane_rgba);

/* ---- CNN (con + affine + relu stacks).  Weights are [cout][cin][k][k] float32, scale/shift per cout. */
void* pvc_net_create(void);
void pvc_net_destroy(void* net);
/* appends a conv layer; returns its index */
int pvc_net_add_conv(void* net, int cout, int cin, int k, int stride, int pad, const float* w, const float* scale,
                     const float* shift, int relu, int round_bf16);
/* MMOD detector: plane RGBA u8 [Hp][Wp][4] -> scores float [OH][OW] (layers 0..n-1 of `net`, the last one has
 * one output channel); mean3/scale = input normalisation; returns 0 and the output size */
int pvc_detector_forward(void* net, const uint8_t* plane_rgba, int Hp, int Wp, const float* mean3, float pixel_scale,
                         int round_bf16, float* scores, int* oh, int* ow);
int pvc_detector_out_size(void* net, int n);
/* ResNet-v1 embedder: net holds conv1 then (a, b) per block in execution order; block_down[i] != 0 for ares_down.
 * chips u8 [M][S][S][3] -> out float [M][128] */
int pvc_embed_forward(void* net, const int* block_down, int n_blocks, const float* fc /*[128][256]*/, const uint8_t* chips,
                      int M, int S, const float* mean3, float pixel_scale, int round_bf16, float* out);

/* ---- MMOD decode: scores [OH][OW] -> boxes int [max][4] + scores; returns the number kept (or -1) */
int pvc_decode(const float* scores, int OH, int OW, const int* rects, const float* fxy, int n_levels, int window,
               int cell_mul, int cell_add, float threshold, double iou_thresh, double covered_thresh, int max_candidates,
               int max_out, int* boxes, float* out_scores);

/* ---- ERT 68-point landmarks (dlib shape_predictor) */
int pvc_ert_predict(const uint8_t* rgb, int H, int W, const float* initial_shape, const int* anchor_idx, const float* deltas,
                    const int* split_idx1, const int* split_idx2, const float* split_thresh, const float* leaf_values,
                    int stages, int trees, int pool, const int* rects, int M, int64_t* out_parts /*[M][68][2]*/);

/* ---- face chips (get_face_chip_details + extract_image_chip) */
int pvc_extract_chips(const uint8_t* rgb, int H, int W, const int64_t* parts, int M, const float* from_pts /*[68][2]*/,
                      const int* pt_idx, int n_pts, int size, uint8_t* chips /*[M][size][size][3]*/);

/* ---- DSST correlation tracker bank (dlib correlation_tracker): double-precision filters like dlib */
void* pvc_trackers_create(int capacity, int use_scale);
void pvc_trackers_destroy(void* bank);
int pvc_trackers_start(void* bank, const uint8_t* rgb, int H, int W, const int* ids, const double* rects, int n);
/* frames: per-track frame pointer index into `frames` [F][H][W][3]; psr out [n] */
int pvc_trackers_update(void* bank, const uint8_t* frames, int H, int W, const int* frame_idx, const int* ids, int n,
                        double* psr);
int pvc_trackers_position(void* bank, int id, double* ltrb);

#ifdef __cplusplus
}
#endif
#endif
