"""Thin Python wrappers over the non-conv kernels (ERT landmarks, chip extraction).

  ShapePredictor.predict  <- dlib.shape_predictor.__call__           pyannote/video/face/face.py:70
  ChipExtractor.extract   <- get_face_chip_details/extract_image_chip (inside compute_face_descriptor,
                             pyannote/video/face/face.py:74-75)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import weights as W


def _dev(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device).contiguous()


class ShapePredictor:
    def __init__(self, model, device):
        if model.get("kind") != "ert_shape_predictor":
            raise RuntimeError("ShapePredictor: not a shape-predictor model")
        self.dev = device
        self.initial_shape = _dev(model["initial_shape"], device, torch.float32)
        self.anchor_idx = _dev(model["anchor_idx"], device, torch.int32)
        self.deltas = _dev(model["deltas"], device, torch.float32)
        self.split_idx1 = _dev(model["split_idx1"], device, torch.int32)
        self.split_idx2 = _dev(model["split_idx2"], device, torch.int32)
        self.split_thresh = _dev(model["split_thresh"], device, torch.float32)
        self.leaf_values = _dev(model["leaf_values"], device, torch.float32)
        S, T, n_split = model["split_thresh"].shape
        if n_split != 15 or model["leaf_values"].shape[2:] != (16, 2 * W.ERT_POINTS):
            raise RuntimeError("ShapePredictor: expected depth-4 trees and 68 points")
        self.num_parts = W.ERT_POINTS
        h = C.c_void_p()
        _lib.check(_lib.lib().pv_ert_create(
            _lib.ptr(self.initial_shape), _lib.ptr(self.anchor_idx), _lib.ptr(self.deltas), _lib.ptr(self.split_idx1),
            _lib.ptr(self.split_idx2), _lib.ptr(self.split_thresh), _lib.ptr(self.leaf_values), int(S), int(T),
            int(model["anchor_idx"].shape[1]), C.byref(h)), "pv_ert_create")
        self.h = h

    def predict(self, frames, rects, frame_idx, out=None):
        """frames uint8 [F,H,W,3]; rects int32 [M,4] (l,t,r,b); frame_idx int32 [M] -> int32 [M,68,2]"""
        M = rects.shape[0]
        if out is None:
            out = torch.empty(M, W.ERT_POINTS, 2, dtype=torch.int32, device=self.dev)
        assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.shape[-1] == 3
        assert rects.dtype == torch.int32 and frame_idx.dtype == torch.int32
        _lib.check(_lib.lib().pv_ert_forward(self.h, _lib.ptr(frames), frames.shape[1], frames.shape[2], _lib.ptr(rects),
                                             _lib.ptr(frame_idx), M, _lib.ptr(out), _lib.stream_ptr()), "pv_ert_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().pv_ert_destroy(self.h)
                self.h = None
        except Exception:
            pass


CHIP_POINTS = list(W.CHIP_POINTS)


def chip_from_points(size=W.EMB_CHIP, padding=W.EMB_CHIP_PADDING):
    """chip-space alignment targets (dlib get_face_chip_details), float32 [68,2]."""
    f32 = np.float32
    mean = W.chip_mean_face()
    frm = np.zeros((68, 2), f32)
    pad = f32(padding)
    scale = f32(2.0) * pad + f32(1.0)
    for i in CHIP_POINTS:
        frm[i, 0] = ((pad + mean[i - 17, 0]) / scale) * f32(size)
        frm[i, 1] = ((pad + mean[i - 17, 1]) / scale) * f32(size)
    return frm


class ChipExtractor:
    def __init__(self, device, size=W.EMB_CHIP):
        self.dev = device
        self.size = size
        self.from_pts = _dev(chip_from_points(size), device, torch.float32)
        self.pt_idx = _dev(np.asarray(CHIP_POINTS, np.int32), device, torch.int32)

    def extract(self, frames, parts, frame_idx, out_rgba):
        """frames uint8 [F,H,W,3]; parts int32 [M,68,2]; out_rgba uint8 [>=M,size,size,4]."""
        M = parts.shape[0]
        assert parts.dtype == torch.int32 and parts.is_contiguous() and frame_idx.dtype == torch.int32
        _lib.check(_lib.lib().pv_chip_extract(_lib.ptr(frames), frames.shape[1], frames.shape[2], _lib.ptr(parts),
                                              _lib.ptr(frame_idx), M, _lib.ptr(self.from_pts), _lib.ptr(self.pt_idx),
                                              len(CHIP_POINTS), self.size, _lib.ptr(out_rgba), _lib.stream_ptr()),
                   "pv_chip_extract")
        return out_rgba[:M]
