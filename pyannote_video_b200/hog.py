"""Host side of csrc/hog.cu: dlib's HOG frontal face detector (`dlib.get_frontal_face_detector()(rgb, 1)`, the detector
the reference really calls, pyannote/video/face/face.py:54,66) on the tiled image pyramid the CNN detector builds.

    frames -> pyramid plane (pv_resize_bilinear / pv_pyramid_tail, shared with nets.DetectorNet)
           -> pv_hog_features: gradients, cell histograms (one warp per cell), 31 Felzenszwalb features -> feature plane
           -> ONE tcgen05 convolution (csrc/rsconv.cu, 10 x 10 x 32 -> 16 filters) = every filter at every position
           -> pv_hog_decode: threshold, sort, boxes (fhog_to_image, rect_up, rect_down), greedy NMS

Same interface as nets.DetectorNet (`detect(frames) -> (boxes, scores, counts)`), so `Face(detector=<hog model>)` works.
dlib's five trained filters are compiled into dlib and not obtainable here: models come from
`weights.make_hog_detector` (seeded) or from a dict / .npz with "filters" [D,31,10,10] and "thresholds" [D].
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from . import weights as W
from .detconv import RsConv, even
from .nets import DetectorNet
from .pyrgeom import pyramid_geometry

GAP = 10        # zero cells between feature tiles = the filter size: a window never covers two levels


def hog_geometry(geo):
    """feature-plane layout for the levels of a pyrgeom.PyramidGeometry: usable levels (both sides >= 80 px) stacked
    vertically, GAP zero cells around each.  Returns (PvHogGeo, [level index per entry])."""
    g = _lib.PvHogGeo()
    g.Hp, g.Wp = geo.plane_h, geo.plane_w
    n = 0
    px = cells = feat = 0
    fy = GAP
    fw_max = 0
    used = []
    for lv, ((x0, y0, w, h)) in enumerate(geo.rects):
        if w < W.HOG_CELL * W.HOG_FILTER or h < W.HOG_CELL * W.HOG_FILTER:
            continue
        cx, cy = int(np.float32(w) / np.float32(W.HOG_CELL) + np.float32(0.5)), int(np.float32(h) / np.float32(W.HOG_CELL) + np.float32(0.5))
        if cx - 2 < 1 or cy - 2 < 1:
            continue
        if n >= _lib.PV_HOG_MAX_LEVELS:
            raise _lib.PvError("hog_geometry: more than %d usable pyramid levels" % _lib.PV_HOG_MAX_LEVELS)
        L = g.lv[n]
        L.x0, L.y0, L.w, L.h = x0, y0, w, h
        L.cx, L.cy = cx, cy
        L.fx0, L.fy0 = GAP, fy
        L.px_off, L.cell_off, L.feat_off = px, cells, feat
        px += w * h
        cells += cx * cy
        feat += (cx - 2) * (cy - 2)
        fy += (cy - 2) + GAP
        fw_max = max(fw_max, cx - 2)
        used.append(lv)
        n += 1
    g.n_levels = n
    g.total_px, g.total_cells, g.total_feat = px, cells, feat
    g.FH, g.FW = fy, fw_max + 2 * GAP
    g.fpitch = even(g.FW)
    return g, used


class HogDetectorNet:
    """Batched HOG detector for frames of one size; drop-in for nets.DetectorNet in face.Face."""

    MAX_CAND = 4096
    MAX_DET = 256
    TAIL_PIXELS = DetectorNet.TAIL_PIXELS
    build_plane = DetectorNet.build_plane
    _init_pyramid = DetectorNet._init_pyramid

    def __init__(self, model, H, W_, upsample, max_batch, device):
        if model.get("kind") != "hog_detector":
            raise RuntimeError("HogDetectorNet: not a HOG detector model")
        filt = np.asarray(model["filters"], np.float32)
        thr = np.asarray(model["thresholds"], np.float32).reshape(-1)
        D = filt.shape[0]
        if filt.shape[1:] != (31, W.HOG_FILTER, W.HOG_FILTER) or thr.shape[0] != D or not (1 <= D <= 8):
            raise RuntimeError("HogDetectorNet: filters must be [D<=8,31,10,10] with D thresholds")
        self.model, self.D = model, D
        self.B = B = int(max_batch)
        self.H, self.W, self.upsample, self.dev = H, W_, int(upsample), device
        self.geo = geo = pyramid_geometry(H, W_, upsample)
        Hp, Wp = geo.plane_h, geo.plane_w
        self.plane = torch.zeros(B, Hp, Wp, 4, dtype=torch.uint8, device=device)
        self._init_pyramid()
        self.hgeo, self.levels = hog_geometry(geo)
        g = self.hgeo
        self.ori = torch.zeros(B, Hp, Wp, dtype=torch.uint8, device=device)
        self.mag = torch.zeros(B, Hp, Wp, dtype=torch.float32, device=device)
        self.hist = torch.zeros(B, max(g.total_cells, 1), 18, dtype=torch.float32, device=device)
        self.nrm = torch.zeros(B, max(g.total_cells, 1), dtype=torch.float32, device=device)
        self.feat = torch.zeros(B, g.FH, g.fpitch, 32, dtype=torch.bfloat16, device=device)     # gaps stay zero
        uv = [math.cos(o * math.pi / 9) for o in range(9)] + [math.sin(o * math.pi / 9) for o in range(9)]
        self.uv = torch.tensor(np.asarray(uv, np.float64).astype(np.float32), device=device)
        # filters as output channels of one 10 x 10 convolution over the feature plane
        w = torch.zeros(16, 32, W.HOG_FILTER, W.HOG_FILTER)
        w[:D, :31] = torch.from_numpy(filt)
        self.conv = RsConv(self.feat, g.FH, g.FW, w, 1, torch.ones(16), torch.zeros(16), False, 32, 16, out_f32=True)
        self.scores = self.conv.out                          # fp32 [B, FH + 1, even(FW + 1), 16]
        self.thr = torch.from_numpy(thr).to(device)
        self.counts = torch.zeros(B, dtype=torch.int32, device=device)
        self.cand_score = torch.zeros(B, self.MAX_CAND, dtype=torch.float32, device=device)
        self.cand_code = torch.zeros(B, self.MAX_CAND, dtype=torch.int32, device=device)
        self.out_boxes = torch.zeros(B, self.MAX_DET, 4, dtype=torch.int32, device=device)
        self.out_scores = torch.zeros(B, self.MAX_DET, dtype=torch.float32, device=device)
        self.out_which = torch.zeros(B, self.MAX_DET, dtype=torch.int32, device=device)
        self.out_counts = torch.zeros(B, dtype=torch.int32, device=device)

    def features(self, M):
        _lib.check(_lib.lib().pv_hog_features(_lib.ptr(self.plane), M, C.byref(self.hgeo), _lib.ptr(self.uv), _lib.ptr(self.ori),
                                              _lib.ptr(self.mag), _lib.ptr(self.hist), _lib.ptr(self.nrm), _lib.ptr(self.feat),
                                              _lib.stream_ptr()), "pv_hog_features")

    def forward_scores(self, M):
        self.features(M)
        self.conv.run(M)
        return self.scores[:M]

    def decode(self, M, threshold=None):
        thr = self.thr if threshold is None else torch.maximum(self.thr, torch.full_like(self.thr, float(threshold)))
        m = self.model
        _lib.check(_lib.lib().pv_hog_decode(_lib.ptr(self.scores), M, self.scores.shape[1], self.scores.shape[2], C.byref(self.hgeo),
                                            _lib.ptr(thr), self.D, int(self.upsample),
                                            C.c_double(float(m.get("iou_thresh", W.HOG_NMS_IOU))),
                                            C.c_double(float(m.get("covered_thresh", W.HOG_NMS_COVERED))), self.MAX_CAND, self.MAX_DET,
                                            _lib.ptr(self.counts), _lib.ptr(self.cand_score), _lib.ptr(self.cand_code),
                                            _lib.ptr(self.out_boxes), _lib.ptr(self.out_scores), _lib.ptr(self.out_which),
                                            _lib.ptr(self.out_counts), _lib.stream_ptr()), "pv_hog_decode")
        return self.out_boxes[:M], self.out_scores[:M], self.out_counts[:M]

    def detect(self, frames):
        """frames uint8 [M,H,W,3] (device).  Returns (boxes int32 [M,MAX_DET,4], scores, counts) on device."""
        M = frames.shape[0]
        assert M <= self.B and frames.shape[1:] == (self.H, self.W, 3) and frames.dtype == torch.uint8
        self.build_plane(frames.contiguous(), M)
        self.forward_scores(M)
        return self.decode(M)

    def level_scores(self, b, k):
        """fp32 [D, cy-2, cx-2] score maps of usable level k of image b (test access; oracle/hog.py score_maps order)"""
        L = self.hgeo.lv[k]
        hy, hx = L.cy - 2, L.cx - 2
        return self.scores[b, L.fy0 + 1:L.fy0 + 1 + hy, L.fx0 + 1:L.fx0 + 1 + hx, :self.D].permute(2, 0, 1)

    def level_features(self, b, k):
        """bf16 [cy-2, cx-2, 31] features of usable level k of image b"""
        L = self.hgeo.lv[k]
        return self.feat[b, L.fy0:L.fy0 + L.cy - 2, L.fx0:L.fx0 + L.cx - 2, :31]

    def check(self):
        self.conv.check()
