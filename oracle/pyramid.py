"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement (numpy float32, unfused ops in a fixed order) of the image-pyramid input stage and
the MMOD decode of dlib's CNN face detector, i.e. what `face_detector_(rgb, 1)` does around the conv
stack (reference call site pyannote/video/face/face.py:66; dlib 19.12 `input_rgb_image_pyramid`,
`pyramid_down<6>`, `pyramid_up`, `loss_mmod::to_label` as recalled in SURVEY.md App. A.1 — parity
unpinned).  Deviations we chose ourselves (dlib's exact tile packing is not recalled) are marked
[OURS]; both this oracle and the CUDA path implement the same definition.
"""
import numpy as np

from . import constants as K
from . import geometry as ogeo

f32 = np.float32


def resize_bilinear_u8(src, oh, ow):
    """dlib resize_image(in,out) with interpolate_bilinear on rgb_pixel images:
    out(r,c) = bilinear(in, r*y_scale, c*x_scale), scale = (in-1)/max(out-1,1); u8 = (int)(v+0.5).
    src: uint8 [H,W,C]."""
    H, W, C = src.shape
    ys = f32(H - 1) / f32(max(oh - 1, 1))
    xs = f32(W - 1) / f32(max(ow - 1, 1))
    y = (np.arange(oh, dtype=f32) * ys).astype(f32)
    x = (np.arange(ow, dtype=f32) * xs).astype(f32)
    top = np.floor(y).astype(np.int64)
    left = np.floor(x).astype(np.int64)
    top = np.minimum(top, H - 1)
    left = np.minimum(left, W - 1)
    bot = np.minimum(top + 1, H - 1)
    right = np.minimum(left + 1, W - 1)
    tb = (y - top.astype(f32)).astype(f32)[:, None, None]
    lr = (x - left.astype(f32)).astype(f32)[None, :, None]
    s = src.astype(f32)
    tl = s[top][:, left]
    tr = s[top][:, right]
    bl = s[bot][:, left]
    br = s[bot][:, right]
    one = f32(1.0)
    a = ((one - lr) * tl).astype(f32) + (lr * tr).astype(f32)
    b = ((one - lr) * bl).astype(f32) + (lr * br).astype(f32)
    v = ((one - tb) * a).astype(f32) + (tb * b).astype(f32)
    v = np.floor(v + f32(0.5))
    return np.clip(v, 0, 255).astype(np.uint8)


def placement_for(H, W, upsample):
    """the tile placement is [OURS] and an INPUT to the oracle: it is taken from the product's packing and
    validated by oracle.geometry.Geometry (sizes, padding, non-overlap) — the only thing the oracle takes
    from the product; every other piece of geometry is restated independently in oracle/geometry.py."""
    from pyannote_video_b200.pyrgeom import pyramid_geometry
    return ogeo.from_product(pyramid_geometry(H, W, upsample))


def build_plane(rgb, upsample=1, geo=None):
    """rgb: uint8 [H,W,3] -> (plane uint8 [Hp,Wp,4] RGBA with A=255 inside pyramid tiles, geometry).
    Level 0 is the (optionally 2x bilinear-upsampled, dlib pyramid_up) image; level i+1 is level i
    resized to floor(5/6) of its size (pyramid_down<6>).  `geo`: an oracle.geometry.Geometry (validated
    placement); by default the product's placement for this frame size, validated."""
    H, W, _ = rgb.shape
    if geo is None:
        geo = placement_for(H, W, upsample)
    plane = np.zeros((geo.plane_h, geo.plane_w, 4), np.uint8)
    cur = rgb
    for lv, (x0, y0, w, h) in enumerate(geo.rects):
        if lv == 0:
            cur = resize_bilinear_u8(rgb, h, w) if upsample else rgb
        else:
            cur = resize_bilinear_u8(cur, h, w)
        plane[y0:y0 + h, x0:x0 + w, :3] = cur
        plane[y0:y0 + h, x0:x0 + w, 3] = 255
    return plane, geo


def normalize_plane(plane_rgba):
    """RGBA u8 plane -> float32 [3,Hp,Wp]: (v-mean)/256 inside tiles, 0 in the padding."""
    v = plane_rgba[..., :3].astype(f32)
    out = (v - np.asarray(K.PIXEL_MEAN, f32)) * f32(K.PIXEL_SCALE)
    out = out * (plane_rgba[..., 3:4] > 0)
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def decode(scores, geo, window, adjust_threshold, iou_thresh, covered_thresh, max_candidates=None):
    """loss_mmod::to_label restated.  scores: float32 [OH,OW] -> list of (l,t,r,b,score) in
    original-image pixel coordinates (integers, dlib `rectangle`, inclusive right/bottom).
    `geo` may be an oracle.geometry.Geometry or any object with H/W/upsample/rects/plane_h/plane_w (it is
    re-validated and only the oracle's own geometry functions are used on it)."""
    if not isinstance(geo, ogeo.Geometry):
        geo = ogeo.from_product(geo)
    ys, xs = np.nonzero(scores > f32(adjust_threshold))
    cands = []
    for r, c in zip(ys.tolist(), xs.tolist()):
        px, py = ogeo.cell_to_plane(c), ogeo.cell_to_plane(r)
        lv = geo.level_at(px, py)
        if lv < 0:
            continue  # [OURS] cells whose centre falls into padding do not produce boxes
        box = geo.box_from_plane(lv, px, py, window)
        cands.append((float(scores[r, c]), r * scores.shape[1] + c, box))
    # sort by score descending; ties by cell index ascending ([OURS]: deterministic)
    cands.sort(key=lambda t: (-t[0], t[1]))
    if max_candidates is not None:
        cands = cands[:max_candidates]
    kept = []
    for s, _, b in cands:
        if any(boxes_overlap(b, k[:4], iou_thresh, covered_thresh) for k in kept):
            continue
        kept.append((b[0], b[1], b[2], b[3], s))
    return kept


def rect_area(l, t, r, b):
    """dlib::rectangle: inclusive coordinates; empty if r < l or b < t."""
    if r < l or b < t:
        return 0
    return (r - l + 1) * (b - t + 1)


def boxes_overlap(a, b, iou_thresh, covered_thresh):
    """dlib test_box_overlap::operator(): true if IoU > iou_thresh or either box is covered by the
    other by more than covered_thresh."""
    inner = rect_area(max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3]))
    if inner == 0:
        return False
    aa, ab = rect_area(*a), rect_area(*b)
    outer = aa + ab - inner
    if inner / outer > iou_thresh:
        return True
    if inner / aa > covered_thresh or inner / ab > covered_thresh:
        return True
    return False
