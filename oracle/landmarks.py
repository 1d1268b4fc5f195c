"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement (numpy float32, unfused ops, fixed summation order) of
  * dlib.shape_predictor.__call__(rgb, rect)              — reference call site
    pyannote/video/face/face.py:70 (68-point ERT cascade, Kazemi & Sullivan CVPR'14)
  * get_face_chip_details + extract_image_chip             — first half of
    face_recognition_model_v1.compute_face_descriptor, pyannote/video/face/face.py:74-75
as recalled in SURVEY.md App. A.3 / A.4 (dlib 19.12 absent: parity unpinned).

Stated deviations from dlib, shared with the CUDA path: the 2-D similarity fit uses the closed
form (equal to Umeyama's optimum for proper rotations) evaluated in float32; chips are sampled with
plain bilinear interpolation (no pyramid_down pre-shrink for large faces).  The 51 mean-face constants
and the landmark exclusion list (eyebrows 17..26, lower lip 55..59 and 65..67) are dlib's, as recalled
(oracle/constants.py).
"""
import numpy as np

from . import constants as W

f32 = np.float32


def similarity_fit(a, b, idx=None):
    """Least-squares similarity b ~ M a + t.  a,b: float32 [M,P,2].  Sequential sums over points in
    index order.  Returns (m00, m01, m10, m11, tx, ty) each float32 [M]."""
    if idx is None:
        idx = range(a.shape[1])
    idx = list(idx)
    n = f32(len(idx))
    Mn = a.shape[0]
    z = lambda: np.zeros(Mn, f32)
    sax, say, sbx, sby = z(), z(), z(), z()
    for i in idx:
        sax = sax + a[:, i, 0]
        say = say + a[:, i, 1]
        sbx = sbx + b[:, i, 0]
        sby = sby + b[:, i, 1]
    max_, may, mbx, mby = sax / n, say / n, sbx / n, sby / n
    A, Bc, den = z(), z(), z()
    for i in idx:
        acx = a[:, i, 0] - max_
        acy = a[:, i, 1] - may
        bcx = b[:, i, 0] - mbx
        bcy = b[:, i, 1] - mby
        A = A + ((acx * bcx) + (acy * bcy))
        Bc = Bc + ((acx * bcy) - (acy * bcx))
        den = den + ((acx * acx) + (acy * acy))
    m00 = A / den
    m10 = Bc / den
    m01 = -m10
    m11 = m00
    tx = mbx - ((m00 * max_) + (m01 * may))
    ty = mby - ((m10 * max_) + (m11 * may))
    return m00, m01, m10, m11, tx, ty


def ert_predict(model, rgb, rects):
    """rgb: uint8 [H,W,3]; rects: int [M,4] (l,t,r,b).  Returns int64 [M,68,2] (x,y) landmark parts."""
    rects = np.asarray(rects, np.int64).reshape(-1, 4)
    Mn = rects.shape[0]
    H, Wd, _ = rgb.shape
    gray = ((rgb[..., 0].astype(np.uint32) + rgb[..., 1] + rgb[..., 2]) // 3).astype(np.uint8)
    init = model["initial_shape"].astype(f32).reshape(-1, 2)
    P = init.shape[0]
    cur = np.broadcast_to(init, (Mn, P, 2)).astype(f32).copy()
    l = rects[:, 0].astype(f32)
    t = rects[:, 1].astype(f32)
    wr = (rects[:, 2] - rects[:, 0]).astype(f32)
    hr = (rects[:, 3] - rects[:, 1]).astype(f32)
    a = np.broadcast_to(init, (Mn, P, 2)).astype(f32)
    stages = model["split_thresh"].shape[0]
    for s in range(stages):
        m00, m01, m10, m11, _, _ = similarity_fit(a, cur)
        anchor = model["anchor_idx"][s]
        d = model["deltas"][s].astype(f32)                       # [pool,2]
        dx = ((m00[:, None] * d[None, :, 0]) + (m01[:, None] * d[None, :, 1])) + cur[:, anchor, 0]
        dy = ((m10[:, None] * d[None, :, 0]) + (m11[:, None] * d[None, :, 1])) + cur[:, anchor, 1]
        px = l[:, None] + (dx * wr[:, None])
        py = t[:, None] + (dy * hr[:, None])
        ix = np.floor(px + f32(0.5)).astype(np.int64)
        iy = np.floor(py + f32(0.5)).astype(np.int64)
        inside = (ix >= 0) & (ix < Wd) & (iy >= 0) & (iy < H)
        feat = np.where(inside, gray[np.clip(iy, 0, H - 1), np.clip(ix, 0, Wd - 1)], 0).astype(f32)  # [M,pool]
        i1, i2, th = model["split_idx1"][s], model["split_idx2"][s], model["split_thresh"][s]
        leaves = model["leaf_values"][s]                         # [trees,16,136]
        n_split = i1.shape[1]
        for tr in range(i1.shape[0]):
            node = np.zeros(Mn, np.int64)
            while True:
                act = node < n_split
                if not act.any():
                    break
                nn = np.minimum(node, n_split - 1)
                diff = feat[np.arange(Mn), i1[tr][nn]] - feat[np.arange(Mn), i2[tr][nn]]
                go_left = diff > th[tr][nn]
                nxt = np.where(go_left, 2 * nn + 1, 2 * nn + 2)
                node = np.where(act, nxt, node)
            leaf = node - n_split
            cur = cur + leaves[tr][leaf].reshape(Mn, P, 2).astype(f32)
    x = l[:, None] + (cur[:, :, 0] * wr[:, None])
    y = t[:, None] + (cur[:, :, 1] * hr[:, None])
    out = np.stack([np.floor(x + f32(0.5)), np.floor(y + f32(0.5))], axis=2).astype(np.int64)
    return out


CHIP_POINTS = W.chip_points()


def chip_from_points(size=W.EMB_CHIP, padding=W.EMB_CHIP_PADDING):
    """chip-space alignment targets of get_face_chip_details: ((padding + mean) / (2 padding + 1)) * size for the
    landmarks that take part (others stay 0), float32 [68,2]"""
    mean = W.mean_face()                                         # [51,2] for landmarks 17..67
    frm = np.zeros((68, 2), f32)
    pad = f32(padding)
    scale = f32(2.0) * pad + f32(1.0)
    for i in CHIP_POINTS:
        frm[i, 0] = ((pad + mean[i - 17, 0]) / scale) * f32(size)
        frm[i, 1] = ((pad + mean[i - 17, 1]) / scale) * f32(size)
    return frm


def chip_transform(parts, size=W.EMB_CHIP, padding=W.EMB_CHIP_PADDING):
    """get_face_chip_details: similarity mapping chip pixel coords -> image coords.  parts int [M,68,2]."""
    parts = np.asarray(parts)
    Mn = parts.shape[0]
    frm = np.broadcast_to(chip_from_points(size, padding), (Mn, 68, 2)).astype(f32)
    to = parts.astype(f32)
    return similarity_fit(frm, to, CHIP_POINTS)


def extract_chips(rgb, parts, size=W.EMB_CHIP):
    """rgb uint8 [H,W,3], parts int [M,68,2] -> uint8 [M,size,size,3] aligned chips."""
    H, Wd, _ = rgb.shape
    m00, m01, m10, m11, tx, ty = chip_transform(parts, size)
    Mn = m00.shape[0]
    c = np.arange(size, dtype=f32)[None, None, :]
    r = np.arange(size, dtype=f32)[None, :, None]
    e = lambda v: v[:, None, None]
    x = ((e(m00) * c) + (e(m01) * r)) + e(tx)
    y = ((e(m10) * c) + (e(m11) * r)) + e(ty)
    left = np.floor(x).astype(np.int64)
    top = np.floor(y).astype(np.int64)
    right, bot = left + 1, top + 1
    ok = (left >= 0) & (top >= 0) & (right < Wd) & (bot < H)
    lr = (x - left.astype(f32)).astype(f32)[..., None]
    tb = (y - top.astype(f32)).astype(f32)[..., None]
    cl = lambda v, hi: np.clip(v, 0, hi)
    s = rgb.astype(f32)
    tl = s[cl(top, H - 1), cl(left, Wd - 1)]
    tr_ = s[cl(top, H - 1), cl(right, Wd - 1)]
    bl = s[cl(bot, H - 1), cl(left, Wd - 1)]
    br = s[cl(bot, H - 1), cl(right, Wd - 1)]
    one = f32(1.0)
    a = ((one - lr) * tl) + (lr * tr_)
    b = ((one - lr) * bl) + (lr * br)
    v = ((one - tb) * a) + (tb * b)
    v = np.clip(np.floor(v + f32(0.5)), 0, 255)
    v = np.where(ok[..., None], v, 0)
    return v.astype(np.uint8)
