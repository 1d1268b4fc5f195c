"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of dlib's HOG frontal face detector, the detector the reference really calls:
`dlib.get_frontal_face_detector()(rgb, 1)` (pyannote/video/face/face.py:54,66) =
`object_detector<scan_fhog_pyramid<pyramid_down<6>>>`: 31-channel Felzenszwalb HOG with cell size 8 of every level of a
6:5 image pyramid (after ONE 2x upsampling), a bank of linear filters of 10 x 10 cells x 31 channels slid over each
feature map, detections above a threshold mapped back to image rectangles, sorted by score and pruned by greedy
non-maximum suppression.

Everything dlib-internal here is RECALLED FROM MEMORY of dlib 19.x (`image_transforms/fhog.h`,
`image_processing/scan_fhog_pyramid.h`, `image_transforms/image_pyramid.h`, `image_processing/object_detector.h`): dlib is
absent from the build environment and its five face filters are a base64 blob compiled into the library, so the filters
used by the tests are seeded random ones — parity unpinned.  Recalled constants and formulas, each one
named so it can be audited: cell size 8, filter window 10 x 10 cells (80 x 80 detection window, padding 1), the
zero border of (filter - 1) / 2 cells that `extract_fhog_features` adds, `fhog_to_image`, `pyramid_down<6>::point_up`
= p * 6/5 + 0.3, the detection box = window minus the padding cells, NMS by `test_box_overlap`.
"""
import numpy as np

f32 = np.float32

CELL = 8
FILT = 10                 # filter rows = cols, in cells (80 x 80 window, padding 1)
PADDING = 1               # scan_fhog_pyramid::padding
DET_BOX = FILT - 2 * PADDING          # 8 cells: the reported box leaves the padding cells out
PAD_OFF = (FILT - 1) // 2             # 4: rows / cols of zeros extract_fhog_features puts before the data
EPS = f32(0.0001)
NMS_IOU = 0.5             # [MEMORY] test_box_overlap stored in the detector: iou threshold ...
NMS_COVERED = 1.0         # ... and percent-covered threshold (1.0 = disabled)

_UU = np.cos(np.arange(9) * np.pi / 9).astype(f32)
_VV = np.sin(np.arange(9) * np.pi / 9).astype(f32)


def fhog_hist(img, cell=CELL):
    """gradient orientation histograms of a uint8 [H,W,3] image: float32 [cells_y, cells_x, 18].  Per interior pixel the
    colour channel with the largest squared gradient, its direction snapped to 18 signed orientations, its magnitude
    shared bilinearly between the 4 nearest cells; contributions are added in pixel raster order (float32, unfused)."""
    H, W, _ = img.shape
    cy, cx = int(f32(H) / f32(cell) + f32(0.5)), int(f32(W) / f32(cell) + f32(0.5))
    c = img.astype(f32)
    dx = c[1:-1, 2:] - c[1:-1, :-2]
    dy = c[2:, 1:-1] - c[:-2, 1:-1]
    m2 = (dx * dx) + (dy * dy)
    ch = np.argmax(m2, axis=2)                                   # first maximum: r, g, b
    ii, jj = np.meshgrid(np.arange(H - 2), np.arange(W - 2), indexing="ij")
    gx, gy, v2 = dx[ii, jj, ch], dy[ii, jj, ch], m2[ii, jj, ch]
    mag = np.sqrt(v2).astype(f32)
    best_dot = np.zeros_like(mag)
    best_o = np.zeros(mag.shape, np.int64)
    for o in range(9):
        dot = ((_UU[o] * gx) + (_VV[o] * gy)).astype(f32)
        pos = dot > best_dot
        best_o = np.where(pos, o, best_o)
        best_dot = np.where(pos, dot, best_dot)
        neg = (~pos) & (-dot > best_dot)
        best_o = np.where(neg, o + 9, best_o)
        best_dot = np.where(neg, -dot, best_dot)
    ys = np.arange(1, H - 1, dtype=f32)
    xs = np.arange(1, W - 1, dtype=f32)
    yp = ((ys + f32(0.5)) / f32(cell) - f32(0.5)).astype(f32)
    xp = ((xs + f32(0.5)) / f32(cell) - f32(0.5)).astype(f32)
    iyp, ixp = np.floor(yp).astype(np.int64), np.floor(xp).astype(np.int64)
    vy0, vx0 = (yp - iyp.astype(f32)).astype(f32), (xp - ixp.astype(f32)).astype(f32)
    vy1, vx1 = (f32(1) - vy0).astype(f32), (f32(1) - vx0).astype(f32)
    hist = np.zeros((cy, cx, 18), f32)
    # one unbuffered np.add.at per block of rows, index arrays in PIXEL-MAJOR order (the four target cells of a pixel are
    # distinct), so every (cell, bin) sum sees its pixels in raster order exactly as fhog_hist_reference does
    ROWS = 128
    for r0 in range(0, H - 2, ROWS):
        r1 = min(r0 + ROWS, H - 2)
        m, bo = mag[r0:r1], best_o[r0:r1]
        Ys, Xs, Ws = [], [], []
        for (oy, wy) in ((0, vy1), (1, vy0)):
            for (ox, wx) in ((0, vx1), (1, vx0)):
                Ws.append(((wx[None, :] * wy[r0:r1, None]).astype(f32) * m).astype(f32))
                Ys.append(np.broadcast_to((iyp[r0:r1] + oy)[:, None], m.shape))
                Xs.append(np.broadcast_to((ixp + ox)[None, :], m.shape))
        Y, X, Wt = np.stack(Ys, -1).reshape(-1), np.stack(Xs, -1).reshape(-1), np.stack(Ws, -1).reshape(-1)
        O = np.repeat(bo.reshape(-1), 4)
        ok = (Y >= 0) & (Y < cy) & (X >= 0) & (X < cx)
        np.add.at(hist, (Y[ok], X[ok], O[ok]), Wt[ok])
    return hist


def fhog_hist_reference(img, cell=CELL):
    """the same histograms, written as obvious python loops (small images only): the definition `fhog_hist` vectorises and
    the CUDA kernel follows — every (cell, bin) sum accumulates its pixels in raster order"""
    H, W, _ = img.shape
    cy, cx = int(f32(H) / f32(cell) + f32(0.5)), int(f32(W) / f32(cell) + f32(0.5))
    c = img.astype(f32)
    hist = np.zeros((cy, cx, 18), f32)
    for y in range(1, H - 1):
        for x in range(1, W - 1):
            best = f32(-1)
            gx = gy = f32(0)
            for k in range(3):
                ddx = c[y, x + 1, k] - c[y, x - 1, k]
                ddy = c[y + 1, x, k] - c[y - 1, x, k]
                v = (ddx * ddx) + (ddy * ddy)
                if v > best:
                    best, gx, gy = v, ddx, ddy
            mag = np.sqrt(best)
            bo, bd = 0, f32(0)
            for o in range(9):
                dot = (_UU[o] * gx) + (_VV[o] * gy)
                if dot > bd:
                    bd, bo = dot, o
                elif -dot > bd:
                    bd, bo = -dot, o + 9
            xp = (f32(x) + f32(0.5)) / f32(cell) - f32(0.5)
            yp = (f32(y) + f32(0.5)) / f32(cell) - f32(0.5)
            ixp, iyp = int(np.floor(xp)), int(np.floor(yp))
            vx0, vy0 = xp - f32(ixp), yp - f32(iyp)
            vx1, vy1 = f32(1) - vx0, f32(1) - vy0
            for (oy, wy) in ((0, vy1), (1, vy0)):
                for (ox, wx) in ((0, vx1), (1, vx0)):
                    Y, X = iyp + oy, ixp + ox
                    if 0 <= Y < cy and 0 <= X < cx:
                        hist[Y, X, bo] += (wx * wy) * mag
    return hist


def fhog_features(hist):
    """31-channel features of the interior cells: float32 [cells_y - 2, cells_x - 2, 31] (18 contrast-sensitive, 9
    contrast-insensitive, 4 texture), four 2x2-block normalisations clipped at 0.2"""
    cy, cx, _ = hist.shape
    oy, ox = max(cy - 2, 0), max(cx - 2, 0)
    out = np.zeros((oy, ox, 31), f32)
    if oy == 0 or ox == 0:
        return out
    nrm = np.zeros((cy, cx), f32)
    for o in range(9):
        s_ = (hist[:, :, o] + hist[:, :, o + 9]).astype(f32)
        nrm = (nrm + (s_ * s_)).astype(f32)

    def blk(dy_, dx_):
        y0, x0 = 1 + dy_, 1 + dx_
        return (((nrm[y0:y0 + oy, x0:x0 + ox] + nrm[y0:y0 + oy, x0 + 1:x0 + 1 + ox]) + nrm[y0 + 1:y0 + 1 + oy, x0:x0 + ox])
                + nrm[y0 + 1:y0 + 1 + oy, x0 + 1:x0 + 1 + ox]).astype(f32)

    ns = [(f32(1) / np.sqrt(blk(dy_, dx_) + EPS)).astype(f32) for (dy_, dx_) in ((-1, -1), (-1, 0), (0, -1), (0, 0))]
    h = hist[1:-1, 1:-1]
    t = [np.zeros((oy, ox), f32) for _ in range(4)]
    for o in range(18):
        hk = [np.minimum(h[:, :, o] * ns[k], f32(0.2)).astype(f32) for k in range(4)]
        out[:, :, o] = f32(0.5) * (((hk[0] + hk[1]) + hk[2]) + hk[3])
        for k in range(4):
            t[k] = (t[k] + hk[k]).astype(f32)
    for o in range(9):
        s_ = (h[:, :, o] + h[:, :, o + 9]).astype(f32)
        hk = [np.minimum(s_ * ns[k], f32(0.2)).astype(f32) for k in range(4)]
        out[:, :, 18 + o] = f32(0.5) * (((hk[0] + hk[1]) + hk[2]) + hk[3])
    for k in range(4):
        out[:, :, 27 + k] = f32(0.2357) * t[k]
    return out


def score_maps(feat, filters):
    """sliding-window scores: feat float32 [fy, fx, 31] (the un-padded feature map of one level), filters float32
    [D, 31, 10, 10] -> float32 [D, fy, fx]; entry (y, x) is the filter whose top-left cell sits at data cell
    (y - 4, x - 4): dlib pads the map with PAD_OFF = 4 zero cells before and 5 after the data and evaluates the filter
    wherever it fits the padded map, i.e. exactly fy x fx positions (dlib's saliency image holds it at padded
    (y + 5, x + 5), the filter's centre)."""
    fy, fx, _ = feat.shape
    D = filters.shape[0]
    P = np.zeros((fy + FILT - 1, fx + FILT - 1, 31), np.float64)
    P[PAD_OFF:PAD_OFF + fy, PAD_OFF:PAD_OFF + fx] = feat                       # row y of `out` reads padded rows y .. y+9
    out = np.zeros((D, fy, fx), np.float64)
    for kh in range(FILT):
        for kw in range(FILT):
            win = P[kh:kh + fy, kw:kw + fx]                                     # [fy, fx, 31]
            out += np.einsum("yxc,dc->dyx", win, filters[:, :, kh, kw].astype(np.float64))
    return out.astype(f32)


def fhog_to_image(px, py):
    """dlib fhog_to_image for cell size 8 and a 10 x 10 filter: feature-map point (padded coordinates) -> pixel"""
    def one(p):
        v = (p + 1 - PAD_OFF) * CELL + 1
        return v + CELL // 2 if v >= 0 else v - CELL // 2
    return one(px), one(py)


def level_box(r, c):
    """detection at padded feature cell (row r, col c) -> (l, t, r, b) in that level's pixel coordinates:
    centered_rect(point(c, r), DET_BOX, DET_BOX) through fhog_to_image"""
    l, t = c - DET_BOX // 2, r - DET_BOX // 2
    rr, b = l + DET_BOX - 1, t + DET_BOX - 1
    L, T = fhog_to_image(l, t)
    R, B = fhog_to_image(rr, b)
    return L, T, R, B


def rect_up(box, levels, upsampled):
    """pyramid_down<6>::rect_up `levels` times (point_up: p * 6/5 + 0.3), then, when the image was upsampled once,
    pyramid_down<2>::rect_down (p / 2); corners rounded to the nearest integer"""
    out = []
    for v in box:
        p = float(v)
        for _ in range(levels):
            p = p * (6.0 / 5.0) + 0.3
        if upsampled:
            p = p / 2.0
        out.append(int(np.floor(p + 0.5)))
    return tuple(out)


def box_overlap(a, b, iou_thresh=NMS_IOU, covered_thresh=NMS_COVERED):
    """dlib test_box_overlap on inclusive integer rectangles"""
    def area(r):
        return max(r[2] - r[0] + 1, 0) * max(r[3] - r[1] + 1, 0)
    inner = area((max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])))
    if inner == 0:
        return False
    outer = area((min(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), max(a[3], b[3])))
    return inner / float(outer) > iou_thresh or inner / float(area(a)) > covered_thresh or inner / float(area(b)) > covered_thresh


def detect_levels(levels, filters, thresholds, upsampled=True, min_side=CELL * FILT):
    """levels: list of uint8 [h,w,3] pyramid levels (level 0 first).  Returns ([(l,t,r,b)], [score], [filter index]),
    sorted by score (descending; ties by level, filter, row, column) after greedy NMS."""
    cands = []
    for lv, img in enumerate(levels):
        if img.shape[0] < min_side or img.shape[1] < min_side:
            continue
        feat = fhog_features(fhog_hist(img))
        if feat.shape[0] < 1 or feat.shape[1] < 1:
            continue
        sc = score_maps(feat, filters)
        for d in range(filters.shape[0]):
            ys, xs = np.nonzero(sc[d] >= f32(thresholds[d]))
            for y, x in zip(ys, xs):
                # score map entry (y, x) = dlib saliency at padded (row y + 5, col x + 5)
                box = rect_up(level_box(int(y) + FILT // 2, int(x) + FILT // 2), lv, upsampled)
                cands.append((float(sc[d, y, x]), lv, d, int(y), int(x), box))
    cands.sort(key=lambda t: (-t[0], t[1], t[2], t[3], t[4]))
    boxes, scores, which = [], [], []
    for s, lv, d, y, x, box in cands:
        if any(box_overlap(box, k) for k in boxes):
            continue
        boxes.append(box)
        scores.append(s)
        which.append(d)
    return boxes, scores, which
