"""Image-pyramid geometry for the CNN (MMOD) detector: level sizes, tile packing, box mapping.

dlib's `input_rgb_image_pyramid<pyramid_down<6>>` tiles every pyramid level into ONE image so a
single pass of the conv stack scans all scales (reference call site: face_detector_(rgb, 1),
pyannote/video/face/face.py:66).  dlib's exact packing is not recalled, so the
packing below is our own shelf layout; it is pure integer host logic shared by the CUDA path and
the oracle, which therefore see the same plane.
"""
import numpy as np

PYR_N = 6            # pyramid_down<6>: each level is 5/6 of the previous one
PYR_PAD = 10         # zero pixels between tiles      (dlib pyramid_padding)
PYR_OUTER_PAD = 11   # zero pixels around the plane    (dlib pyramid_outer_padding)
PYR_MIN_SIDE = 5     # stop when a side would drop below this


def det_cell_to_plane(c, r):
    """centre (x,y) in plane pixels of detector output cell (col c, row r): through the conv stack
    output->input, p -> p*stride - pad + k/2 per layer = 8p + 14 for the MMOD face net."""
    from .weights import DET_CONVS, conv_pad
    def back(p):
        for (_, _, k, s) in reversed(DET_CONVS):
            p = p * s - conv_pad(k, s) + k // 2
        return p
    return back(c), back(r)


class PyramidGeometry:
    def __init__(self, H, W, upsample):
        self.H, self.W, self.upsample = H, W, int(upsample)
        h, w = (2 * H, 2 * W) if upsample else (H, W)
        sizes = []
        while min(h, w) >= PYR_MIN_SIDE:
            sizes.append((w, h))
            h, w = ((PYR_N - 1) * h) // PYR_N, ((PYR_N - 1) * w) // PYR_N
        self.sizes = sizes
        # ---- shelf packing: columns of tiles, each column as tall as level 0 ----
        w0, h0 = sizes[0]
        col_h = h0
        cols = []      # [x0, width, y_cursor]
        rects = []
        x_cursor = PYR_OUTER_PAD
        for (w, h) in sizes:
            placed = False
            for col in cols:
                if w <= col[1] and col[2] + h <= PYR_OUTER_PAD + col_h:
                    rects.append((col[0], col[2], w, h))
                    col[2] += h + PYR_PAD
                    placed = True
                    break
            if not placed:
                cols.append([x_cursor, w, PYR_OUTER_PAD + h + PYR_PAD])
                rects.append((x_cursor, PYR_OUTER_PAD, w, h))
                x_cursor += w + PYR_PAD
        self.rects = rects                                    # (x0, y0, w, h) per level
        self.plane_w = x_cursor - PYR_PAD + PYR_OUTER_PAD
        self.plane_w = (self.plane_w + 3) & ~3    # 16-byte row pitch: the first conv reads raw pixel rows by TMA
        self.plane_h = col_h + 2 * PYR_OUTER_PAD
        # ---- float32 factors mapping level-local coordinates to original-image coordinates ----
        f32 = np.float32
        fx, fy = [], []
        for (w, h) in sizes:
            sx = f32(w0 - 1) / f32(max(w - 1, 1))
            sy = f32(h0 - 1) / f32(max(h - 1, 1))
            if upsample:
                sx = f32(sx * (f32(W - 1) / f32(max(w0 - 1, 1))))
                sy = f32(sy * (f32(H - 1) / f32(max(h0 - 1, 1))))
            fx.append(f32(sx))
            fy.append(f32(sy))
        self.fx = np.asarray(fx, f32)
        self.fy = np.asarray(fy, f32)

    @property
    def n_levels(self):
        return len(self.sizes)

    def level_table(self):
        """int32 [L,4] rects and float32 [L,2] factors, as passed to the decode kernel."""
        return (np.asarray(self.rects, np.int32).reshape(-1, 4),
                np.stack([self.fx, self.fy], axis=1).astype(np.float32))

    def level_at(self, px, py):
        for lv, (x0, y0, w, h) in enumerate(self.rects):
            if x0 <= px < x0 + w and y0 <= py < y0 + h:
                return lv
        return -1

    def box_from_plane(self, lv, px, py, window):
        """window x window box centred on plane pixel (px,py) of level lv -> integer (l,t,r,b) in
        the original image (dlib centered_rect, then level -> image scaling, round to nearest)."""
        f32 = np.float32
        x0, y0, _, _ = self.rects[lv]
        l = px - window // 2 - x0
        t = py - window // 2 - y0
        r = l + window - 1
        b = t + window - 1
        half = f32(0.5)
        def m(v, f):
            return int(np.floor(f32(f32(v) * f) + half))
        return (m(l, self.fx[lv]), m(t, self.fy[lv]), m(r, self.fx[lv]), m(b, self.fy[lv]))

    def total_level_pixels(self):
        return sum(w * h for w, h in self.sizes)


_cache = {}


def pyramid_geometry(H, W, upsample=1):
    key = (H, W, int(upsample))
    if key not in _cache:
        _cache[key] = PyramidGeometry(H, W, upsample)
    return _cache[key]
