"""Image-pyramid geometry for the CNN (MMOD) detector: level sizes, tile packing, box mapping.

dlib's `input_rgb_image_pyramid<pyramid_down<6>>` tiles every pyramid level into ONE image so a
single pass of the conv stack scans all scales (reference call site: face_detector_(rgb, 1),
pyannote/video/face/face.py:66).  dlib's exact packing is not recalled, so the
packing below is our own guillotine layout; it is pure integer host logic shared by the CUDA path and
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


def _guillotine(sizes, pw, ph):
    """place the tiles (each followed by PYR_PAD zero pixels to its right and below) into a pw x ph area;
    returns [(x, y, w, h)] or None if they do not fit"""
    free = [(0, 0, pw + PYR_PAD, ph + PYR_PAD)]
    out = []
    for (w, h) in sizes:
        ww, hh = w + PYR_PAD, h + PYR_PAD
        best = None
        for i, (fx, fy, fw, fh) in enumerate(free):
            if ww <= fw and hh <= fh:
                score = min(fw - ww, fh - hh)
                if best is None or score < best[0]:
                    best = (score, i)
        if best is None:
            return None
        fx, fy, fw, fh = free.pop(best[1])
        out.append((fx, fy, w, h))
        r1, b1 = (fx + ww, fy, fw - ww, hh), (fx, fy + hh, fw, fh - hh)      # split along the tile's bottom edge
        r2, b2 = (fx + ww, fy, fw - ww, fh), (fx, fy + hh, ww, fh - hh)      # split along the tile's right edge
        cand = (r1, b1) if max(r1[2] * r1[3], b1[2] * b1[3]) > max(r2[2] * r2[3], b2[2] * b2[3]) else (r2, b2)
        for c in cand:
            if c[2] > PYR_PAD + PYR_MIN_SIDE - 1 and c[3] > PYR_PAD + PYR_MIN_SIDE - 1:
                free.append(c)
    return out


class PyramidGeometry:
    def __init__(self, H, W, upsample):
        self.H, self.W, self.upsample = H, W, int(upsample)
        h, w = (2 * H, 2 * W) if upsample else (H, W)
        sizes = []
        while min(h, w) >= PYR_MIN_SIDE:
            sizes.append((w, h))
            h, w = ((PYR_N - 1) * h) // PYR_N, ((PYR_N - 1) * w) // PYR_N
        self.sizes = sizes
        # ---- guillotine packing (our own): the conv stack scans the whole plane, so every
        # padding pixel costs as much as an image pixel.  Levels go, largest first, into the free rectangle
        # with the best short-side fit; a few plane heights are tried and the smallest plane wins
        # (1080p, upsample 1: 27.1 M level pixels in a 28.9 M-pixel plane, 93.8 %).
        w0, h0 = sizes[0]
        best = None
        heights = [h0] + [h0 + sizes[i][1] + PYR_PAD for i in (5, 3, 2, 1) if i < len(sizes)]
        for ph in heights:
            for pw in range(w0, 8 * w0 + 8, 8):
                rects = _guillotine(sizes, pw, ph)
                if rects is not None:
                    pw_al = (pw + 2 * PYR_OUTER_PAD + 3) & ~3
                    tot = pw_al * (ph + 2 * PYR_OUTER_PAD)
                    if best is None or tot < best[0]:
                        best = (tot, pw_al, ph + 2 * PYR_OUTER_PAD, rects)
                    break
        assert best is not None
        self.rects = [(x + PYR_OUTER_PAD, y + PYR_OUTER_PAD, w, h) for (x, y, w, h) in best[3]]   # (x0, y0, w, h) per level
        self.plane_w = best[1]    # multiple of 4 pixels = 16-byte row pitch: the first conv reads raw pixel rows by TMA
        self.plane_h = best[2]
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
