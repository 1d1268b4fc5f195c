"""ORACLE (test infrastructure — never imported by the product path).

The oracle's OWN, deliberately slow and obvious, geometry of the tiled image pyramid that dlib's
`input_rgb_image_pyramid<pyramid_down<6>>` feeds the MMOD detector (reference call site
pyannote/video/face/face.py:66): level sizes, a validator for a tile placement, the map from a
detector output cell back to a plane pixel (by walking the conv stack), the level lookup and the
box back-projection into image coordinates.  Nothing here imports the product's `pyrgeom`.

The tile PLACEMENT itself is [OURS] (dlib's exact packing is not recalled): it
is an input to the oracle (`Geometry(H, W, upsample, rects, plane_h, plane_w)`), never trusted — the
constructor checks it against the invariants any correct placement must satisfy (every level
present at its own size, inside the plane with the outer padding, tiles at least PYR_PAD apart), so
that a packing bug on the product side fails here instead of passing bit-exact parity.
"""
import numpy as np

from . import constants as K

f32 = np.float32


def level_sizes(H, W, upsample):
    """[(w, h)] of every pyramid level: level 0 is the (optionally 2x upsampled) image, each next
    level is floor(5/6) of the previous (pyramid_down<6>), until a side drops below PYR_MIN_SIDE."""
    if upsample:
        h, w = 2 * H, 2 * W
    else:
        h, w = H, W
    out = []
    while h >= K.PYR_MIN_SIDE and w >= K.PYR_MIN_SIDE:
        out.append((w, h))
        h = ((K.PYR_N - 1) * h) // K.PYR_N
        w = ((K.PYR_N - 1) * w) // K.PYR_N
    return out


def cell_to_plane(c):
    """plane coordinate of the centre of detector output cell index c (same formula for x and y):
    walk the conv stack from the output back to the input, p -> p*stride - pad + k//2."""
    p = c
    for (_, _, k, s) in reversed(K.DET_CONVS):
        p = p * s - K.conv_pad(k, s) + k // 2
    return p


class Geometry(object):
    def __init__(self, H, W, upsample, rects, plane_h, plane_w):
        self.H, self.W, self.upsample = int(H), int(W), int(upsample)
        self.sizes = level_sizes(H, W, upsample)
        self.rects = [tuple(int(v) for v in r) for r in rects]
        self.plane_h, self.plane_w = int(plane_h), int(plane_w)
        self.check_placement()

    # ---- validation of a placement handed to the oracle ----
    def check_placement(self):
        if len(self.rects) != len(self.sizes):
            raise AssertionError("placement has %d tiles, the pyramid has %d levels" % (len(self.rects), len(self.sizes)))
        for lv, ((x0, y0, w, h), (sw, sh)) in enumerate(zip(self.rects, self.sizes)):
            if (w, h) != (sw, sh):
                raise AssertionError("level %d: tile is %dx%d, level is %dx%d" % (lv, w, h, sw, sh))
            if x0 < K.PYR_OUTER_PAD or y0 < K.PYR_OUTER_PAD or x0 + w + K.PYR_OUTER_PAD > self.plane_w \
                    or y0 + h + K.PYR_OUTER_PAD > self.plane_h:
                raise AssertionError("level %d: tile violates the outer padding" % lv)
        n = len(self.rects)
        for i in range(n):
            ax, ay, aw, ah = self.rects[i]
            for j in range(i + 1, n):
                bx, by, bw, bh = self.rects[j]
                gap_x = max(bx - (ax + aw), ax - (bx + bw))
                gap_y = max(by - (ay + ah), ay - (by + bh))
                if max(gap_x, gap_y) < K.PYR_PAD:
                    raise AssertionError("tiles %d and %d are closer than PYR_PAD" % (i, j))

    # ---- lookups ----
    def level_at(self, px, py):
        hit = -1
        for lv, (x0, y0, w, h) in enumerate(self.rects):
            if x0 <= px <= x0 + w - 1 and y0 <= py <= y0 + h - 1:
                assert hit < 0
                hit = lv
        return hit

    def level_factors(self, lv):
        """float32 (fx, fy): level-local pixel coordinate -> original-image coordinate.  Resizing maps corners to
        corners (scale (in-1)/(out-1), dlib resize_image), so level lv -> level 0 is (w0-1)/(w-1) and, with an
        upsampled level 0, level 0 -> image is (W-1)/(w0-1); both evaluated in float32 in this order."""
        w0, h0 = self.sizes[0]
        w, h = self.sizes[lv]
        fx = f32(w0 - 1) / f32(max(w - 1, 1))
        fy = f32(h0 - 1) / f32(max(h - 1, 1))
        if self.upsample:
            fx = f32(fx * (f32(self.W - 1) / f32(max(w0 - 1, 1))))
            fy = f32(fy * (f32(self.H - 1) / f32(max(h0 - 1, 1))))
        return f32(fx), f32(fy)

    def box_from_plane(self, lv, px, py, window):
        """window x window box centred (dlib centered_rect) on plane pixel (px, py) of level lv, mapped to the
        original image and rounded to the nearest integer: inclusive (l, t, r, b)."""
        x0, y0, _, _ = self.rects[lv]
        l = (px - x0) - window // 2
        t = (py - y0) - window // 2
        r = l + window - 1
        b = t + window - 1
        fx, fy = self.level_factors(lv)

        def m(v, f):
            return int(np.floor(f32(f32(v) * f) + f32(0.5)))
        return (m(l, fx), m(t, fy), m(r, fx), m(b, fy))

    def total_level_pixels(self):
        return sum(w * h for (w, h) in self.sizes)


def from_product(geo):
    """adopt (and validate) the placement chosen by the product's pyramid_geometry object"""
    return Geometry(geo.H, geo.W, geo.upsample, geo.rects, geo.plane_h, geo.plane_w)
