"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of dlib.correlation_tracker (DSST, Danelljan et al. BMVC'14) as the reference uses
it: `start_track(frame, drectangle)`, `update(frame) -> PSR`, `get_position()`
(pyannote/video/tracking.py:203,231,250-251), following dlib 19.12's correlation_tracker.h as
recalled in SURVEY.md App. A.5 (dlib absent: parity unpinned).

Restated: translation filter over 32 feature planes of a 64x64 chip cut from the position rectangle grown by
1.4 — the 31-channel FHOG (cell size 1) plus, as the 32nd plane, the chip's intensity / 255 (dlib's make_chip:
"the HOG features ignore the overall brightness ... so we add it as the 32nd feature") — cosine window,
per-plane numerators A_i = conj(G) F_i and shared denominator B = sum |F_i|^2, response =
ifft2(sum F_i conj(A_i) / (B + 0.001)), sub-pixel peak as in dlib's max_point_interpolated (Newton step of the
3x3 finite-difference quadratic INCLUDING the cross term, clamped to +-1 pixel), PSR over everything outside the
8x8 peak window, running update with nu = 0.025.  Still not restated: extract_image_chip's pyramid_down
pre-shrink of chips whose source region is more than twice the chip (DESIGN.md audit list).
Also restated: the 1-D scale filter of `update` — 32 scales alpha^(k-16) (alpha = 1.020) of the
position rectangle, each resampled to 23x23, FHOG with cell size 4 (4x4 cells x 31 = 496 features),
Hann window over scales, per-feature numerators As_j = conj(Gs) Fs_j and denominator Bs = sum |Fs_j|^2
over a length-32 DFT, interpolated argmax, position *= alpha^(p-16), running update with nu = 0.025.
Chip sampling and FHOG are float32/unfused like the CUDA path; the FFTs are numpy float64.
"""
import numpy as np

f32 = np.float32

FS = 64
NCH = 32          # 31 FHOG planes + the intensity plane
PADDING = 1.4
LAMBDA = 0.001
NU = 0.025
EPS = f32(0.0001)


def extract_chip(rgb, rect):
    """bilinear 64x64 RGB chip (uint8) of rect*PADDING; chip pixel (cx,cy) <-> image point
    R.l + cx*(R.r-R.l)/63 (corners map to corners); outside the image -> 0."""
    H, W, _ = rgb.shape
    l, t, r, b = [f32(v) for v in rect]
    cx, cy = (l + r) * f32(0.5), (t + b) * f32(0.5)
    hw, hh = (r - l) * f32(0.5) * f32(PADDING), (b - t) * f32(0.5) * f32(PADDING)
    rl, rt = cx - hw, cy - hh
    sx, sy = (f32(2.0) * hw) / f32(FS - 1), (f32(2.0) * hh) / f32(FS - 1)
    xs = (rl + np.arange(FS, dtype=f32) * sx).astype(f32)
    ys = (rt + np.arange(FS, dtype=f32) * sy).astype(f32)
    left = np.floor(xs).astype(np.int64)
    top = np.floor(ys).astype(np.int64)
    lr = (xs - left.astype(f32)).astype(f32)[None, :, None]
    tb = (ys - top.astype(f32)).astype(f32)[:, None, None]
    ok = ((left >= 0) & (left + 1 < W))[None, :] & ((top >= 0) & (top + 1 < H))[:, None]
    lc, tc = np.clip(left, 0, W - 2), np.clip(top, 0, H - 2)
    s = rgb.astype(f32)
    tl = s[tc][:, lc]
    tr = s[tc][:, lc + 1]
    bl = s[tc + 1][:, lc]
    br = s[tc + 1][:, lc + 1]
    one = f32(1)
    a = ((one - lr) * tl) + (lr * tr)
    bb = ((one - lr) * bl) + (lr * br)
    v = ((one - tb) * a) + (tb * bb)
    v = np.clip(np.floor(v + f32(0.5)), 0, 255)
    v = np.where(ok[..., None], v, 0)
    return v.astype(np.uint8), (rl, rt, sx, sy)


_UU = np.cos(np.arange(9) * np.pi / 9).astype(f32)
_VV = np.sin(np.arange(9) * np.pi / 9).astype(f32)


def fhog_cell1(chip):
    """31-channel FHOG with cell size 1 of a uint8 [64,64,3] chip -> float32 [31,64,64]; the
    outermost ring of cells is zero (dlib pads the (n-2)^2 valid cells back to n^2)."""
    n = chip.shape[0]
    c = chip.astype(f32)
    dx = np.zeros((n, n, 3), f32)
    dy = np.zeros((n, n, 3), f32)
    dx[1:-1, 1:-1] = c[1:-1, 2:] - c[1:-1, :-2]
    dy[1:-1, 1:-1] = c[2:, 1:-1] - c[:-2, 1:-1]
    m2 = (dx * dx) + (dy * dy)
    ch = np.argmax(m2, axis=2)                       # first max: channel order r,g,b
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    gx, gy, v2 = dx[ii, jj, ch], dy[ii, jj, ch], m2[ii, jj, ch]
    mag = np.sqrt(v2).astype(f32)
    best_dot = np.zeros((n, n), f32)
    best_o = np.zeros((n, n), np.int64)
    for o in range(9):
        dot = (_UU[o] * gx) + (_VV[o] * gy)
        pos = dot > best_dot
        best_o = np.where(pos, o, best_o)
        best_dot = np.where(pos, dot, best_dot)
        neg = (~pos) & (-dot > best_dot)
        best_o = np.where(neg, o + 9, best_o)
        best_dot = np.where(neg, -dot, best_dot)
    interior = np.zeros((n, n), bool)
    interior[1:-1, 1:-1] = True
    mag = np.where(interior, mag, 0).astype(f32)
    nrm = (mag * mag).astype(f32)
    P = np.pad(nrm, 1)

    def blk(dy_, dx_):
        # sum of the 2x2 block whose top-left cell is (y+dy_, x+dx_)
        y0, x0 = 1 + dy_, 1 + dx_
        return (((P[y0:y0 + n, x0:x0 + n] + P[y0:y0 + n, x0 + 1:x0 + n + 1]) + P[y0 + 1:y0 + n + 1, x0:x0 + n])
                + P[y0 + 1:y0 + n + 1, x0 + 1:x0 + n + 1]).astype(f32)

    ns = [f32(1) / np.sqrt(blk(-1, -1) + EPS), f32(1) / np.sqrt(blk(-1, 0) + EPS),
          f32(1) / np.sqrt(blk(0, -1) + EPS), f32(1) / np.sqrt(blk(0, 0) + EPS)]
    h = [np.minimum(mag * nk, f32(0.2)).astype(f32) for nk in ns]
    osum = (f32(0.5) * (((h[0] + h[1]) + h[2]) + h[3])).astype(f32)
    out = np.zeros((31, n, n), f32)
    valid = np.zeros((n, n), bool)
    valid[1:-1, 1:-1] = True
    for o in range(18):
        out[o] = np.where(valid & (best_o == o), osum, 0)
    for o in range(9):
        out[18 + o] = np.where(valid & ((best_o % 9) == o), osum, 0)
    for k in range(4):
        out[27 + k] = np.where(valid, f32(0.2357) * h[k], 0)
    return out


def gray_plane(chip):
    """dlib assign_pixel rgb -> grayscale: (r + g + b) / 3 in unsigned integers, then / 255 (float32)"""
    g = (chip.astype(np.uint32).sum(axis=2) // 3).astype(f32)
    return (g / f32(255)).astype(f32)


def interpolated_peak(R):
    """dlib max_point_interpolated: argmax (first in raster order); inside the border a Newton step of the quadratic
    through the 3x3 neighbourhood (central differences, cross term included), each component clamped to [-1, 1]"""
    py, px = np.unravel_index(int(np.argmax(R)), R.shape)
    ppx, ppy = float(px), float(py)
    if 0 < px < R.shape[1] - 1 and 0 < py < R.shape[0] - 1:
        dx = 0.5 * (R[py, px + 1] - R[py, px - 1])
        dy = 0.5 * (R[py + 1, px] - R[py - 1, px])
        dxx = R[py, px + 1] - 2 * R[py, px] + R[py, px - 1]
        dyy = R[py + 1, px] - 2 * R[py, px] + R[py - 1, px]
        dxy = 0.25 * ((R[py + 1, px + 1] + R[py - 1, px - 1]) - (R[py + 1, px - 1] + R[py - 1, px + 1]))
        det = dxx * dyy - dxy * dxy
        if det != 0:
            ppx += float(np.clip(-(dyy * dx - dxy * dy) / det, -1.0, 1.0))
            ppy += float(np.clip(-(dxx * dy - dxy * dx) / det, -1.0, 1.0))
    return int(py), int(px), ppx, ppy


def hann2d(n=FS):
    w = (f32(0.5) - f32(0.5) * np.cos(2 * np.pi * np.arange(n) / (n - 1))).astype(f32)
    return (w[:, None] * w[None, :]).astype(f32)


def gaussian_target(px, py, n=FS):
    x = np.arange(n, dtype=f32)[None, :]
    y = np.arange(n, dtype=f32)[:, None]
    return np.exp(-(((x - f32(px)) ** 2) + ((y - f32(py)) ** 2)) / f32(3.0)).astype(f32)


# ---------------------------------------------------------------------------------------------
# scale filter
# ---------------------------------------------------------------------------------------------
N_SCALES = 32
SCALE_WINDOW = 23
SCALE_ALPHA = 1.020
SCALE_LAMBDA = 0.001
SCALE_NU = 0.025
SCALE_CELL = 4


def extract_chip_n(rgb, rect, n):
    """bilinear n x n RGB chip of `rect` (no padding); corners map to corners; outside -> 0"""
    H, W, _ = rgb.shape
    l, t, r, b = [f32(v) for v in rect]
    sx, sy = (r - l) / f32(n - 1), (b - t) / f32(n - 1)
    xs = (l + np.arange(n, dtype=f32) * sx).astype(f32)
    ys = (t + np.arange(n, dtype=f32) * sy).astype(f32)
    left = np.floor(xs).astype(np.int64)
    top = np.floor(ys).astype(np.int64)
    lr = (xs - left.astype(f32)).astype(f32)[None, :, None]
    tb = (ys - top.astype(f32)).astype(f32)[:, None, None]
    ok = ((left >= 0) & (left + 1 < W))[None, :] & ((top >= 0) & (top + 1 < H))[:, None]
    lc, tc = np.clip(left, 0, W - 2), np.clip(top, 0, H - 2)
    s = rgb.astype(f32)
    tl, tr, bl, br = s[tc][:, lc], s[tc][:, lc + 1], s[tc + 1][:, lc], s[tc + 1][:, lc + 1]
    one = f32(1)
    a = ((one - lr) * tl) + (lr * tr)
    bb = ((one - lr) * bl) + (lr * br)
    v = ((one - tb) * a) + (tb * bb)
    v = np.clip(np.floor(v + f32(0.5)), 0, 255)
    return np.where(ok[..., None], v, 0).astype(np.uint8)


def fhog_cell4(chip):
    """31-channel FHOG, cell size 4, of a uint8 [n,n,3] chip -> float32 [31, cells-2, cells-2]
    (cells = round(n/4)); votes are bilinearly shared between the 4 nearest cells, in raster order."""
    n = chip.shape[0]
    cells = int(round(n / float(SCALE_CELL)))
    c = chip.astype(f32)
    hist = np.zeros((cells, cells, 18), f32)
    for y in range(1, n - 1):
        for x in range(1, n - 1):
            best = f32(-1)
            gx = gy = f32(0)
            for ch in range(3):
                dx = c[y, x + 1, ch] - c[y, x - 1, ch]
                dy = c[y + 1, x, ch] - c[y - 1, x, ch]
                v = (dx * dx) + (dy * dy)
                if v > best:
                    best, gx, gy = v, dx, dy
            mag = np.sqrt(best)
            bo, best_dot = 0, f32(0)
            for o in range(9):
                dot = (_UU[o] * gx) + (_VV[o] * gy)
                if dot > best_dot:
                    best_dot, bo = dot, o
                elif -dot > best_dot:
                    best_dot, bo = -dot, o + 9
            xp = (f32(x) + f32(0.5)) / f32(SCALE_CELL) - f32(0.5)
            yp = (f32(y) + f32(0.5)) / f32(SCALE_CELL) - f32(0.5)
            ixp, iyp = int(np.floor(xp)), int(np.floor(yp))
            vx0, vy0 = xp - f32(ixp), yp - f32(iyp)
            vx1, vy1 = f32(1) - vx0, f32(1) - vy0
            if ixp >= 0 and iyp >= 0:
                hist[iyp, ixp, bo] += (vx1 * vy1) * mag
            if ixp + 1 < cells and iyp >= 0:
                hist[iyp, ixp + 1, bo] += (vx0 * vy1) * mag
            if ixp >= 0 and iyp + 1 < cells:
                hist[iyp + 1, ixp, bo] += (vx1 * vy0) * mag
            if ixp + 1 < cells and iyp + 1 < cells:
                hist[iyp + 1, ixp + 1, bo] += (vx0 * vy0) * mag
    nrm = np.zeros((cells, cells), f32)
    for o in range(9):
        s_ = hist[:, :, o] + hist[:, :, o + 9]
        nrm = nrm + (s_ * s_)
    oc = cells - 2
    out = np.zeros((31, oc, oc), f32)
    for y in range(oc):
        for x in range(oc):
            Y, X = y + 1, x + 1
            h = hist[Y, X]
            ns = []
            for (dy_, dx_) in ((-1, -1), (-1, 0), (0, -1), (0, 0)):
                y0, x0 = Y + dy_, X + dx_
                blk = ((nrm[y0, x0] + nrm[y0, x0 + 1]) + nrm[y0 + 1, x0]) + nrm[y0 + 1, x0 + 1]
                ns.append(f32(1) / np.sqrt(blk + EPS))
            t = [f32(0)] * 4
            for o in range(18):
                hk = [min(h[o] * ns[k], f32(0.2)) for k in range(4)]
                out[o, y, x] = f32(0.5) * (((hk[0] + hk[1]) + hk[2]) + hk[3])
                for k in range(4):
                    t[k] = t[k] + hk[k]
            for o in range(9):
                s_ = h[o] + h[o + 9]
                hk = [min(s_ * ns[k], f32(0.2)) for k in range(4)]
                out[18 + o, y, x] = f32(0.5) * (((hk[0] + hk[1]) + hk[2]) + hk[3])
            for k in range(4):
                out[27 + k, y, x] = f32(0.2357) * t[k]
    return out


def scale_rect(rect, factor):
    l, t, r, b = [float(v) for v in rect]
    cx, cy = 0.5 * (l + r), 0.5 * (t + b)
    hw, hh = 0.5 * (r - l) * factor, 0.5 * (b - t) * factor
    return (cx - hw, cy - hh, cx + hw, cy + hh)


def scale_space(rgb, position):
    """float64 [496, 32]: feature j at scale k, windowed over scales"""
    w = (f32(0.5) - f32(0.5) * np.cos(2 * np.pi * np.arange(N_SCALES) / (N_SCALES - 1))).astype(f32)
    cols = []
    for k in range(N_SCALES):
        factor = f32(SCALE_ALPHA) ** f32(k - N_SCALES // 2)
        l, t, r, b = [f32(v) for v in position]
        cx, cy = (l + r) * f32(0.5), (t + b) * f32(0.5)
        hw, hh = ((r - l) * f32(0.5)) * factor, ((b - t) * f32(0.5)) * factor
        chip = extract_chip_n(rgb, (cx - hw, cy - hh, cx + hw, cy + hh), SCALE_WINDOW)
        cols.append((fhog_cell4(chip).reshape(-1) * w[k]).astype(f32))
    return np.stack(cols, axis=1).astype(np.float64)


def scale_target(p):
    k = np.arange(N_SCALES, dtype=f32)
    return np.exp(-((k - f32(p)) ** 2) / f32(1.0)).astype(f32)


class CorrelationTracker(object):
    def __init__(self, use_scale=True):
        self.use_scale = use_scale

    def _features(self, rgb, rect):
        chip, tf = extract_chip(rgb, rect)
        planes = np.concatenate([fhog_cell1(chip), gray_plane(chip)[None]], axis=0)
        F = (planes * hann2d()[None]).astype(f32)
        return np.fft.fft2(F.astype(np.float64)), tf

    def start_track(self, rgb, rect):
        self.position = tuple(float(v) for v in rect)
        F, _ = self._features(rgb, self.position)
        c = (FS - 1) / 2.0
        G = np.conj(np.fft.fft2(gaussian_target(c, c).astype(np.float64)))
        self.A = G[None] * F
        self.B = (np.abs(F) ** 2).sum(0)
        if self.use_scale:
            Fs = np.fft.fft(scale_space(rgb, self.position), axis=1)
            Gs = np.conj(np.fft.fft(scale_target(N_SCALES // 2).astype(np.float64)))
            self.As = Gs[None] * Fs
            self.Bs = (np.abs(Fs) ** 2).sum(0)

    def update_scale(self, rgb):
        Fs = np.fft.fft(scale_space(rgb, self.position), axis=1)
        r = np.real(np.fft.ifft((Fs * np.conj(self.As)).sum(0) / (self.Bs + SCALE_LAMBDA)))
        pk = int(np.argmax(r))
        p = float(pk)
        if 0 < pk < N_SCALES - 1:
            d = r[pk - 1] - 2 * r[pk] + r[pk + 1]
            if d != 0:
                # dlib's 1-D max_point_interpolated (lagrange_poly_min_extrap) stays inside [pk - 1, pk + 1]
                p += float(np.clip(0.5 * (r[pk - 1] - r[pk + 1]) / d, -1.0, 1.0))
        self.position = scale_rect(self.position, SCALE_ALPHA ** (p - N_SCALES // 2))
        Gs = np.conj(np.fft.fft(scale_target(p).astype(np.float64)))
        self.As = (1 - SCALE_NU) * self.As + SCALE_NU * (Gs[None] * Fs)
        self.Bs = (1 - SCALE_NU) * self.Bs + SCALE_NU * (np.abs(Fs) ** 2).sum(0)
        return p

    def update(self, rgb):
        psr = self.update_noscale(rgb)
        if self.use_scale:
            self.update_scale(rgb)
        return psr

    def update_noscale(self, rgb):
        guess = self.position
        F, (rl, rt, sx, sy) = self._features(rgb, guess)
        R = np.real(np.fft.ifft2((F * np.conj(self.A)).sum(0) / (self.B + LAMBDA)))
        py, px, ppx, ppy = interpolated_peak(R)
        mask = np.ones_like(R, bool)
        mask[max(py - 4, 0):py + 4, max(px - 4, 0):px + 4] = False
        side = R[mask]
        psr = float((R[py, px] - side.mean()) / side.std(ddof=1))
        ix, iy = float(rl) + ppx * float(sx), float(rt) + ppy * float(sy)
        cx, cy = 0.5 * (guess[0] + guess[2]), 0.5 * (guess[1] + guess[3])
        ddx, ddy = ix - cx, iy - cy
        self.position = (guess[0] + ddx, guess[1] + ddy, guess[2] + ddx, guess[3] + ddy)
        G = np.conj(np.fft.fft2(gaussian_target(ppx, ppy).astype(np.float64)))
        self.A = (1 - NU) * self.A + NU * (G[None] * F)
        self.B = (1 - NU) * self.B + NU * (np.abs(F) ** 2).sum(0)
        return psr

    def get_position(self):
        return self.position
