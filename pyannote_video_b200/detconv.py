"""Host side of csrc/detconv.cu: the detector's conv layers 2..7 as 2-D tiled implicit GEMMs.

Activations are plain NHWC bf16 tensors [B, H, pitch, C] (pitch = W rounded up to even; the extra
column stays zero), weights are packed on the host into the exact shared-memory image the kernel's
UMMA descriptors read (include/pv_b200.h, PvDetconvDesc).  Replaces the `con` layers of dlib's MMOD
CNN behind face_detector_(rgb, 1), pyannote/video/face/face.py:66.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

INSTANCES = {  # (c_in, n_out, kh, kw, stride, out_f32)
    (16, 32, 5, 5, 2, 0), (32, 32, 5, 5, 2, 0), (32, 48, 5, 5, 1, 0), (48, 48, 5, 5, 1, 0), (48, 16, 9, 1, 1, 1)}
# rsconv additionally serves the embedder's 3x3 layers of levels 4 and 3 (faces packed side by side in an image row)
# ... and the HOG detector's sliding-window filters (10 x 10 cells x 32 features, filters as output channels)
INSTANCES_RS = INSTANCES | {(32, 32, 3, 3, 1, 0), (32, 64, 3, 3, 2, 0), (64, 64, 3, 3, 1, 0), (32, 16, 10, 10, 1, 1)}


def even(n):
    return (n + 1) // 2 * 2


def pack_weight_image(w, c_in, n_out):
    """w float [Cout, Cin, KH, KW] -> uint8 image: for tap t = kh*KW+kw and 16-channel chunk j an
    n_out x 16 bf16 tile at byte (t*(c_in/16)+j)*n_out*32, element (n,k) at
    (k>>3)*(n_out*16) + (n>>3)*128 + (n&7)*16 + (k&7)*2."""
    w = torch.as_tensor(w).float()
    Cout, Cin, KH, KW = w.shape
    assert Cout <= n_out and Cin <= c_in and c_in % 16 == 0 and n_out % 16 == 0
    kch = c_in // 16
    full = torch.zeros(KH, KW, kch, n_out, 16)
    wp = torch.zeros(n_out, c_in, KH, KW)
    wp[:Cout, :Cin] = w
    full[:] = wp.permute(2, 3, 1, 0).reshape(KH, KW, kch, 16, n_out).permute(0, 1, 2, 4, 3)   # [kh,kw,j,n,k]
    # tile layout: [k>>3][n>>3][n&7][k&7]
    t = full.reshape(KH * KW * kch, n_out // 8, 8, 2, 8).permute(0, 3, 1, 2, 4).contiguous()
    img = t.to(torch.bfloat16).view(torch.uint8).reshape(-1)
    assert img.numel() == KH * KW * kch * n_out * 32
    return img


def unpack_weight_image(img, c_in, n_out, KH, KW):
    """inverse of pack_weight_image (element-address formula written out; used by the CPU test)."""
    kch = c_in // 16
    v = img.view(torch.bfloat16).float().numpy()
    out = np.zeros((n_out, c_in, KH, KW), np.float32)
    for t in range(KH * KW):
        for j in range(kch):
            base = (t * kch + j) * n_out * 32
            for n in range(n_out):
                for k in range(16):
                    off = base + (k >> 3) * (n_out * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2
                    out[n, j * 16 + k, t // KW, t % KW] = v[off // 2]
    return out


class DetConv:
    """One bound conv layer: x [Bmax,H,pitch,c_in] bf16 -> out [Bmax,OH,out_pitch,out_cs] (bf16, or f32
    when out_f32).  Same `run(B)` / `check()` protocol as plan.Srgemm."""

    _create, _run, _check, _destroy = "pv_detconv_create", "pv_detconv_run", "pv_detconv_check", "pv_detconv_destroy"

    @staticmethod
    def _pack(weight, c_in, n_out, stride):
        return pack_weight_image(weight, c_in, n_out)

    _instances = INSTANCES

    def __init__(self, x, H, W, weight, stride, scale, shift, relu, c_in, n_out, out_f32=False, out_cs=None, out=None,
                 resid=None, gap=None):
        dev = x.device
        Cout, Cin, KH, KW = weight.shape
        key = (c_in, n_out, KH, KW, stride, int(out_f32))
        if key not in self._instances:
            raise _lib.PvError("%s: no kernel instance for %r" % (type(self).__name__, key))
        B, Hx, pitch, Cx = x.shape
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and (Hx, Cx) == (H, c_in) and pitch == even(W)
        pad_y, pad_x = (KH // 2, KW // 2) if stride == 1 else (0, 0)
        self.OH = (H + 2 * pad_y - KH) // stride + 1
        self.OW = (W + 2 * pad_x - KW) // stride + 1
        self.out_pitch = even(self.OW)
        self.out_cs = out_cs or n_out
        if out is None:
            out = torch.zeros(B, self.OH, self.out_pitch, self.out_cs, dtype=torch.float32 if out_f32 else torch.bfloat16,
                              device=dev)
        assert tuple(out.shape) == (B, self.OH, self.out_pitch, self.out_cs) and out.is_contiguous()
        self.out = out
        self.resid = resid
        if resid is not None:
            assert tuple(resid.shape) == tuple(out.shape) and resid.dtype == torch.bfloat16 and resid.is_contiguous()
        self.w_img = self._pack(weight, c_in, n_out, stride).to(dev)
        sc = torch.zeros(n_out, dtype=torch.float32)
        sh = torch.zeros(n_out, dtype=torch.float32)
        sc[:Cout] = torch.as_tensor(scale).float()
        sh[:Cout] = torch.as_tensor(shift).float()
        self.scale, self.shift = sc.to(dev), sh.to(dev)
        self.x = x
        self.Bmax = B
        d = _lib.PvDetconvDesc()
        d.x = x.data_ptr()
        d.B, d.H, d.W, d.pitch = B, H, W, pitch
        d.c_in, d.n_out, d.kh, d.kw, d.stride, d.out_f32 = c_in, n_out, KH, KW, stride, int(out_f32)
        d.w_img, d.w_bytes = self.w_img.data_ptr(), self.w_img.numel()
        d.scale, d.shift = self.scale.data_ptr(), self.shift.data_ptr()
        d.relu = int(relu)
        d.out = self.out.data_ptr()
        d.out_pitch, d.out_cs = self.out_pitch, self.out_cs
        d.resid = resid.data_ptr() if resid is not None else None
        d.gap_period, d.gap_pos = (int(gap[0]), int(gap[1])) if gap else (0, 0)
        h = C.c_void_p()
        _lib.check(getattr(_lib.lib(), self._create)(C.byref(d), C.byref(h)), self._create)
        self.h = h
        self.flops_per_image = 2 * self.OH * self.OW * Cout * Cin * KH * KW

    def info(self):
        a, b, c, e = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().pv_detconv_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(e)), "pv_detconv_info")
        return dict(n_stages=a.value, smem_bytes=b.value, tiles_x=c.value, tiles_y=e.value)

    def run(self, B=None):
        _lib.check(getattr(_lib.lib(), self._run)(self.h, int(B or self.Bmax), _lib.stream_ptr()), self._run)

    def check(self):
        _lib.check(getattr(_lib.lib(), self._check)(self.h, _lib.stream_ptr()), self._check)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                getattr(_lib.lib(), self._destroy)(self.h)
                self.h = None
        except Exception:
            pass


def pack_weight_image_rs(w, c_in, n_out, stride):
    """weight image of csrc/rsconv.cu (include/pv_b200.h): tiles ordered [q][kw][j]; tile (q,kw,j) has nq*n_out
    rows (block b = filter row kh = q + stride*(nq-1-b)) x 16 channels of chunk j; K-major rows of 32 bytes with
    the 32-byte swizzle: element (nn,k) at nn*32 + (((k>>3) ^ ((nn>>2)&1)) * 16) + (k&7)*2."""
    w = torch.as_tensor(w).float()
    Cout, Cin, KH, KW = w.shape
    assert Cout <= n_out and Cin <= c_in and c_in % 16 == 0 and n_out % 16 == 0 and stride in (1, 2)
    kch = c_in // 16
    wp = torch.zeros(n_out, c_in, KH, KW)
    wp[:Cout, :Cin] = w
    parts = []
    for q in range(stride):
        khs = list(range(q, KH, stride))[::-1]            # decreasing kh
        nq = len(khs)
        if nq == 0:
            continue
        nn = torch.arange(nq * n_out)
        swap = ((nn >> 2) & 1).bool()
        for kw in range(KW):
            for j in range(kch):
                # [nq*n_out, 2, 8]: row nn = b*n_out + n, two 16-byte chunks
                tile = torch.stack([wp[:, j * 16:(j + 1) * 16, kh, kw] for kh in khs], dim=0).reshape(nq * n_out, 2, 8)
                t = tile.clone()
                t[swap] = tile[swap].flip(1)              # chunk c is stored at c ^ ((nn>>2)&1)
                parts.append(t.reshape(-1))
    img = torch.cat(parts).to(torch.bfloat16).view(torch.uint8).reshape(-1)
    return img


def unpack_weight_image_rs(img, c_in, n_out, KH, KW, stride):
    """inverse of pack_weight_image_rs, element-address formula written out (CPU test)"""
    kch = c_in // 16
    v = img.view(torch.bfloat16).float().numpy()
    out = np.zeros((n_out, c_in, KH, KW), np.float32)
    base = 0
    for q in range(stride):
        khs = list(range(q, KH, stride))[::-1]
        nq = len(khs)
        for kw in range(KW):
            for j in range(kch):
                for b, kh in enumerate(khs):
                    for n in range(n_out):
                        nn = b * n_out + n
                        for k in range(16):
                            off = base + nn * 32 + (((k >> 3) ^ ((nn >> 2) & 1)) * 16) + (k & 7) * 2
                            out[n, j * 16 + k, kh, kw] = v[off // 2]
                base += nq * n_out * 32
    return out


class RsConv(DetConv):
    """Same layer contract as DetConv on the row-streaming kernel (csrc/rsconv.cu)."""

    _create, _run, _check, _destroy = "pv_rsconv_create", "pv_rsconv_run", "pv_rsconv_check", "pv_rsconv_destroy"
    _instances = INSTANCES_RS

    @staticmethod
    def _pack(weight, c_in, n_out, stride):
        return pack_weight_image_rs(weight, c_in, n_out, stride)

    def info(self):
        a, b, c, e, f = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().pv_rsconv_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(e), C.byref(f)), "pv_rsconv_info")
        return dict(n_stages=a.value, smem_bytes=b.value, strips=c.value, segs=e.value, seg_rows=f.value)


# ---------------------------------------------------------------------------------------------------------------
# c12: detector conv1 + conv2 in one strip kernel (csrc/c12.cu)
# ---------------------------------------------------------------------------------------------------------------
C12_W1_TILES = (("M0", (4, 2, 0)), ("M1", (3, 1)), ("P", ((4, None), (2, 3), (0, 1))))


def _c12_block(w, kind, spec):
    """one 16 x 16 block (output channel n, k) of the conv1 weight image; w float [16,3,5,5]"""
    blk = torch.zeros(16, 16)
    if kind in ("M0", "M1"):
        kh = spec
        for kw in range(4):
            blk[:, kw * 4:kw * 4 + 3] = w[:, :, kh, kw]            # k = kw*4 + c, c = 3 stays zero
    else:
        for half, kh in enumerate(spec):                           # k < 8: even plane row, k >= 8: odd plane row
            if kh is not None:
                blk[:, half * 8:half * 8 + 3] = w[:, :, kh, 4]     # kw = 4; k 3..7 of the half stay zero (kw = 5 does not exist)
    return blk


def pack_c12_w1(w):
    """conv1 weight image of csrc/c12.cu (include/pv_b200.h, PvC12Desc): tiles M0 / M1 / P, blocks of 16 output channels,
    32-byte K-major rows with the 32-byte swizzle counted from the tile's first row."""
    w = torch.as_tensor(w).float()
    assert tuple(w.shape) == (16, 3, 5, 5)
    parts = []
    for kind, specs in C12_W1_TILES:
        tile = torch.cat([_c12_block(w, kind, sp) for sp in specs], dim=0).reshape(-1, 2, 8)      # [nn, chunk, 8]
        nn = torch.arange(tile.shape[0])
        swap = ((nn >> 2) & 1).bool()
        t = tile.clone()
        t[swap] = tile[swap].flip(1)
        parts.append(t.reshape(-1))
    img = torch.cat(parts).to(torch.bfloat16).view(torch.uint8).reshape(-1)
    assert img.numel() == 4096
    return img


def unpack_c12_w1(img):
    """inverse of pack_c12_w1 with the element-address formula written out (CPU test): returns {tile: [blocks][16][16]}"""
    v = img.view(torch.bfloat16).float().numpy()
    out, base = {}, 0
    for kind, specs in C12_W1_TILES:
        nb = len(specs)
        t = np.zeros((nb, 16, 16), np.float32)
        for b in range(nb):
            for n in range(16):
                nn = b * 16 + n
                for k in range(16):
                    off = base + nn * 32 + (((k >> 3) ^ ((nn >> 2) & 1)) * 16) + (k & 7) * 2
                    t[b, n, k] = v[off // 2]
        out[kind] = t
        base += nb * 512
    return out


class FusedC12:
    """Detector conv1 (5x5 s2, RGB -> 16) + conv2 (5x5 s2, 16 -> 32), affine + ReLU each, in one launch reading the RGBA u8
    pyramid plane (csrc/c12.cu).  out: bf16 [B, OH2, even(OW2), 32].  Same run(B) / check() protocol as DetConv."""

    def __init__(self, plane, Hp, Wp, w1, scale1, shift1, w2, scale2, shift2, mean, out=None):
        dev = plane.device
        B = plane.shape[0]
        assert plane.dtype == torch.uint8 and tuple(plane.shape) == (B, Hp, Wp, 4) and plane.is_contiguous()
        self.OH1, self.OW1 = (Hp - 5) // 2 + 1, (Wp - 5) // 2 + 1
        self.OH, self.OW = (self.OH1 - 5) // 2 + 1, (self.OW1 - 5) // 2 + 1
        self.out_pitch = even(self.OW)
        if out is None:
            out = torch.zeros(B, self.OH, self.out_pitch, 32, dtype=torch.bfloat16, device=dev)
        assert tuple(out.shape) == (B, self.OH, self.out_pitch, 32) and out.is_contiguous()
        self.out, self.plane, self.Bmax = out, plane, B
        self.w1_img = pack_c12_w1(w1).to(dev)
        self.w2_img = pack_weight_image_rs(w2, 16, 32, 2).to(dev)
        f = lambda a: torch.as_tensor(a).float().contiguous().to(dev)
        self.scale1, self.shift1, self.scale2, self.shift2 = f(scale1), f(shift1), f(scale2), f(shift2)
        self._mean = (C.c_float * 3)(*[float(m) for m in mean])
        d = _lib.PvC12Desc()
        d.plane = plane.data_ptr()
        d.B, d.Hp, d.Wp = B, Hp, Wp
        d.w1_img, d.w1_bytes = self.w1_img.data_ptr(), self.w1_img.numel()
        d.w2_img, d.w2_bytes = self.w2_img.data_ptr(), self.w2_img.numel()
        d.scale1, d.shift1 = self.scale1.data_ptr(), self.shift1.data_ptr()
        d.scale2, d.shift2 = self.scale2.data_ptr(), self.shift2.data_ptr()
        d.out, d.out_pitch = out.data_ptr(), self.out_pitch
        d.mean_host = self._mean
        h = C.c_void_p()
        _lib.check(_lib.lib().pv_c12_create(C.byref(d), C.byref(h)), "pv_c12_create")
        self.h = h
        self.flops_per_image = 2 * (self.OH1 * self.OW1 * 16 * 3 * 25 + self.OH * self.OW * 32 * 16 * 25)

    def info(self):
        v = [C.c_int() for _ in range(6)]
        _lib.check(_lib.lib().pv_c12_info(self.h, *[C.byref(x) for x in v]), "pv_c12_info")
        return dict(zip(("smem_bytes", "strips", "segs", "seg_rows", "oh2", "ow2"), [x.value for x in v]))

    def debug(self):
        out = (C.c_longlong * 16)()
        _lib.check(_lib.lib().pv_c12_debug(self.h, out), "pv_c12_debug")
        names = ("c1_wait_px", "c1_issue_tile1", "c1_issue", "c1_quads", "c2_wait_a", "c2_wait_slot", "c2_issue", "c2_rows",
                 "cv_wait_raw", "cv_wait_slot", "cv_work", "e1_wait_row", "e1_wait_a", "e1_work", "e2_wait_row", "e2_work")
        return dict(zip(names, [int(x) for x in out]))

    def run(self, B=None):
        _lib.check(_lib.lib().pv_c12_run(self.h, int(B or self.Bmax), _lib.stream_ptr()), "pv_c12_run")

    def check(self):
        _lib.check(_lib.lib().pv_c12_check(self.h, _lib.stream_ptr()), "pv_c12_check")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().pv_c12_destroy(self.h)
                self.h = None
        except Exception:
            pass
