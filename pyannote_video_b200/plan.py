"""Host-side planning for the srgemm kernel: activation row layouts and conv -> tap tables.

An activation tensor lives in HBM as a row-major bf16 matrix ("row layout"), one row per
pixel position, so that a convolution becomes  D[q] = sum_t X[q + off_t] . W_t  (see
include/pv_b200.h).  Three layouts exist:

  padded   [B, H+2p, W+2p, C]            zero border of width p, consumed by stride-1 convs
  parity   [4, B, ceil(Hp/2), ceil(Wp/2), C]   the 4 (row,col)-parity planes of the padded
                                          tensor, consumed by stride-2 convs and 2x2 pools
  gathered [2, B, ceil(H/2), ceil(W/2), kw*3 -> 16|32]   first-layer input: each row holds the
                                          kw RGB pixels a stride-2 filter row needs

The layer graphs follow dlib's `mmod_human_face_detector` and `face_recognition_resnet_model_v1`
network definitions as restated in SURVEY.md App. A.1/A.4 (reference call sites:
pyannote/video/face/face.py:66 and :74-75).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def round_up(x, m):
    return (x + m - 1) // m * m


class RowLayout:
    def __init__(self, kind, B, H, W, C, pad=0, kw=None):
        assert kind in ("padded", "parity", "gathered", "pixrows")
        self.kind, self.B, self.H, self.W, self.C, self.pad, self.kw = kind, B, H, W, C, pad, kw
        self.row_stride_bytes = 0
        if kind == "padded":
            self.Hq, self.Wq = H + 2 * pad, W + 2 * pad
            self.planes = 1
            self.cols = C
        elif kind == "parity":
            self.Hq, self.Wq = (H + 2 * pad + 1) // 2, (W + 2 * pad + 1) // 2
            self.planes = 4
            self.cols = C
        elif kind == "pixrows":
            # first-layer input read in place: bf16 RGBX pixels (8 B) of the two row-parity half
            # planes; matrix row j = the 8 pixels starting at pixel pair j (rows overlap: stride 16 B)
            assert pad == 0 and C == 3 and kw is not None and kw <= 8 and W % 2 == 0
            self.Hq, self.Wq = (H + 1) // 2, W // 2
            self.planes = 2
            self.cols = 32
            self.row_stride_bytes = 16
        else:
            assert pad == 0 and kw is not None
            self.Hq, self.Wq = (H + 1) // 2, (W + 1) // 2
            self.planes = 2
            self.cols = 16 if kw * C <= 16 else 32
            assert kw * C <= 32
        assert self.cols % 8 == 0
        self.img = self.Hq * self.Wq
        self.plane_rows = B * self.img
        self.rows = self.planes * self.plane_rows

    # ---- device-side row map (destination of an epilogue) ----
    def rowmap(self):
        assert self.kind in ("padded", "parity")
        m = _lib.PvRowMap()
        m.kind = 0 if self.kind == "padded" else 1
        m.cols = self.cols
        m.w = self.Wq
        m.py = m.px = self.pad
        m.img = self.img
        m.plane_rows = self.plane_rows
        return m

    def alloc(self, device):
        if self.kind == "pixrows":
            # physical buffer: [2, B, Hq, W] pixels of 4 bf16 + 8 pixels of zero slack
            return torch.zeros(self.rows * 8 + 32, dtype=torch.bfloat16, device=device)
        return torch.zeros(self.rows, self.cols, dtype=torch.bfloat16, device=device)

    # ---- CPU reference scatter/gather (tests, emulator) ----
    def row_index(self, n, y, x):
        """row of interior position (n,y,x) (numpy broadcasting)"""
        Y, X = y + self.pad, x + self.pad
        if self.kind == "padded":
            return n * self.img + Y * self.Wq + X
        assert self.kind == "parity"
        plane = (Y & 1) * 2 + (X & 1)
        return plane * self.plane_rows + n * self.img + (Y >> 1) * self.Wq + (X >> 1)

    def to_rows(self, x_nhwc):
        """NHWC float tensor [B,H,W,C'] (C' <= C) -> bf16 row matrix (CPU)."""
        B, H, W, Cc = x_nhwc.shape
        assert (B, H, W) == (self.B, self.H, self.W)
        out = torch.zeros(self.rows, self.cols, dtype=torch.float32)
        if self.kind == "pixrows":
            phys = torch.zeros(self.rows * 8 + 32)
            v = phys[:self.rows * 8].view(2, B, self.Hq, W, 4)
            for ph in range(2):
                rows_ph = x_nhwc[:, ph::2]
                v[ph, :, :rows_ph.shape[1], :, :Cc] = rows_ph
            idx = (torch.arange(self.rows)[:, None] * 8 + torch.arange(32)[None, :])
            return phys[idx].to(torch.bfloat16)
        if self.kind == "gathered":
            xp = torch.zeros(B, 2 * self.Hq, 2 * self.Wq + self.kw, Cc)
            xp[:, :H, :W] = x_nhwc
            for ph in range(2):
                rowsel = xp[:, ph::2][:, :self.Hq]  # [B,Hq,Wtot,C]
                cols = [rowsel[:, :, k:k + 2 * self.Wq:2] for k in range(self.kw)]  # each [B,Hq,Wq,C]
                g = torch.stack(cols, dim=3).reshape(B, self.Hq, self.Wq, self.kw * Cc)
                out[ph * self.plane_rows:(ph + 1) * self.plane_rows, :self.kw * Cc] = g.reshape(-1, self.kw * Cc)
            return out.to(torch.bfloat16)
        n, y, x = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
        r = torch.from_numpy(self.row_index(n, y, x).reshape(-1).astype(np.int64))
        out[r, :Cc] = x_nhwc.reshape(-1, Cc).float()
        return out.to(torch.bfloat16)

    def from_rows(self, mat, C=None):
        """row matrix -> NHWC interior [B,H,W,C] float32 (CPU)."""
        assert self.kind in ("padded", "parity")
        C = C or self.C
        n, y, x = np.meshgrid(np.arange(self.B), np.arange(self.H), np.arange(self.W), indexing="ij")
        r = torch.from_numpy(self.row_index(n, y, x).reshape(-1).astype(np.int64))
        return mat.detach().cpu().float()[r, :C].reshape(self.B, self.H, self.W, C)


def conv_out_size(n, k, s, p):
    return (n + 2 * p - k) // s + 1


def _segments(kc):
    """split a row of kc elements into TMA segments of 64/32/16 elements (<= 2 distinct widths)"""
    segs = []
    col = 0
    while kc - col >= 64:
        segs.append((col, 64))
        col += 64
    if kc - col >= 32:
        segs.append((col, 32))
        col += 32
    if kc - col >= 16:
        segs.append((col, 16))
        col += 16
    assert col == kc, "row length %d is not a multiple of 16" % kc
    widths = sorted({w for _, w in segs}, reverse=True)
    assert len(widths) <= 2, "row length %d needs more than two segment widths" % kc
    return segs, widths


RESIDENT_MAX_BYTES = 120 * 1024     # weights at most this large stay in shared memory for the launch
SLOT_TARGET_BYTES = 40 * 1024        # operands per ring slot (one mbarrier round trip)
SMEM_AVAILABLE = 227 * 1024 - 16 * 1024


class ConvPlan:
    """Tap table + packed weights for one convolution over a RowLayout."""

    def __init__(self, lin, weight, stride, pad, group="tap", max_b_bytes=40 * 1024, resident_max=RESIDENT_MAX_BYTES,
                 ctas_per_sm=0):
        """weight: float tensor [Cout, Cin, KH, KW] (dlib/torch order), CPU."""
        self.ctas_req = ctas_per_sm
        Cout, Cin, KH, KW = weight.shape
        self.lin, self.stride, self.pad, self.KH, self.KW = lin, stride, pad, KH, KW
        N = round_up(Cout, 16)
        Kc = lin.cols
        OH = conv_out_size(lin.H, KH, stride, pad)
        OW = conv_out_size(lin.W, KW, stride, pad)
        taps = []  # (row offset, W_t [N, Kc])
        w = weight.float()
        if lin.kind == "padded":
            assert stride == 1 and lin.pad >= pad and Cin <= lin.C
            d = lin.pad - pad
            for kh in range(KH):
                for kw in range(KW):
                    wt = torch.zeros(N, Kc)
                    wt[:Cout, :Cin] = w[:, :, kh, kw]
                    taps.append(((kh + d) * lin.Wq + (kw + d), wt))
        elif lin.kind == "parity":
            assert stride == 2 and lin.pad >= pad and Cin <= lin.C
            d = lin.pad - pad
            for kh in range(KH):
                for kw in range(KW):
                    wt = torch.zeros(N, Kc)
                    wt[:Cout, :Cin] = w[:, :, kh, kw]
                    plane = ((kh + d) & 1) * 2 + ((kw + d) & 1)
                    taps.append((plane * lin.plane_rows + ((kh + d) >> 1) * lin.Wq + ((kw + d) >> 1), wt))
        elif lin.kind == "pixrows":
            assert stride == 2 and pad == 0 and KW == lin.kw and Cin == 3
            for kh in range(KH):
                wt = torch.zeros(N, 8, 4)
                wt[:Cout, :KW, :3] = w[:, :, kh, :].permute(0, 2, 1)     # [n][kw][c]
                taps.append(((kh & 1) * lin.plane_rows + (kh >> 1) * lin.Wq, wt.reshape(N, 32)))
        else:
            assert stride == 2 and pad == 0 and KW == lin.kw and Cin == lin.C
            for kh in range(KH):
                wt = torch.zeros(N, Kc)
                # gathered row = [kw][c]
                wt[:Cout, :KW * Cin] = w[:, :, kh, :].permute(0, 2, 1).reshape(Cout, KW * Cin)
                taps.append(((kh & 1) * lin.plane_rows + (kh >> 1) * lin.Wq, wt))
        self._build(taps, Cout, OH, OW, group, max_b_bytes, resident_max)

    @classmethod
    def from_taps(cls, lin, taps, Cout, OH, OW, group="tap", max_b_bytes=40 * 1024, resident_max=RESIDENT_MAX_BYTES,
                  kernel=(1, 1), ctas_per_sm=0):
        """taps: list of (row offset, W_t float [round_up(Cout,16), lin.cols])"""
        self = cls.__new__(cls)
        self.ctas_req = ctas_per_sm
        self.lin, self.stride, self.pad = lin, 1, 0
        self.KH, self.KW = kernel
        self._build(list(taps), Cout, OH, OW, group, max_b_bytes, resident_max)
        return self

    def _build(self, taps, Cout, OH, OW, group, max_b_bytes, resident_max):
        lin = self.lin
        self.Cout, self.OH, self.OW = Cout, OH, OW
        self.N = N = round_up(Cout, 16)
        self.n_taps_total = len(taps)
        self.Kc = Kc = lin.cols
        segs, widths = _segments(Kc)
        self.widths = widths
        w_bytes = len(taps) * N * Kc * 2
        self.resident = w_bytes <= resident_max
        self.resident_max = resident_max if self.resident else 0
        # ---- group taps that share a slab ----
        taps.sort(key=lambda t: t[0])
        if group == "tap":
            max_span = 0
        elif group == "row":
            max_span = 8  # taps of one filter row (kw offsets)
        elif group == "all":
            max_span = 120
        else:
            max_span = int(group)
        maxw = max(widths)
        max_taps = _lib.PV_SR_MAX_TAPS if self.resident else max(1, min(_lib.PV_SR_MAX_TAPS, max_b_bytes // (N * maxw * 2)))
        groups = []
        for off, wt in taps:
            if groups and off - groups[-1][0][0] <= max_span and len(groups[-1]) < max_taps:
                groups[-1].append((off, wt))
            else:
                groups.append([(off, wt)])
        span = max(g[-1][0] - g[0][0] for g in groups)
        self.tail_rows = round_up(span, 8)
        assert self.tail_rows <= 128
        # ---- entries + packed weights ----
        entries = []
        packed = {wd: [] for wd in widths}
        nrows = {wd: 0 for wd in widths}
        for g in groups:
            base = g[0][0]
            for col, wd in segs:
                st = _lib.PvSrEntry()
                st.a_row_off = base
                st.b_row = nrows[wd]
                st.a_col = col
                st.cls = widths.index(wd)
                st.n_taps = len(g)
                st.use_tail = 1 if self.tail_rows > 0 else 0
                for i, (off, wt) in enumerate(g):
                    st.tap_rel[i] = off - base
                    packed[wd].append(wt[:, col:col + wd])
                    nrows[wd] += N
                entries.append(st)
        assert len(entries) <= _lib.PV_SR_MAX_ENTRIES, "too many entries: %d" % len(entries)
        # ---- co-resident CTAs per SM and ring-slot packing ----
        res_bytes = round_up(w_bytes, 1024) + 1024 if self.resident else 0
        n_mma = sum(st.n_taps * (widths[st.cls] // 16) for st in entries)

        def ebytes(st):
            wd = widths[st.cls]
            b = round_up((128 + (self.tail_rows if st.use_tail else 0)) * wd * 2, 1024)
            if not self.resident:
                b += round_up(st.n_taps * N * wd * 2, 1024)
            return b

        max_entry = max(ebytes(st) for st in entries)
        fixed = 52 * len(entries) + 16 * n_mma + 8 * N + 512

        def per_cta(cps):
            return (227 * 1024) // cps - 2048 - fixed - res_bytes

        cands = [self.ctas_req] if self.ctas_req else ([4, 2, 1] if N <= 64 else ([2, 1] if N <= 128 else [1]))
        cps = 1
        for c in cands:
            if per_cta(c) >= 2 * max_entry and 2 * N <= 512 // c:
                cps = c
                break
        self.ctas_per_sm = cps
        slot_cap = max(max_entry, min(SLOT_TARGET_BYTES // cps, per_cta(cps) // 3))

        cur = 0
        for i, st in enumerate(entries):
            eb = ebytes(st)
            if i == 0 or cur + eb > slot_cap:
                st.flags = 1
                if i > 0:
                    entries[i - 1].flags |= 2
                cur = 0
            cur += eb
        entries[-1].flags |= 2
        self.entries = entries
        self.n_slots = sum(1 for st in entries if st.flags & 1)
        self.w_packed = [torch.cat(packed[wd], dim=0).to(torch.bfloat16).contiguous() for wd in widths]
        self.mma_per_tile = sum(st.n_taps * (widths[st.cls] // 16) for st in entries)

    @property
    def stages(self):
        return self.entries

    def flops_per_row(self):
        return 2 * self.N * 16 * self.mma_per_tile


class Srgemm:
    """A bound srgemm plan: conv + fused affine/residual/ReLU epilogue between device buffers."""

    def __init__(self, cp, x, out, lout, scale, shift, relu, resid=None, lres=None, out_f32=False,
                 out_rows_f32=False, max_ctas=0, x_rows=None, x_row_stride_bytes=0, acc_split=0, ctas_per_sm=None, mma_warps=None):
        lin = cp.lin
        dev = x.device
        if lin.kind == "pixrows":
            x_rows, x_row_stride_bytes = lin.rows, lin.row_stride_bytes
            assert x.dtype == torch.bfloat16 and x.numel() == lin.rows * 8 + 32 and x.is_contiguous()
        if x_rows is None:
            assert x.dtype == torch.bfloat16 and x.shape == (lin.rows, lin.cols) and x.is_contiguous()
        self.cp = cp
        self.keep = [x, out, resid]
        self.w_dev = [w.to(dev) for w in cp.w_packed]
        N = cp.N
        sc = torch.zeros(N, dtype=torch.float32)
        sh = torch.zeros(N, dtype=torch.float32)
        sc[:cp.Cout] = scale.float()
        sh[:cp.Cout] = shift.float()
        self.scale, self.shift = sc.to(dev), sh.to(dev)
        d = _lib.PvSrgemmDesc()
        d.x = x.data_ptr()
        d.x_rows, d.x_cols = (x_rows if x_rows is not None else lin.rows), lin.cols
        d.x_row_stride_bytes = x_row_stride_bytes
        d.weights_resident_max_bytes = cp.resident_max
        d.n_out = N
        d.n_classes = len(cp.widths)
        for i, wd in enumerate(cp.widths):
            d.class_width[i] = wd
            d.w_packed[i] = self.w_dev[i].data_ptr()
            d.w_rows[i] = self.w_dev[i].shape[0]
        d.tail_rows = cp.tail_rows
        d.n_entries = len(cp.entries)
        self._entries = (_lib.PvSrEntry * len(cp.entries))(*cp.entries)
        d.entries = C.cast(self._entries, C.POINTER(_lib.PvSrEntry))
        d.scale, d.shift = self.scale.data_ptr(), self.shift.data_ptr()
        d.hq, d.wq, d.oh, d.ow = lin.Hq, lin.Wq, cp.OH, cp.OW
        d.relu = int(relu)
        d.out = out.data_ptr()
        if out_f32:
            assert out.dtype == torch.float32
            d.out_mode = 2
            m = _lib.PvRowMap()
            m.kind, m.cols, m.w, m.py, m.px, m.img, m.plane_rows = 0, 1, cp.OW, 0, 0, cp.OH * cp.OW, 0
            d.dst = m
        elif out_rows_f32:
            assert out.dtype == torch.float32 and out.shape == (lout.rows, lout.cols)
            d.out_mode = 3
            d.dst = lout.rowmap()
        else:
            assert out.dtype == torch.bfloat16 and out.shape == (lout.rows, lout.cols)
            assert (lout.H, lout.W) == (cp.OH, cp.OW) or (lout.H >= cp.OH and lout.W >= cp.OW)
            d.out_mode = 0
            d.dst = lout.rowmap()
        if resid is not None:
            assert resid.dtype == torch.bfloat16 and resid.shape == (lres.rows, lres.cols)
            d.resid = resid.data_ptr()
            d.res = lres.rowmap()
        d.max_ctas = max_ctas
        d.acc_split = acc_split
        d.ctas_per_sm = cp.ctas_per_sm if ctas_per_sm is None else ctas_per_sm
        if mma_warps is None:
            # aim for 4 MMA-issuing lanes per SM: co-resident CTAs x issuing warps per CTA.  Every warp must
            # own at least one MMA in every ring slot, and the accumulators (warps x N) must leave room
            # for two tiles in TMEM.
            per_slot, cur = [], 0
            for st in cp.entries:
                cur += st.n_taps * (cp.widths[st.cls] // 16)
                if st.flags & 2:
                    per_slot.append(cur)
                    cur = 0
            mma_warps = 1
            for w in (4, 2):
                if w * d.ctas_per_sm <= 4 and min(per_slot) >= w and 2 * w * cp.N <= 512 // d.ctas_per_sm and cp.N <= 64:
                    mma_warps = w
                    break
        d.mma_warps = mma_warps
        self.mma_warps = mma_warps
        self.desc = d
        self.q_rows = lin.plane_rows
        h = C.c_void_p()
        _lib.check(_lib.lib().pv_srgemm_create(C.byref(d), C.byref(h)), "pv_srgemm_create")
        self.h = h

    def info(self):
        a, b, c, e, f = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().pv_srgemm_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(e), C.byref(f)),
                   "pv_srgemm_info")
        return dict(n_ring=a.value, slot_bytes=b.value, resident=c.value, n_acc=e.value, acc_split=f.value,
                    mma_warps=self.mma_warps)

    def run(self, q_rows=None):
        _lib.check(_lib.lib().pv_srgemm_run(self.h, C.c_int64(q_rows or self.q_rows), _lib.stream_ptr()),
                   "pv_srgemm_run")

    def check(self):
        _lib.check(_lib.lib().pv_srgemm_check(self.h, _lib.stream_ptr()), "pv_srgemm_check")

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().pv_srgemm_destroy(self.h)
                self.h = None
        except Exception:
            pass
