"""Device-resident execution plans for the two CNNs on the path (host orchestration only).

  DetectorNet  dlib CNN/MMOD face detector: tiled image pyramid -> 7 convs -> decode + NMS
               (replaces face_detector_(rgb, 1), pyannote/video/face/face.py:66)
  EmbedNet     dlib face_recognition_resnet_model_v1: 150x150 chip -> 29 convs -> 128-d
               (replaces compute_face_descriptor, pyannote/video/face/face.py:74-75)

Every convolution is one tcgen05 kernel launch with the dlib `affine` (frozen BN), bias, residual add and
ReLU fused into its epilogue — `conv1_fused` + `rsconv` (row streaming) for the detector, `srgemm` (shifted-row
GEMM) for the embedder; the remaining layers are the small kernels of csrc/layers.cu and csrc/detect.cu.  All buffers are allocated once; nothing here touches the CPU
oracle.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, config
from . import weights as W
from .plan import ConvPlan, RowLayout, Srgemm
from .detconv import DetConv, RsConv, FusedC12, even
from .pyrgeom import pyramid_geometry, det_cell_to_plane


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _affine(c, affine=True):
    g, b, be = _t(c["gamma"]).float(), _t(c["b"]).float(), _t(c["beta"]).float()
    if affine:
        return g, g * b + be
    return torch.ones_like(b), b


def _mean3():
    return (C.c_float * 3)(*W.PIXEL_MEAN)


def _packed_rowmap(F, H, Wp, P, C):
    """PvRowMap kind 2: faces side by side in an image row (csrc/layers.cu row_of)"""
    rm = _lib.PvRowMap()
    rm.kind, rm.cols, rm.w, rm.py, rm.px, rm.img, rm.plane_rows = 2, C, Wp, H, P, F, 0
    return rm


class EmbedNet:
    """face_recognition_resnet_model_v1 (29 convs).  impl 'rs' (default): the 3x3 layers of levels 4 and 3 (14 convs,
    54 % of the FLOPs, 32 / 64 channels) run on the row-streaming kernel csrc/rsconv.cu with FACES_PER_ROW faces packed
    side by side in one image row (one zero column between neighbours = the convs' padding), so that one MMA multiplies
    an input row against the three filter rows at once (N = 96 / 192 instead of 32 / 64); conv1 and levels 2..0 stay on
    the shifted-row GEMM csrc/srgemm.cu.  impl 'srgemm': every conv on srgemm (the round-1 path, kept as cross-check)."""

    FACES_PER_ROW = 7          # 7 x (35 + 1) = 252 columns = two 128-column strips (126 used each); 7 x 18 = 126 at level 3

    def __init__(self, model, max_batch, device, group=None, impl=None):
        if model.get("kind") != "resnet_v1_embedder":
            raise RuntimeError("EmbedNet: not an embedder model")
        group = config.SRGEMM_GROUP if group is None else group
        self.impl = impl = (config.EMBED_IMPL if impl is None else impl)
        self.B = B = int(max_batch)
        self.dev = device
        S = W.EMB_CHIP
        self.chips = torch.zeros(B, S, S, 4, dtype=torch.uint8, device=device)
        self.ops = []          # (kind, payload)
        self.flops_per_face = 0
        # conv1 7x7 s2 on the gathered layout
        lg = RowLayout("gathered", B, S, S, 3, kw=7)
        self.lg, self.xg = lg, lg.alloc(device)
        w1 = _t(model["conv1"]["w"])
        cp = ConvPlan(lg, w1, 2, 0, group=group)
        l1 = RowLayout("padded", B, cp.OH, cp.OW, 32, pad=0)
        b1 = l1.alloc(device)
        sc, sh = _affine(model["conv1"])
        self._add_conv(cp, self.xg, b1, l1, sc, sh, True, None, None)
        # max_pool 3x3 s2
        H = (cp.OH - 3) // 2 + 1
        blocks = model["blocks"]
        first = 0
        if impl == "rs":
            cur, lcur, first = self._build_rs(model, b1, l1, H, device)
        else:
            lcur = RowLayout("parity" if blocks[0]["type"] == "ares_down" else "padded", B, H, H, 32,
                             pad=0 if blocks[0]["type"] == "ares_down" else 1)
            cur = lcur.alloc(device)
            self.ops.append(("maxpool", (b1, cur, l1.H, l1.W, 32, lcur.rowmap())))
        for i, blk in enumerate(blocks):
            if i < first:
                continue
            nxt = blocks[i + 1]["type"] if i + 1 < len(blocks) else None
            ch = blk["a"]["w"].shape[0]
            wa, wb = _t(blk["a"]["w"]), _t(blk["b"]["w"])
            sca, sha = _affine(blk["a"])
            scb, shb = _affine(blk["b"])

            def out_layout(h, w):
                if nxt is None:
                    return RowLayout("padded", B, h, w, ch, pad=0)
                if nxt == "ares_down":
                    return RowLayout("parity", B, h, w, ch, pad=0)
                return RowLayout("padded", B, h, w, ch, pad=1)

            if blk["type"] == "ares":
                assert lcur.kind == "padded" and lcur.pad == 1
                la = RowLayout("padded", B, lcur.H, lcur.W, ch, pad=1)
                t = la.alloc(device)
                self._add_conv(ConvPlan(lcur, wa, 1, 1, group=group), cur, t, la, sca, sha, True, None, None)
                lo = out_layout(lcur.H, lcur.W)
                o = lo.alloc(device)
                self._add_conv(ConvPlan(la, wb, 1, 1, group=group), t, o, lo, scb, shb, True, cur, lcur)
            else:
                assert lcur.kind == "parity" and lcur.pad == 0
                cpa = ConvPlan(lcur, wa, 2, 0, group=group)
                la = RowLayout("padded", B, cpa.OH, cpa.OW, ch, pad=1)
                t = la.alloc(device)
                self._add_conv(cpa, cur, t, la, sca, sha, True, None, None)
                ph, pw = (lcur.H - 2) // 2 + 1, (lcur.W - 2) // 2 + 1
                ho, wo = max(ph, cpa.OH), max(pw, cpa.OW)
                assert (ho, wo) == (ph, pw)
                lo = out_layout(ho, wo)
                o = lo.alloc(device)
                skip = lo.alloc(device)
                self.ops.append(("avgpool", (cur, lcur, skip, o, ph, pw, ch, lo.rowmap())))
                self._add_conv(ConvPlan(la, wb, 1, 1, group=group), t, o, lo, scb, shb, True, skip, lo)
            cur, lcur = o, lo
        assert lcur.kind == "padded" and lcur.pad == 0
        self.fc = _t(model["fc"]).float().contiguous().to(device)
        self.out = torch.zeros(B, W.EMB_DIM, dtype=torch.float32, device=device)
        self.ops.append(("head", (cur, lcur.H * lcur.W, lcur.C)))
        self.final_layout = lcur

    # ---- levels 4 and 3 on rsconv, faces packed side by side ----
    def _build_rs(self, model, b1, l1, H, device):
        """returns (tensor, layout, index of the first block that is NOT handled here)"""
        B, F = self.B, self.FACES_PER_ROW
        G = (B + F - 1) // F
        self.G, blocks = G, model["blocks"]

        def pr(h, w, c):
            P = w + 1
            wp = even(F * P)
            return torch.zeros(G, h, wp, c, dtype=torch.bfloat16, device=device), wp, P

        def rs(x, h, wp, conv, stride, c_in, n_out, gap, out=None, resid=None, face=None):
            sc, sh = _affine(conv)
            op = RsConv(x, h, wp, _t(conv["w"]), stride, sc, sh, True, c_in, n_out, out=out, resid=resid, gap=gap)
            fh, fw = face
            cout, cin = conv["w"].shape[0], conv["w"].shape[1]
            fl = 2 * fh * fw * cout * cin * 9
            self.ops.append(("rsconv", (op, fl)))
            self.flops_per_face += fl
            return op.out

        # max_pool: conv1 output rows -> level-4 packed tensor
        x, wp, P = pr(H, H, 32)
        self.ops.append(("maxpool", (b1, x, l1.H, l1.W, 32, _packed_rowmap(F, H, wp, P, 32))))
        h, w, c = H, H, 32
        i = 0
        while i < len(blocks):
            blk = blocks[i]
            ch = blk["a"]["w"].shape[0]
            if ch > 64:
                break
            if blk["type"] == "ares":
                gap = (P, w)
                t = rs(x, h, wp, blk["a"], 1, c, ch, gap, face=(h, w))
                x = rs(t, h, wp, blk["b"], 1, ch, ch, gap, resid=x, face=(h, w))
            else:
                oh, ow = (h - 3) // 2 + 1, (w - 3) // 2 + 1
                skip, owp, OP = pr(oh, ow, ch)
                assert OP * 2 == P and (oh, ow) == ((h - 2) // 2 + 1, (w - 2) // 2 + 1)
                gap = (OP, ow)
                t = rs(x, h, wp, blk["a"], 2, c, ch, gap, face=(oh, ow))
                assert tuple(t.shape) == tuple(skip.shape), (t.shape, skip.shape)
                self.ops.append(("pr_avgpool", (x, G, F, h, w, wp, c, skip, owp, ch)))
                x = rs(t, oh, owp, blk["b"], 1, ch, ch, gap, resid=skip, face=(oh, ow))
                h, w, c, wp, P = oh, ow, ch, owp, OP
            i += 1
        # hand over to the srgemm layers: the layout the next block reads
        nxt = blocks[i]["type"] if i < len(blocks) else None
        if nxt is None:
            lo = RowLayout("padded", B, h, w, c, pad=0)
        elif nxt == "ares_down":
            lo = RowLayout("parity", B, h, w, c, pad=0)
        else:
            lo = RowLayout("padded", B, h, w, c, pad=1)
        o = lo.alloc(device)
        self.ops.append(("pr_unpack", (x, F, h, w, wp, c, o, lo.rowmap())))
        return o, lo, i

    def _add_conv(self, cp, x, out, lout, sc, sh, relu, resid, lres):
        op = Srgemm(cp, x, out, lout, sc, sh, relu, resid=resid, lres=lres)
        self.ops.append(("conv", (op, cp.lin.img)))
        Cin = cp.lin.C if cp.lin.kind != "gathered" else 3
        self.flops_per_face += 2 * cp.OH * cp.OW * cp.Cout * Cin * cp.KH * cp.KW

    def conv_ops(self):
        """[(op, flops per face)] of every tensor-core conv launch, in execution order (bench / probes patch op.run)"""
        out = []
        for kind, a in self.ops:
            if kind == "conv":
                cp = a[0].cp
                cin = cp.lin.C if cp.lin.kind != "gathered" else 3
                out.append((a[0], 2 * cp.OH * cp.OW * cp.Cout * cin * cp.KH * cp.KW))
            elif kind == "rsconv":
                out.append(a)
        return out

    def forward_chips(self, M):
        """Run the network on self.chips[:M] (RGBA u8).  Returns self.out[:M] (float32 [M,128])."""
        assert 0 < M <= self.B
        L = _lib.lib()
        st = _lib.stream_ptr()
        S = W.EMB_CHIP
        Gm = (M + self.FACES_PER_ROW - 1) // self.FACES_PER_ROW
        _lib.check(L.pv_pack_gathered(_lib.ptr(self.chips), _lib.ptr(self.xg), M, S, S, 7,
                                      C.c_int64(self.lg.plane_rows), _mean3(), st), "pv_pack_gathered")
        for kind, a in self.ops:
            if kind == "conv":
                op, img = a
                op.run(M * img)
            elif kind == "rsconv":
                a[0].run(Gm)
            elif kind == "maxpool":
                src, dst, h, w, c, rm = a
                _lib.check(L.pv_maxpool3x3s2(_lib.ptr(src), _lib.ptr(dst), M, h, w, c, C.byref(rm), st), "pv_maxpool3x3s2")
            elif kind == "avgpool":
                src, lsrc, skip, o, ph, pw, ch, rm = a
                _lib.check(L.pv_avgpool_skip(_lib.ptr(src), lsrc.C, C.c_int64(lsrc.plane_rows), lsrc.Hq, lsrc.Wq,
                                             _lib.ptr(skip), _lib.ptr(o), M, ph, pw, ch, C.byref(rm), st),
                           "pv_avgpool_skip")
            elif kind == "pr_avgpool":
                src, G, F, h, w, wp, c, skip, owp, ch = a
                _lib.check(L.pv_pr_avgpool(_lib.ptr(src), Gm, F, h, w, wp, c, _lib.ptr(skip), owp, ch, st), "pv_pr_avgpool")
            elif kind == "pr_unpack":
                src, F, h, w, wp, c, o, rm = a
                _lib.check(L.pv_pr_unpack(_lib.ptr(src), M, F, h, w, wp, c, _lib.ptr(o), C.byref(rm), st), "pv_pr_unpack")
            else:
                src, hw, c = a
                _lib.check(L.pv_embed_head(_lib.ptr(src), M, hw, c, _lib.ptr(self.fc), _lib.ptr(self.out),
                                           W.EMB_DIM, st), "pv_embed_head")
        return self.out[:M]

    def check(self):
        for kind, a in self.ops:
            if kind in ("conv", "rsconv"):
                a[0].check()


class FusedConv1:
    """First detector conv (5x5, stride 2, 3 -> 16) reading the RGBA u8 plane in place
    (csrc/conv1_fused.cu); same `run(q_rows)` / `check()` protocol as Srgemm."""

    def __init__(self, plane, Hp, Wp, conv, out, lout, device):
        w = _t(conv["w"]).float()                                  # [16, 3, 5, 5]
        assert tuple(w.shape) == (16, 3, 5, 5)
        self.w = w.to(torch.bfloat16).contiguous().to(device)     # the kernel lays out its own MMA operands
        sc, sh = _affine(conv)
        self.scale, self.shift = sc.float().contiguous().to(device), sh.float().contiguous().to(device)
        self.plane, self.Hp, self.Wp, self.out, self.lout = plane, Hp, Wp, out, lout
        self.rm = lout.rowmap()
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.OH, self.OW = (Hp - 5) // 2 + 1, (Wp - 5) // 2 + 1
        self.img = ((Hp + 1) // 2) * ((Wp + 1) // 2)

    def run(self, q_rows=None):
        M = (q_rows // self.img) if q_rows else self.plane.shape[0]
        _lib.check(_lib.lib().pv_conv1_fused(_lib.ptr(self.plane), M, self.Hp, self.Wp, _lib.ptr(self.w), _lib.ptr(self.scale),
                                             _lib.ptr(self.shift), 1, _lib.ptr(self.out), C.byref(self.rm), self.OH, self.OW,
                                             _mean3(), _lib.ptr(self.err), _lib.stream_ptr()), "pv_conv1_fused")

    def check(self):
        torch.cuda.synchronize()
        code = int(self.err.item())
        if code:
            self.err.zero_()
            raise _lib.PvError("conv1_fused: device-side pipeline timeout (role code %d)" % code)


class DetectorNet:
    """Batched CNN detector for frames of one size."""

    MAX_CAND = 4096
    MAX_DET = 256
    TAIL_PIXELS = 250000      # pyramid levels at most this large are built by the one-launch tail kernel

    def __init__(self, model, H, W_, upsample, max_batch, device, group=None, conv1_mode=None, conv_impl=None):
        if model.get("kind") != "mmod_detector":
            raise RuntimeError("DetectorNet: not a detector model")
        group = config.SRGEMM_GROUP if group is None else group
        self.model = model
        self.B = B = int(max_batch)
        self.H, self.W, self.upsample, self.dev = H, W_, int(upsample), device
        self.geo = geo = pyramid_geometry(H, W_, upsample)
        Hp, Wp = geo.plane_h, geo.plane_w
        self.plane = torch.zeros(B, Hp, Wp, 4, dtype=torch.uint8, device=device)
        self.planes = [self.plane]        # a second plane (enable_double_buffer) lets the next batch's pyramid be built early
        self.conv1_mode = config.DET_CONV1 if conv1_mode is None else conv1_mode
        self.conv_impl = config.DET_CONVS if conv_impl is None else conv_impl
        if self.conv1_mode == "c12" and self.conv_impl != "rsconv":
            self.conv1_mode = "fused"            # the conv1+conv2 strip kernel feeds the rsconv layers only
        if self.conv_impl in ("detconv", "rsconv"):
            self._init_detconv(model, device)
        else:
            self._init_srgemm(model, device, group)
        self._init_tail(model, device)
        # algorithmic work: MACs per PYRAMID pixel (not per plane pixel: computing on the
        # padding between tiles is overhead, not work) -> 3701.48 MAC/pixel for the MMOD face net
        P = float(geo.total_level_pixels())
        cum = 1
        self.algorithmic_flops_per_layer = []
        for (cout, cin, k, s) in W.DET_CONVS:
            cum *= s
            self.algorithmic_flops_per_layer.append(2.0 * P / (cum * cum) * cout * cin * k * k)
        self.algorithmic_flops_per_frame = sum(self.algorithmic_flops_per_layer)
        # the same, per launch of self.convs (conv1 + conv2 share one launch in "c12" mode)
        names = getattr(self, "layer_names", None) or ["conv%d" % (i + 1) for i in range(len(self.convs))]
        self.layer_names = names
        af = list(self.algorithmic_flops_per_layer)
        self.algorithmic_flops_per_op = ([af[0] + af[1]] + af[2:]) if names[0] == "conv1+2" else af

    def _init_detconv(self, model, device):
        """layers 2..7 on csrc/rsconv.cu (row streaming) or csrc/detconv.cu (2-D tiles): plain NHWC activations
        [B, H, even(W), C], no padding in HBM"""
        Conv = RsConv if self.conv_impl == "rsconv" else DetConv
        B, geo = self.B, self.geo
        Hp, Wp = geo.plane_h, geo.plane_w
        convs = model["convs"]
        self.convs = []
        self.flops_per_frame = 0
        self.flops_per_layer = []
        self.layer_names = []
        first = 1
        if self.conv_impl == "rsconv" and self.conv1_mode == "c12":
            # conv1 + conv2 in ONE strip kernel (csrc/c12.cu): the conv1 activations never reach HBM
            self.lg, self.xg = None, None
            (sc1, sh1), (sc2, sh2) = _affine(convs[0]), _affine(convs[1])
            op = FusedC12(self.plane, Hp, Wp, _t(convs[0]["w"]), sc1, sh1, _t(convs[1]["w"]), sc2, sh2, W.PIXEL_MEAN)
            self._c12_args = (Hp, Wp, _t(convs[0]["w"]), sc1, sh1, _t(convs[1]["w"]), sc2, sh2, W.PIXEL_MEAN)
            self.convs.append((op, 1))
            self.layer_names.append("conv1+2")
            for i in range(2):
                cout, cin, k, s = W.DET_CONVS[i]
                oh, ow = (op.OH1, op.OW1) if i == 0 else (op.OH, op.OW)
                self.flops_per_layer.append(2 * oh * ow * cout * cin * k * k)
            x, h, w = op.out, op.OH, op.OW
            first = 2
        if first == 1:
            x, h, w = self._init_conv1(model, device, B, Hp, Wp)
        n = len(convs)
        for i in range(first, n):
            cout, cin, k, s = W.DET_CONVS[i]
            c = convs[i]
            last = i == n - 1
            c_in = x.shape[3]
            if last:
                # 9x9, one output channel: the 9 filter columns become 9 output channels of a 9x1 conv;
                # pv_det_shift_sum_nhwc adds the kw-shifted channels back together.
                wt = _t(c["w"]).float()                                   # [1, cin, 9, 9]
                w9 = wt[0].permute(2, 0, 1).unsqueeze(-1).contiguous()    # [kw, cin, kh, 1]
                op = Conv(x, h, w, w9, 1, torch.ones(k), torch.zeros(k), False, c_in, 16, out_f32=True)
                self.partial = op.out
                self.OH, self.OW = op.OH, w
                self.scores = torch.zeros(B, self.OH, self.OW, dtype=torch.float32, device=device)
                self.score_bias = float(_affine(c, affine=False)[1][0])
                self.flops_per_layer.append(2 * self.OH * self.OW * cout * cin * k * k)
            else:
                sc, sh = _affine(c)
                n_out = (cout + 15) // 16 * 16
                op = Conv(x, h, w, _t(c["w"]), s, sc, sh, True, c_in, n_out)
                self.flops_per_layer.append(2 * op.OH * op.OW * cout * cin * k * k)
                x, h, w = op.out, op.OH, op.OW
            self.convs.append((op, 1))
            self.layer_names.append("conv%d" % (i + 1))
        self.flops_per_frame = sum(self.flops_per_layer)

    def _init_conv1(self, model, device, B, Hp, Wp):
        """conv1 (5x5 s2, RGB -> 16) as its own launch: writes straight into the NHWC tensor through a row map"""
        convs = model["convs"]
        cout, cin, k, s = W.DET_CONVS[0]
        OH1, OW1 = (Hp - k) // s + 1, (Wp - k) // s + 1
        l1 = RowLayout("padded", B, OH1, even(OW1), 16, pad=0)
        a1 = l1.alloc(device)
        if self.conv1_mode == "fused":
            self.lg, self.xg = None, None
            op = FusedConv1(self.plane, Hp, Wp, convs[0], a1, l1, device)
            img1 = op.img
        else:
            lg = RowLayout(self.conv1_mode, B, Hp, Wp, 3, kw=5)
            self.lg, self.xg = lg, lg.alloc(device)
            cp = ConvPlan(lg, _t(convs[0]["w"]), s, 0, group=config.SRGEMM_GROUP)
            sc, sh = _affine(convs[0])
            op = Srgemm(cp, self.xg, a1, l1, sc, sh, relu=True)
            img1 = lg.img
        self.convs.append((op, img1))
        self.flops_per_layer.append(2 * OH1 * OW1 * cout * cin * k * k)
        self.layer_names.append("conv1")
        return a1.view(B, OH1, even(OW1), 16), OH1, OW1

    def _init_srgemm(self, model, device, group):
        B, geo = self.B, self.geo
        Hp, Wp = geo.plane_h, geo.plane_w
        lg = RowLayout("gathered" if self.conv1_mode == "fused" else self.conv1_mode, B, Hp, Wp, 3, kw=5)
        self.lg = lg
        self.xg = lg.alloc(device) if self.conv1_mode != "fused" else None
        self.convs = []
        self.flops_per_frame = 0
        self.flops_per_layer = []
        convs = model["convs"]
        lcur, cur = lg, self.xg
        n = len(convs)
        for i, c in enumerate(convs):
            cout, cin, k, s = W.DET_CONVS[i]
            pad = W.conv_pad(k, s)
            last = i == n - 1
            sc, sh = _affine(c, affine=not last)
            if last:
                # 9x9, one output channel: the 9 filter columns become 9 output channels of a 9x1
                # conv (K = 9 taps x 48 instead of 81 taps x 48); pv_det_shift_sum adds the
                # kw-shifted channels back together.
                assert cout == 1 and lcur.kind == "padded" and lcur.pad == pad
                wt = _t(c["w"]).float()                       # [1, cin, k, k]
                taps = []
                for kh in range(k):
                    m = torch.zeros(16, lcur.cols)
                    m[:k, :cin] = wt[0, :, kh, :].t()         # row kw, col c
                    taps.append((kh * lcur.Wq, m))
                OHs, OWs = lcur.H + 2 * pad - k + 1, lcur.W + 2 * pad - k + 1
                cp = ConvPlan.from_taps(lcur, taps, k, OHs, lcur.Wq, group="tap", kernel=(k, 1))
                self.OH, self.OW = OHs, OWs
                self.lpart = RowLayout("padded", B, lcur.Hq, lcur.Wq, 16, pad=0)
                self.partial = torch.zeros(self.lpart.rows, 16, dtype=torch.float32, device=device)
                self.scores = torch.zeros(B, OHs, OWs, dtype=torch.float32, device=device)
                self.score_bias = float(sh[0])
                op = Srgemm(cp, cur, self.partial, self.lpart, torch.ones(k), torch.zeros(k), relu=False, out_rows_f32=True)
            elif i == 0 and self.conv1_mode == "fused":
                OH1, OW1 = (Hp - k) // s + 1, (Wp - k) // s + 1
                lo = RowLayout("parity", B, OH1, OW1, 16, pad=0)
                o = lo.alloc(device)
                op = FusedConv1(self.plane, Hp, Wp, c, o, lo, device)

                class _CP(object):
                    pass
                cp = _CP()
                cp.OH, cp.OW, cp.lin = OH1, OW1, lg
            else:
                cp = ConvPlan(lcur, _t(c["w"]), s, pad, group=group)
                _, _, nk, ns = W.DET_CONVS[i + 1]
                if ns == 2:
                    lo = RowLayout("parity", B, cp.OH, cp.OW, cp.N, pad=0)
                else:
                    lo = RowLayout("padded", B, cp.OH, cp.OW, cp.N, pad=W.conv_pad(nk, ns))
                o = lo.alloc(device)
                op = Srgemm(cp, cur, o, lo, sc, sh, relu=True)
            self.convs.append((op, cp.lin.img))
            self.flops_per_layer.append(2 * (self.OH * self.OW if last else cp.OH * cp.OW) * cout * cin * k * k)
            self.flops_per_frame += self.flops_per_layer[-1]
            if not last:
                lcur, cur = lo, o

    def _init_pyramid(self):
        """tables of build_plane (shared with hog.HogDetectorNet, which scans the same tiled pyramid plane)"""
        geo = self.geo
        # pyramid: big levels one launch each, the small tail (<= TAIL_PIXELS per level) in one launch
        f32 = np.float32
        tail_from = next((i for i, (w, h) in enumerate(geo.sizes) if i >= 1 and w * h <= self.TAIL_PIXELS), geo.n_levels)
        self._tail_from = tail_from
        self._tail_n = geo.n_levels - tail_from
        if self._tail_n > 0:
            self._tail_rects = np.asarray(geo.rects[tail_from - 1:], np.int32).reshape(-1, 4).copy()
            sc = []
            for lv in range(tail_from, geo.n_levels):
                (pw, ph), (w, h) = geo.sizes[lv - 1], geo.sizes[lv]
                sc.append((f32(pw - 1) / f32(max(w - 1, 1)), f32(ph - 1) / f32(max(h - 1, 1))))
            self._tail_scales = np.asarray(sc, np.float32).reshape(-1, 2).copy()

    def _init_tail(self, model, device):
        B, geo = self.B, self.geo
        self._init_pyramid()
        rects, fxy = geo.level_table()
        self.level_rects = _t(rects).to(device)
        self.level_fxy = _t(fxy).to(device)
        self.counts = torch.zeros(B, dtype=torch.int32, device=device)
        self.cand_score = torch.zeros(B, self.MAX_CAND, dtype=torch.float32, device=device)
        self.cand_cell = torch.zeros(B, self.MAX_CAND, dtype=torch.int32, device=device)
        self.out_boxes = torch.zeros(B, self.MAX_DET, 4, dtype=torch.int32, device=device)
        self.out_scores = torch.zeros(B, self.MAX_DET, dtype=torch.float32, device=device)
        self.out_counts = torch.zeros(B, dtype=torch.int32, device=device)
        cm, ca = det_cell_to_plane(1, 1)[0] - det_cell_to_plane(0, 0)[0], det_cell_to_plane(0, 0)[0]
        self.cell_mul, self.cell_add = cm, ca

    def enable_double_buffer(self):
        """allocate a second pyramid plane: build_plane of batch s+1 (issue-bound CUDA-core work) can then run on
        another stream while the tensor-core convs of batch s read the first plane"""
        if len(self.planes) == 1:
            self.planes.append(torch.zeros_like(self.planes[0]))

    def use_plane(self, slot):
        """select the plane that the next build_plane / forward_scores calls bind (host-side pointer switch)"""
        self.plane = self.planes[slot]
        op = self.convs[0][0]
        if isinstance(op, FusedConv1):
            op.plane = self.plane
        elif isinstance(op, FusedC12):
            # the plane pointer lives in the kernel's tensor map: one bound op per plane, same output tensor
            if not hasattr(self, "_c12_ops"):
                self._c12_ops = {0: op}
            if slot not in self._c12_ops:
                Hp, Wp, w1, sc1, sh1, w2, sc2, sh2, mean = self._c12_args
                self._c12_ops[slot] = FusedC12(self.plane, Hp, Wp, w1, sc1, sh1, w2, sc2, sh2, mean, out=op.out)
            self.convs[0] = (self._c12_ops[slot], 1)

    def build_plane(self, frames, M):
        """frames: uint8 [M,H,W,3] device tensor -> tiled pyramid plane (RGBA u8)."""
        L = _lib.lib()
        st = _lib.stream_ptr()
        geo = self.geo
        Hp, Wp = geo.plane_h, geo.plane_w
        f32 = np.float32
        prev = None
        tail_from = self._tail_from
        for lv, (x0, y0, w, h) in enumerate(geo.rects[:tail_from]):
            if lv == 0:
                sw, sh = self.W, self.H
                xs = float(f32(sw - 1) / f32(max(w - 1, 1)))
                ys = float(f32(sh - 1) / f32(max(h - 1, 1)))
                _lib.check(L.pv_resize_bilinear(_lib.ptr(frames), 3, C.c_int64(self.H * self.W * 3), self.W, 0, 0, sw, sh,
                                                _lib.ptr(self.plane), C.c_int64(Hp * Wp), Wp, x0, y0, w, h,
                                                C.c_float(xs), C.c_float(ys), M, 0 if self.upsample else 1, st),
                           "pv_resize_bilinear")
            else:
                px0, py0, pw, ph = prev
                xs = float(f32(pw - 1) / f32(max(w - 1, 1)))
                ys = float(f32(ph - 1) / f32(max(h - 1, 1)))
                _lib.check(L.pv_resize_bilinear(_lib.ptr(self.plane), 4, C.c_int64(Hp * Wp * 4), Wp, px0, py0, pw, ph,
                                                _lib.ptr(self.plane), C.c_int64(Hp * Wp), Wp, x0, y0, w, h,
                                                C.c_float(xs), C.c_float(ys), M, 0, st), "pv_resize_bilinear")
            prev = (x0, y0, w, h)
        if self._tail_n > 0:
            _lib.check(L.pv_pyramid_tail(_lib.ptr(self.plane), C.c_int64(Hp * Wp), Wp, M, self._tail_n,
                                         self._tail_rects.ctypes.data_as(C.POINTER(C.c_int)),
                                         self._tail_scales.ctypes.data_as(C.POINTER(C.c_float)), st), "pv_pyramid_tail")

    def forward_scores(self, M):
        L = _lib.lib()
        st = _lib.stream_ptr()
        geo = self.geo
        if self.conv1_mode in ("fused", "c12"):
            pass   # conv1 reads the plane in place
        elif self.conv1_mode == "pixrows":
            # NB: the pixel buffer is [2, Bcap, Hq, W]; a partial batch writes the first M images of
            # each parity plane, which is where rows n < M of the layout live.
            _lib.check(L.pv_plane_to_pixrows(_lib.ptr(self.plane), _lib.ptr(self.xg), M, geo.plane_h, geo.plane_w,
                                             C.c_int64(self.lg.plane_rows), _mean3(), st), "pv_plane_to_pixrows")
        else:
            _lib.check(L.pv_pack_gathered(_lib.ptr(self.plane), _lib.ptr(self.xg), M, geo.plane_h, geo.plane_w, 5,
                                          C.c_int64(self.lg.plane_rows), _mean3(), st), "pv_pack_gathered")
        for op, img in self.convs:
            op.run(M * img)
        if self.conv_impl in ("detconv", "rsconv"):
            _lib.check(L.pv_det_shift_sum_nhwc(_lib.ptr(self.partial), M, self.OH, self.OW, self.partial.shape[2], 16, 9,
                                               C.c_float(self.score_bias), _lib.ptr(self.scores), st),
                       "pv_det_shift_sum_nhwc")
            return self.scores[:M]
        lp = self.lpart
        _lib.check(L.pv_det_shift_sum(_lib.ptr(self.partial), M, lp.Hq, lp.Wq, 16, self.OH, self.OW, 9,
                                      C.c_float(self.score_bias), _lib.ptr(self.scores), st), "pv_det_shift_sum")
        return self.scores[:M]

    def decode(self, M, threshold=None):
        L = _lib.lib()
        st = _lib.stream_ptr()
        m = self.model
        thr = float(m["adjust_threshold"]) if threshold is None else max(float(threshold), float(m["adjust_threshold"]))
        _lib.check(L.pv_det_candidates(_lib.ptr(self.scores), M, self.OH * self.OW, C.c_float(thr),
                                       _lib.ptr(self.counts), _lib.ptr(self.cand_score), _lib.ptr(self.cand_cell),
                                       self.MAX_CAND, st), "pv_det_candidates")
        _lib.check(L.pv_det_nms(_lib.ptr(self.counts), _lib.ptr(self.cand_score), _lib.ptr(self.cand_cell), self.MAX_CAND,
                                M, _lib.ptr(self.level_rects), _lib.ptr(self.level_fxy), self.geo.n_levels,
                                int(m["window"]), self.OW, self.cell_mul, self.cell_add,
                                C.c_double(float(m["iou_thresh"])), C.c_double(float(m["covered_thresh"])), self.MAX_DET,
                                _lib.ptr(self.out_boxes), _lib.ptr(self.out_scores), _lib.ptr(self.out_counts), st),
                   "pv_det_nms")
        return self.out_boxes[:M], self.out_scores[:M], self.out_counts[:M]

    def detect(self, frames):
        """frames uint8 [M,H,W,3] (device).  Returns (boxes int32 [M,MAX_DET,4], scores, counts) on device."""
        M = frames.shape[0]
        assert M <= self.B and frames.shape[1:] == (self.H, self.W, 3) and frames.dtype == torch.uint8
        self.build_plane(frames.contiguous(), M)
        self.forward_scores(M)
        return self.decode(M)

    def check(self):
        for op, _ in self.convs:
            op.check()
