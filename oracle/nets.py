"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the two dlib networks the reference reaches through
  * pyannote/video/face/face.py:66   face_detector_(rgb, 1)        (as dlib's CNN/MMOD net)
  * pyannote/video/face/face.py:74   compute_face_descriptor(...)  (face_recognition_resnet_model_v1)
following SURVEY.md App. A.1 / A.4 (dlib 19.12 `dnn` layer semantics, recalled — parity unpinned:
dlib, its weights and its tests are absent from the build environment).

torch CPU `conv2d` is used for the contractions (a floating-point kernel; the tier rules allow a
torch fp32 reference there).  `bf16=True` rounds weights and every stored activation to bfloat16
exactly where the CUDA path stores bf16, so kernel parity can be asserted tightly; `bf16=False` is
the plain fp32 network used to state the embedding tolerance.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import constants as W


def _r(x, bf16):
    return x.to(torch.bfloat16).float() if bf16 else x


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).float()


def _conv_affine(x, c, stride, bf16, relu=True, affine=True):
    """dlib: relu<affine<con<...>>>; con has a bias, affine is y = gamma*x + beta."""
    k = c["w"].shape[-1]
    y = F.conv2d(x, _r(_t(c["w"]), bf16), None, stride=stride, padding=W.conv_pad(k, stride))
    g, b, be = _t(c["gamma"]), _t(c["b"]), _t(c["beta"])
    if affine:
        scale, shift = g, g * b + be
    else:
        scale, shift = torch.ones_like(b), b
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    return torch.relu(y) if relu else y


def detector_forward(model, plane, bf16=False):
    """plane: float [B,3,H,W], already (v - mean)/256.  Returns score map [B,OH,OW] fp32."""
    x = _r(plane, bf16)
    n = len(model["convs"])
    for i, c in enumerate(model["convs"]):
        stride = W.DET_CONVS[i][3]
        last = i == n - 1
        x = _conv_affine(x, c, stride, bf16, relu=not last, affine=not last)
        if not last:
            x = _r(x, bf16)
    return x[:, 0]


def detector_out_size(n):
    for (_, _, k, s) in W.DET_CONVS:
        n = (n + 2 * W.conv_pad(k, s) - k) // s + 1
    return n


def _zero_extend_add(a, b):
    """dlib add_prev: result has the max extent in every dim, missing entries are zero."""
    C = max(a.shape[1], b.shape[1])
    H = max(a.shape[2], b.shape[2])
    Wd = max(a.shape[3], b.shape[3])
    out = torch.zeros(a.shape[0], C, H, Wd)
    out[:, :a.shape[1], :a.shape[2], :a.shape[3]] += a
    out[:, :b.shape[1], :b.shape[2], :b.shape[3]] += b
    return out


def embed_forward(model, chips, bf16=False, return_taps=False):
    """chips: float [B,3,150,150], already (v - mean)/256.  Returns [B,128] fp32 (no L2 norm)."""
    taps = {}
    x = _r(chips, bf16)
    x = _r(_conv_affine(x, model["conv1"], 2, bf16), bf16)        # 150 -> 72
    taps["conv1"] = x
    x = F.max_pool2d(x, 3, stride=2, padding=0)                    # 72 -> 35
    taps["pool1"] = x
    for i, blk in enumerate(model["blocks"]):
        if blk["type"] == "ares":
            t = _r(_conv_affine(x, blk["a"], 1, bf16), bf16)
            u = _conv_affine(t, blk["b"], 1, bf16, relu=False)
            x = _r(torch.relu(x + u), bf16)
        else:
            t = _r(_conv_affine(x, blk["a"], 2, bf16), bf16)     # 3x3 s2 p0
            u = _conv_affine(t, blk["b"], 1, bf16, relu=False)
            s = _r(F.avg_pool2d(x, 2, stride=2, padding=0), bf16)  # skip path (stored bf16 on GPU)
            x = _r(torch.relu(_zero_extend_add(s, u)), bf16)
        taps["block%d" % i] = x
    g = x.mean(dim=(2, 3))                                         # avg_pool_everything
    out = g @ _t(model["fc"]).t()                                  # fc_no_bias<128>
    return (out, taps) if return_taps else out


def normalize_rgb(img_u8_nhwc):
    """input_rgb_image(_pyramid/_sized): (v - mean) / 256 per channel; returns NCHW float."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8_nhwc)).float()
    mean = torch.tensor(W.PIXEL_MEAN)
    x = (x - mean) * W.PIXEL_SCALE
    return x.permute(0, 3, 1, 2).contiguous()
