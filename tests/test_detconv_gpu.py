"""GPU: csrc/detconv.cu (2-D tiled implicit-GEMM convs of the detector) against a CPU fp32 convolution of the
same bf16-rounded operands (torch conv2d = the oracle's `con` layer, oracle/nets.py).  fp32 accumulation order
differs, so the tolerance is 1 bf16 ulp of the result + 0.02 absolute; fp32 outputs 2e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

from pyannote_video_b200.detconv import DetConv, RsConv, even

pytestmark = pytest.mark.gpu

CASES = [
    # c_in, cin, n_out, cout, kh, kw, stride, f32, B, H, W
    (16, 16, 32, 32, 5, 5, 2, False, 2, 37, 53),
    (32, 32, 32, 32, 5, 5, 2, False, 1, 69, 40),
    (32, 32, 48, 45, 5, 5, 1, False, 2, 19, 27),
    (48, 45, 48, 45, 5, 5, 1, False, 2, 33, 21),
    (48, 45, 16, 9, 9, 1, 1, True, 1, 35, 18),
    (48, 45, 48, 45, 5, 5, 1, False, 1, 5, 6),          # smaller than one tile
    (16, 16, 32, 32, 5, 5, 2, False, 3, 5, 5),          # a single output pixel per image
]


CASES_RS = CASES + [
    (48, 45, 48, 45, 5, 5, 1, False, 1, 70, 150),      # two strips, several row segments, slot-ring wraps
    (16, 16, 32, 32, 5, 5, 2, False, 2, 131, 300),
    (48, 45, 16, 9, 9, 1, 1, True, 2, 90, 140),
]


@pytest.mark.parametrize("impl", ["rsconv", "detconv"])
@pytest.mark.parametrize("c_in,cin,n_out,cout,kh,kw,stride,f32,B,H,W", CASES_RS)
def test_detconv_matches_conv2d(cuda, impl, c_in, cin, n_out, cout, kh, kw, stride, f32, B, H, W):
    Conv = RsConv if impl == "rsconv" else DetConv
    torch.manual_seed(3)
    x = torch.randn(B, H, W, cin).to(torch.bfloat16)
    w = (torch.randn(cout, cin, kh, kw) / (cin * kh * kw) ** 0.5).to(torch.bfloat16).float()
    scale, shift = torch.rand(cout) + 0.5, torch.randn(cout) * 0.1
    xd = torch.zeros(B, H, even(W), c_in, dtype=torch.bfloat16, device=cuda)
    xd[:, :, :W, :cin] = x.to(cuda)
    op = Conv(xd, H, W, w, stride, scale, shift, not f32, c_in, n_out, out_f32=f32)
    op.out.fill_(7.0)     # positions outside the valid extent must stay untouched
    op.run()
    op.check()
    pad = (kh // 2, kw // 2) if stride == 1 else (0, 0)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=pad)
    ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if not f32:
        ref = ref.clamp_min(0)
    ref = ref.permute(0, 2, 3, 1)
    assert (op.OH, op.OW) == tuple(ref.shape[1:3])
    got = op.out.cpu().float()
    d = (got[:, :, :op.OW, :cout] - ref).abs()
    tol = (2e-3 * ref.abs() + 1e-3) if f32 else (0.02 + 0.01 * ref.abs())
    assert (d <= tol).all(), float(d.max())
    assert (got[:, :, op.OW:] == 7.0).all()
    if n_out > cout and not f32:
        assert (got[:, :, :op.OW, cout:n_out] == 0).all()     # padded output channels are exact zeros
    # partial batch: only the first image is recomputed
    if B > 1:
        op.out.fill_(7.0)
        op.run(1)
        op.check()
        got1 = op.out.cpu().float()
        assert torch.equal(got1[0, :, :op.OW], got[0, :, :op.OW]) and (got1[1:] == 7.0).all()
