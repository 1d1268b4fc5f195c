"""GPU: the tcgen05 srgemm kernel against the CPU emulator of its contract (bit-level up to
fp32 accumulation order; tolerance 1 bf16 ulp of the result + 0.02 absolute)."""
import pytest
import torch

from pyannote_video_b200.plan import ConvPlan, RowLayout, Srgemm
from srgemm_emu import emulate

pytestmark = pytest.mark.gpu

CASES = [
    ("padded", 2, 35, 35, 32, 32, 3, 1, 1),
    ("padded", 2, 17, 17, 64, 64, 3, 1, 1),
    ("padded", 3, 8, 8, 128, 128, 3, 1, 1),
    ("padded", 5, 4, 4, 256, 256, 3, 1, 1),
    ("padded", 1, 40, 60, 48, 45, 5, 1, 2),
    ("parity", 1, 61, 83, 16, 32, 5, 2, 0),
    ("parity", 2, 35, 35, 32, 64, 3, 2, 0),
    ("gathered", 1, 63, 90, 3, 16, 5, 2, 0),
    ("gathered", 2, 150, 150, 3, 32, 7, 2, 0),
]


@pytest.mark.parametrize("kind,B,H,W,cin,cout,k,stride,pad", CASES)
def test_srgemm_matches_emulator(cuda, kind, B, H, W, cin, cout, k, stride, pad):
    from pyannote_video_b200 import config
    torch.manual_seed(5)
    x = torch.randn(B, H, W, cin).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).to(torch.bfloat16).float()
    lin = RowLayout("gathered", B, H, W, 3, kw=k) if kind == "gathered" else RowLayout(kind, B, H, W, cin, pad=pad)
    cp = ConvPlan(lin, w, stride, pad, group=config.SRGEMM_GROUP)
    lout = RowLayout("padded", B, cp.OH, cp.OW, cp.N, pad=1)
    scale, shift = torch.rand(cout) + 0.5, torch.randn(cout) * 0.1
    xr = lin.to_rows(x)
    resid = torch.randn(lout.rows, lout.cols).to(torch.bfloat16)
    out = lout.alloc(cuda)
    op = Srgemm(cp, xr.to(cuda), out, lout, scale, shift, relu=True, resid=resid.to(cuda), lres=lout)
    op.run()
    op.check()
    ref = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    emulate(cp, xr, ref, lout, scale, shift, relu=True, resid=resid, lres=lout)
    d = (out.cpu().float() - ref.float()).abs()
    assert (d <= 0.02 + 0.01 * ref.float().abs()).all(), float(d.max())
