"""GPU: csrc/c12.cu (detector conv1 + conv2 in one strip kernel) against a CPU fp32 restatement of the two `con` layers on
the same bf16-rounded operands (torch conv2d = the oracle's layer, oracle/nets.py), and against the two-launch CUDA path
(conv1_fused + rsconv).  The conv1 activations are rounded to bf16 in both; fp32 accumulation order differs, so the
tolerance is ~1 bf16 ulp of the result + 0.02 absolute."""
import pytest
import torch
import torch.nn.functional as F

from pyannote_video_b200.detconv import FusedC12
from pyannote_video_b200 import weights as Wt

pytestmark = pytest.mark.gpu

CASES = [
    # B, Hp, Wp
    (1, 29, 40),            # a single conv2 row of 7 columns: one item, one strip
    (2, 131, 300),          # partial strip, several quads
    (1, 93, 1100),          # three strips (2 x 124 + 22 conv2 columns)
    (3, 277, 520),          # row segments (ring wraps across items), two strips
    (1, 13, 16),            # the smallest plane with one conv2 output
]


def _reference(plane, w1, sc1, sh1, w2, sc2, sh2):
    mean = torch.tensor(Wt.PIXEL_MEAN, dtype=torch.float32)
    rgb = plane[..., :3].float()
    x = ((rgb - mean) / 256.0)
    x = torch.where(plane[..., 3:4] != 0, x, torch.zeros_like(x)).to(torch.bfloat16).float()
    a1 = F.conv2d(x.permute(0, 3, 1, 2), w1, stride=2)
    a1 = (a1 * sc1.view(1, -1, 1, 1) + sh1.view(1, -1, 1, 1)).clamp_min(0).to(torch.bfloat16).float()
    a2 = F.conv2d(a1, w2, stride=2)
    a2 = (a2 * sc2.view(1, -1, 1, 1) + sh2.view(1, -1, 1, 1)).clamp_min(0)
    return a1.permute(0, 2, 3, 1), a2.permute(0, 2, 3, 1)


@pytest.mark.parametrize("B,Hp,Wp", CASES)
def test_c12_matches_two_conv2d(cuda, B, Hp, Wp):
    torch.manual_seed(5)
    plane = torch.randint(0, 256, (B, Hp, Wp, 4), dtype=torch.uint8)
    plane[..., 3] = 255
    plane[:, Hp // 3:Hp // 3 + 7, Wp // 4:Wp // 4 + 9, :] = 0            # a patch of pyramid padding (alpha 0 -> exact zeros)
    w1 = (torch.randn(16, 3, 5, 5) / 75 ** 0.5).to(torch.bfloat16).float()
    w2 = (torch.randn(32, 16, 5, 5) / 400 ** 0.5).to(torch.bfloat16).float()
    sc1, sh1 = torch.rand(16) + 0.5, torch.randn(16) * 0.1
    sc2, sh2 = torch.rand(32) + 0.5, torch.randn(32) * 0.1
    pd = plane.to(cuda)
    op = FusedC12(pd, Hp, Wp, w1, sc1, sh1, w2, sc2, sh2, Wt.PIXEL_MEAN)
    op.out.fill_(7.0)
    op.run()
    op.check()
    _, ref = _reference(plane, w1, sc1, sh1, w2, sc2, sh2)
    assert (op.OH, op.OW) == tuple(ref.shape[1:3])
    got = op.out.cpu().float()
    d = (got[:, :, :op.OW] - ref).abs()
    tol = 0.02 + 0.01 * ref.abs()
    assert (d <= tol).all(), (float(d.max()), op.info())
    assert (got[:, :, op.OW:] == 7.0).all()               # the pitch column is never written
    if B > 1:                                              # partial batch
        op.out.fill_(7.0)
        op.run(1)
        op.check()
        got1 = op.out.cpu().float()
        assert torch.equal(got1[0, :, :op.OW], got[0, :, :op.OW]) and (got1[1:] == 7.0).all()


def test_c12_equals_the_two_launch_path_on_a_detector_plane(cuda):
    """same detector, conv1_mode "c12" vs "fused": conv2 activations agree to bf16 rounding, score maps to 1e-2"""
    from pyannote_video_b200.nets import DetectorNet
    from pyannote_video_b200.synth import make_frames
    model = Wt.make_detector(seed=2)
    frames = make_frames(2, 270, 480, seed=3, device=cuda)
    nets = {m: DetectorNet(model, 270, 480, 1, 2, cuda, conv1_mode=m, conv_impl="rsconv") for m in ("fused", "c12")}
    outs = {}
    for m, net in nets.items():
        net.build_plane(frames, 2)
        s = net.forward_scores(2).clone()
        for op, _ in net.convs:
            op.check()
        outs[m] = s
    assert torch.equal(nets["fused"].plane, nets["c12"].plane)
    a = nets["fused"].convs[1][0].out.float()
    b = nets["c12"].convs[0][0].out.float()
    assert a.shape == b.shape
    assert ((a - b).abs() <= 0.02 + 0.01 * a.abs()).all(), float((a - b).abs().max())
    rng = float(outs["fused"].max() - outs["fused"].min())
    assert float((outs["fused"] - outs["c12"]).abs().max()) <= 0.01 * max(rng, 1.0)
