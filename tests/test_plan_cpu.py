"""CPU: conv -> srgemm tap tables reproduce torch conv2d through the kernel emulator."""
import pytest
import torch
import torch.nn.functional as F

from pyannote_video_b200.plan import ConvPlan, RowLayout
from srgemm_emu import emulate


def _bf(x):
    return x.to(torch.bfloat16).float()


@pytest.mark.parametrize("group", ["tap", "row", "all"])
@pytest.mark.parametrize("cin,cout,k,pad", [(32, 32, 3, 1), (48, 45, 5, 2), (16, 32, 3, 1), (64, 64, 3, 1), (48, 1, 9, 4)])
def test_stride1_padded(group, cin, cout, k, pad):
    torch.manual_seed(0)
    B, H, W = 2, 11, 13
    x = _bf(torch.randn(B, H, W, cin))
    w = _bf(torch.randn(cout, cin, k, k) * 0.1)
    lin = RowLayout("padded", B, H, W, cin, pad=pad)
    cp = ConvPlan(lin, w, 1, pad, group=group)
    lout = RowLayout("padded", B, cp.OH, cp.OW, cp.N, pad=1)
    out = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    scale = torch.rand(cout) + 0.5
    shift = torch.randn(cout)
    emulate(cp, lin.to_rows(x), out, lout, scale, shift, relu=True)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=pad) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    got = lout.from_rows(out, C=cout)
    assert torch.allclose(got, ref, atol=2e-2, rtol=2e-2)
    # border stays zero
    full = out.float().reshape(B, lout.Hq, lout.Wq, -1)
    assert full[:, 0].abs().max() == 0 and full[:, :, 0].abs().max() == 0


@pytest.mark.parametrize("group", ["tap", "row"])
@pytest.mark.parametrize("cin,cout,k,H,W", [(16, 32, 5, 21, 18), (32, 64, 3, 17, 17), (128, 256, 3, 8, 8), (256, 256, 3, 4, 4)])
def test_stride2_parity(group, cin, cout, k, H, W):
    torch.manual_seed(1)
    B = 2
    x = _bf(torch.randn(B, H, W, cin))
    w = _bf(torch.randn(cout, cin, k, k) * 0.1)
    lin = RowLayout("parity", B, H, W, cin, pad=0)
    cp = ConvPlan(lin, w, 2, 0, group=group)
    lout = RowLayout("parity", B, cp.OH, cp.OW, cp.N, pad=0)
    out = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    scale = torch.ones(cout)
    shift = torch.zeros(cout)
    emulate(cp, lin.to_rows(x), out, lout, scale, shift, relu=False)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2).permute(0, 2, 3, 1)
    got = lout.from_rows(out, C=cout)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=3e-2, rtol=3e-2)


@pytest.mark.parametrize("k,cout,H,W", [(5, 16, 23, 30), (7, 32, 150, 150)])
def test_stride2_gathered_first_layer(k, cout, H, W):
    torch.manual_seed(2)
    B = 2
    x = _bf(torch.randn(B, H, W, 3))
    w = _bf(torch.randn(cout, 3, k, k) * 0.1)
    lin = RowLayout("gathered", B, H, W, 3, kw=k)
    cp = ConvPlan(lin, w, 2, 0, group="tap")
    lout = RowLayout("padded", B, cp.OH, cp.OW, cp.N, pad=0)
    out = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    emulate(cp, lin.to_rows(x), out, lout, torch.ones(cout), torch.zeros(cout), relu=False)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2).permute(0, 2, 3, 1)
    got = lout.from_rows(out, C=cout)
    assert torch.allclose(got, ref, atol=3e-2, rtol=3e-2)


def test_residual_and_f32_out():
    torch.manual_seed(3)
    B, H, W, C = 1, 6, 7, 32
    x = _bf(torch.randn(B, H, W, C))
    w = _bf(torch.randn(C, C, 3, 3) * 0.1)
    lin = RowLayout("padded", B, H, W, C, pad=1)
    cp = ConvPlan(lin, w, 1, 1, group="row")
    lout = RowLayout("parity", B, H, W, C, pad=0)
    out = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    xr = lin.to_rows(x)
    emulate(cp, xr, out, lout, torch.ones(C), torch.zeros(C), relu=True, resid=xr, lres=lin)
    ref = torch.relu(F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1) + x)
    assert torch.allclose(lout.from_rows(out), ref, atol=3e-2, rtol=3e-2)
    # f32 single-channel output
    w1 = _bf(torch.randn(1, C, 3, 3) * 0.1)
    cp1 = ConvPlan(lin, w1, 1, 1, group="tap")
    o = torch.zeros(B, cp1.OH, cp1.OW)
    emulate(cp1, xr, o, None, torch.ones(1), torch.zeros(1), relu=False, out_f32=True)
    ref1 = F.conv2d(x.permute(0, 3, 1, 2), w1, padding=1)[:, 0]
    assert torch.allclose(o, ref1, atol=1e-3, rtol=1e-3)


def test_single_channel_9x9_as_columns_in_n():
    """detector last layer: 9x9 conv with one output channel computed as a 9x1 conv whose 9 output
    channels are the filter columns, then re-assembled by a kw-shifted sum."""
    torch.manual_seed(7)
    B, H, W, C, k, pad = 2, 12, 15, 48, 9, 4
    x = _bf(torch.randn(B, H, W, 45))
    w = _bf(torch.randn(1, 45, k, k) * 0.05)
    lin = RowLayout("padded", B, H, W, C, pad=pad)
    taps = []
    for kh in range(k):
        m = torch.zeros(16, lin.cols)
        m[:k, :45] = w[0, :, kh, :].t()
        taps.append((kh * lin.Wq, m))
    OH, OW = H, W
    cp = ConvPlan.from_taps(lin, taps, k, OH, lin.Wq, group="tap", kernel=(k, 1))
    assert cp.mma_per_tile == 9 * 3 and cp.resident
    lpart = RowLayout("padded", B, lin.Hq, lin.Wq, 16, pad=0)
    D = torch.zeros(lpart.rows, 16)
    emulate(cp, lin.to_rows(x), D, lpart, torch.ones(k), torch.zeros(k), relu=False, out_rows_f32=True)
    D = D.reshape(B, lin.Hq, lin.Wq, 16)
    score = sum(D[:, :OH, kw:kw + OW, kw] for kw in range(k))
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=pad)[:, 0]
    assert torch.allclose(score, ref, atol=2e-3, rtol=2e-3)


def test_slots_cover_entries():
    w = _bf(torch.randn(45, 48, 5, 5) * 0.1)
    lin = RowLayout("padded", 1, 20, 30, 48, pad=2)
    cp = ConvPlan(lin, w, 1, 2, group="row")
    flags = [e.flags for e in cp.entries]
    assert flags[0] & 1 and flags[-1] & 2
    for a, b in zip(flags[:-1], flags[1:]):
        assert bool(a & 2) == bool(b & 1)
    assert cp.resident and 1 <= cp.n_slots <= len(cp.entries)


def test_stride2_first_layer_read_in_place():
    """'pixrows': conv1 reads 8-pixel runs straight out of a bf16 RGBX plane (overlapping matrix rows)."""
    torch.manual_seed(8)
    B, H, W, k, cout = 2, 23, 30, 5, 16
    x = _bf(torch.randn(B, H, W, 3))
    w = _bf(torch.randn(cout, 3, k, k) * 0.1)
    lin = RowLayout("pixrows", B, H, W, 3, kw=k)
    cp = ConvPlan(lin, w, 2, 0, group="tap")
    assert cp.mma_per_tile == 10
    lout = RowLayout("parity", B, cp.OH, cp.OW, cp.N, pad=0)
    out = torch.zeros(lout.rows, lout.cols, dtype=torch.bfloat16)
    emulate(cp, lin.to_rows(x), out, lout, torch.ones(cout), torch.zeros(cout), relu=False)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2).permute(0, 2, 3, 1)
    assert torch.allclose(lout.from_rows(out, C=cout), ref, atol=3e-2, rtol=3e-2)


def test_detconv_weight_image_roundtrip():
    """the shared-memory weight image of csrc/detconv.cu: pack (vectorised) == element-address formula"""
    from pyannote_video_b200.detconv import pack_weight_image, unpack_weight_image
    w = torch.randn(45, 45, 5, 5).to(torch.bfloat16).float()
    img = pack_weight_image(w, 48, 48)
    assert img.numel() == 25 * 3 * 48 * 32
    u = unpack_weight_image(img, 48, 48, 5, 5)
    assert (u[:45, :45] == w.numpy()).all() and (u[45:] == 0).all() and (u[:, 45:] == 0).all()
    w9 = torch.randn(9, 45, 9, 1).to(torch.bfloat16).float()
    u9 = unpack_weight_image(pack_weight_image(w9, 48, 16), 48, 16, 9, 1)
    assert (u9[:9, :45] == w9.numpy()).all()


def test_rsconv_weight_image_roundtrip():
    """the shared-memory weight image of csrc/rsconv.cu (filter rows of one parity class side by side)"""
    from pyannote_video_b200.detconv import pack_weight_image_rs, unpack_weight_image_rs
    for (cout, cin, kh, kw, c_in, n_out, s) in [(45, 45, 5, 5, 48, 48, 1), (32, 16, 5, 5, 16, 32, 2), (9, 45, 9, 1, 48, 16, 1)]:
        w = torch.randn(cout, cin, kh, kw).to(torch.bfloat16).float()
        img = pack_weight_image_rs(w, c_in, n_out, s)
        assert img.numel() == kh * kw * (c_in // 16) * n_out * 32
        u = unpack_weight_image_rs(img, c_in, n_out, kh, kw, s)
        assert (u[:cout, :cin] == w.numpy()).all() and (u[cout:] == 0).all() and (u[:, cin:] == 0).all()
