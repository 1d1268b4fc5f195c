"""Seeded synthetic video frames (there is no video, ffmpeg or dataset in the build environment).

Frames are low-pass filtered noise plus a linear gradient so that FHOG / ERT / the CNNs see
structure.  Generated with torch on whichever device is asked for; the same seed
gives the same frames on CPU and GPU only per device type, so parity tests always move the SAME
tensor to both sides.
"""
import torch


def make_frames(n, H, W, seed=0, device="cpu", shift_per_frame=(0.0, 0.0)):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pad = 8 + int(abs(shift_per_frame[0]) * n + abs(shift_per_frame[1]) * n) + 1
    base = torch.rand(1, 3, (H + 3) // 4 + pad, (W + 3) // 4 + pad, generator=g, device=device)
    base = torch.nn.functional.interpolate(base, scale_factor=4, mode="bilinear", align_corners=False)
    fine = torch.rand(1, 3, H + 4 * pad, W + 4 * pad, generator=g, device=device)
    canvas = 0.75 * base[:, :, :H + 4 * pad, :W + 4 * pad] + 0.25 * torch.nn.functional.avg_pool2d(fine, 5, 1, 2)
    yy = torch.linspace(0, 1, H, device=device)[None, None, :, None]
    xx = torch.linspace(0, 1, W, device=device)[None, None, None, :]
    out = torch.empty(n, H, W, 3, dtype=torch.uint8, device=device)
    for i in range(n):
        dx = int(round(shift_per_frame[0] * i))
        dy = int(round(shift_per_frame[1] * i))
        crop = canvas[:, :, dy:dy + H, dx:dx + W]
        jitter = 0.02 * torch.rand(1, 3, H, W, generator=g, device=device)
        img = (0.8 * crop + 0.15 * (0.5 * yy + 0.5 * xx) + jitter).clamp(0, 1)
        out[i] = (img[0].permute(1, 2, 0) * 255.0).to(torch.uint8)
    return out


def make_boxes(n_frames, per_frame, H, W, seed=1, min_side=80, max_side=400):
    """seeded square face boxes fully inside the frame: int32 [n_frames*per_frame, 4] (l,t,r,b) and
    frame indices int32 [n_frames*per_frame] (mirrors `extract`, where boxes come from the track file)."""
    g = torch.Generator()
    g.manual_seed(seed)
    M = n_frames * per_frame
    max_side = min(max_side, H - 2, W - 2)
    side = torch.randint(min_side, max_side + 1, (M,), generator=g)
    l = (torch.rand(M, generator=g) * (W - side).float()).long()
    t = (torch.rand(M, generator=g) * (H - side).float()).long()
    boxes = torch.stack([l, t, l + side - 1, t + side - 1], dim=1).to(torch.int32)
    fidx = torch.arange(n_frames, dtype=torch.int32).repeat_interleave(per_frame)
    return boxes, fidx
