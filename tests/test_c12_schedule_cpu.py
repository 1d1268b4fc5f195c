"""CPU: numpy emulation of the fused conv1+conv2 strip kernel's schedule (csrc/c12.cu) — the conv1 weight image (tiles M0 /
M1 / P, 32-byte swizzle), the per-quad MMA sequence with its row clamps, the overlapping A rows of the pixel-row buffer
and the hand-swizzled conv2 A operand the conv1 epilogue writes — against a direct convolution."""
import numpy as np
import torch

from pyannote_video_b200.detconv import pack_c12_w1, unpack_c12_w1, C12_W1_TILES

T1 = 126          # conv1 outputs per M tile
PXW = 520         # pixels per pixel-row buffer row


def _a_main(px, t, j):
    """A operand of the main MMA of plane row t, tile j: row m = pixels 2m .. 2m+3 (k = kw*4 + c)"""
    A = np.zeros((128, 16), np.float32)
    for m in range(128):
        for k in range(16):
            A[m, k] = px[t, 2 * T1 * j + 2 * m + (k >> 2), k & 3]
    return A


def _a_pair(px, t_even, j):
    """A operand of the pair MMA: k < 8 -> pixels 2m+4, 2m+5 of the even row, k >= 8 -> of the odd row"""
    A = np.zeros((128, 16), np.float32)
    for m in range(128):
        for k in range(16):
            A[m, k] = px[t_even + (k >> 3), 2 * T1 * j + 2 * m + 4 + ((k & 7) >> 2), k & 3]
    return A


def _issue(acc, opened, A, tile, r_first, r_max, new_row):
    """c1_issue: rows [r_first, r_first + nblk) take blocks 0..nblk-1; clamp to [0, r_max]; new_row starts from zero"""
    nblk = tile.shape[0]
    lo, hi = max(r_first, 0), min(r_first + nblk - 1, r_max)
    for r in range(lo, hi + 1):
        d = A @ tile[r - r_first].T                      # [128, 16]
        if r == new_row:
            assert r not in opened
            acc[r] = d.copy()
            opened.add(r)
        else:
            assert r in opened, "accumulating into a row that was never opened"
            acc[r] += d


def test_conv1_weight_image_roundtrip_and_schedule():
    rng = np.random.default_rng(0)
    w = torch.from_numpy(rng.standard_normal((16, 3, 5, 5)).astype(np.float32)).to(torch.bfloat16).float()
    img = pack_c12_w1(w)
    tiles = unpack_c12_w1(img)
    assert [k for k, _ in C12_W1_TILES] == list(tiles)
    L = 3                                    # conv2 rows of the item
    r_max = 2 * L + 2
    nq = L + 3
    px = np.zeros((4 * nq, PXW, 4), np.float32)
    px[:, :512, :3] = rng.standard_normal((4 * nq, 512, 3)).astype(np.float32)
    wn = w.numpy()
    for j in range(2):
        acc, opened, done = {}, set(), []
        for q in range(nq):
            for tt in range(4):
                t = 4 * q + tt
                r_top = 2 * q + (tt >> 1)
                if tt % 2 == 0:
                    opens = r_top <= r_max
                    _issue(acc, opened, _a_main(px, t, j), tiles["M0"], r_top - 2, r_max, r_top if opens else -1)
                else:
                    _issue(acc, opened, _a_main(px, t, j), tiles["M1"], r_top - 1, r_max, -1)
                    _issue(acc, opened, _a_pair(px, t - 1, j), tiles["P"], r_top - 2, r_max, -1)
                    rc = r_top - 2
                    if 0 <= rc <= r_max:
                        done.append(rc)
                        # the completed row must equal the direct convolution
                        ref = np.zeros((T1, 16), np.float32)
                        for m in range(T1):
                            x = T1 * j + m
                            patch = px[2 * rc:2 * rc + 5, 2 * x:2 * x + 5, :3]          # [kh, kw, c]
                            ref[m] = np.einsum("hwc,nchw->n", patch, wn)
                        np.testing.assert_allclose(acc[rc][:T1], ref, rtol=1e-4, atol=1e-4)
        assert done == list(range(r_max + 1))


def test_conv2_a_operand_swizzle_written_by_the_conv1_epilogue():
    """the epilogue's two 16-byte stores per conv1 column land where a 64-byte-swizzled K-major descriptor starting at
    pair row (kw >> 1), byte (kw & 1) * 32 reads channel chunk c of conv1 column 2m + kw"""
    entry = np.full(8704, -1, np.int64)                   # byte -> (conv1 column, channel) id
    for j in range(2):
        for m in range(T1):
            x1 = j * T1 + m
            pr = x1 >> 1
            sw = (pr >> 1) & 3
            for c in range(2):
                off = pr * 64 + ((((x1 & 1) * 2 + c) ^ sw) << 4)
                assert (entry[off:off + 16] == -1).all()
                for b in range(16):
                    entry[off + b] = x1 * 16 + c * 8 + b // 2
    for kw in range(5):
        for m in range(124):
            for chunk in range(2):
                logical = (m + (kw >> 1)) * 64 + (kw & 1) * 32 + chunk * 16
                phys = logical ^ (((logical >> 7) & 3) << 4)
                want = (2 * m + kw) * 16 + chunk * 8
                assert entry[phys] == want and entry[phys + 15] == want + 7, (kw, m, chunk)
