"""CPU: the row-streaming schedule of csrc/rsconv.cu and the multi-row MMA schedule of csrc/conv1_fused.cu, emulated in
numpy with the same index algebra as the device code (slot ring, parity classes, new-row MMA with accumulate = 0, ring
wrap splits, row-complete commits) and compared with a direct convolution.  Guards the algorithms without a GPU."""
import numpy as np
import pytest


def conv_direct(x, w, stride, pad_y, pad_x):
    """x [H,W,C], w [N,C,KH,KW] -> [OH,OW,N] (float64)"""
    H, W, C = x.shape
    N, _, KH, KW = w.shape
    xp = np.zeros((H + 2 * pad_y, W + 2 * pad_x, C))
    xp[pad_y:pad_y + H, pad_x:pad_x + W] = x
    OH, OW = (H + 2 * pad_y - KH) // stride + 1, (W + 2 * pad_x - KW) // stride + 1
    out = np.zeros((OH, OW, N))
    for kh in range(KH):
        for kw in range(KW):
            patch = xp[kh:kh + stride * (OH - 1) + 1:stride, kw:kw + stride * (OW - 1) + 1:stride]
            out += patch @ w[:, :, kh, kw].T
    return out


def rsconv_emulate(x, w, S, seg_rows, nslot):
    """mirror of the MMA-issuer / epilogue loops of rsconv_kernel for one 128-column strip (all columns at once here)"""
    H, W, C = x.shape
    N, _, KH, KW = w.shape
    P = KH // 2 if S == 1 else 0
    PX = KW // 2 if S == 1 else 0
    OH, OW = (H + 2 * P - KH) // S + 1, (W + 2 * PX - KW) // S + 1
    nq0 = (KH + S - 1) // S
    nq1 = KH // 2 if S == 2 else 0
    dc = (KH - 1) // S
    out = np.full((OH, OW, N), np.nan)
    slots = [None] * nslot            # accumulator of the row currently owning the slot
    owner = [None] * nslot
    drained = [True] * nslot
    slot_base = 0

    def in_row(y):                    # input row y as the A operand sees it: [OW, KW, C] windows (zero fill outside)
        a = np.zeros((OW, KW, C))
        if 0 <= y < H:
            for kw in range(KW):
                xs = np.arange(OW) * S + kw - PX
                ok = (xs >= 0) & (xs < W)
                a[ok, kw] = x[y, xs[ok]]
        return a

    for ra in range(0, OH, seg_rows):
        rb = min(ra + seg_rows, OH)
        t0, t1 = S * ra, S * (rb - 1) + KH - 1
        slot_top = slot_base
        for t in range(t0, t1 + 1):
            y = t - P
            q = (t & 1) if S == 2 else 0
            odd = S == 2 and q == 1
            nq = nq1 if odd else nq0
            r_top = (t >> 1) if S == 2 else t
            r_first = r_top - (nq - 1)
            r_hi, r_lo = min(r_top, rb - 1), max(r_first, ra)
            new_row = q == 0 and r_top <= rb - 1
            n_all = r_hi - r_lo + 1
            assert n_all >= 1
            slot_lo = slot_top - (r_top - r_lo)
            if slot_lo < 0:
                slot_lo += nslot
            n1 = min(n_all, nslot - slot_lo)
            n2 = n_all - n1
            if new_row:
                assert drained[slot_top], "slot of the new row has not been drained"
                slots[slot_top] = None
                owner[slot_top] = r_top
                drained[slot_top] = False
            a = in_row(y)
            khs = list(range(q, KH, S))[::-1]                       # B blocks: decreasing kh
            blk0 = r_lo - r_first
            first = True
            for kw in range(KW):
                # (chunks of 16 channels are summed in one go here)
                for seg_start, seg_n, seg_slot in ((0, n1, slot_lo), (n1, n2, 0)):
                    for i in range(seg_n):
                        r = r_lo + seg_start + i
                        s = seg_slot + i
                        assert owner[s] == r, "MMA writes a slot that belongs to another row"
                        kh = khs[blk0 + seg_start + i]
                        assert kh == t - S * r
                        contrib = a[:, kw] @ w[:, :, kh, kw].T
                        if first and new_row and r == r_top:
                            assert slots[s] is None
                            slots[s] = contrib.copy()               # the accumulate = 0 MMA
                        else:
                            assert slots[s] is not None, "accumulate into a slot that was never opened"
                            slots[s] += contrib
                first = False
            if S == 1 or q == 0:
                rc = r_top - dc
                if ra <= rc < rb:
                    sc = slot_top - dc
                    if sc < 0:
                        sc += nslot
                    assert owner[sc] == rc
                    out[rc] = slots[sc]                             # epilogue drains the slot
                    drained[sc] = True
            if S == 1 or q == 1:
                slot_top = (slot_top + 1) % nslot
        slot_base = (slot_base + (rb - ra)) % nslot
    return out


@pytest.mark.parametrize("S,KH,KW,nslot,seg_rows,H,W", [
    (1, 5, 5, 10, 7, 23, 17), (1, 5, 5, 10, 26, 40, 9), (1, 5, 5, 10, 100, 31, 12), (1, 9, 1, 16, 32, 45, 10),
    (2, 5, 5, 8, 8, 37, 21), (2, 5, 5, 16, 5, 53, 30), (2, 5, 5, 8, 100, 29, 19),
])
def test_rsconv_schedule_equals_direct_convolution(S, KH, KW, nslot, seg_rows, H, W):
    rng = np.random.default_rng(0)
    C, N = 3, 4
    x = rng.standard_normal((H, W, C))
    w = rng.standard_normal((N, C, KH, KW))
    P, PX = (KH // 2, KW // 2) if S == 1 else (0, 0)
    ref = conv_direct(x, w, S, P, PX)
    got = rsconv_emulate(x, w, S, seg_rows, nslot)
    assert not np.isnan(got).any()
    assert np.allclose(got, ref, atol=1e-9)


def test_conv1_multi_row_schedule_covers_every_tap_once():
    """the 17 MMAs of a conv1_fused tile (4 output rows, 11 input rows): ent_* helpers restated"""
    k_rows, k_in = 4, 11

    def pair(e):
        return e >= 11

    def row(e):
        return 4 if e == 0 else (10 if e == 1 else ((e - 2 if e - 2 < 4 else e - 1) if e < 11 else 2 * (e - 11)))

    def rlo(e):
        return (row(e) - 3) >> 1 if row(e) - 3 > 0 else 0

    def rhi(e):
        return min((row(e) + (1 if pair(e) else 0)) >> 1, k_rows - 1)

    cover, first = {}, {}
    for e in range(17):
        for r in range(rlo(e), rhi(e) + 1):
            first.setdefault(r, e)
            for k in range(16):
                ch = k & 3
                if ch == 3:
                    continue
                if not pair(e):
                    kw, kh, ok = k >> 2, row(e) - 2 * r, True
                else:
                    ii = row(e) + (k >> 3)
                    kw, kh, ok = 4 + ((k >> 2) & 1), ii - 2 * r, (((k >> 2) & 1) == 0 and ii < k_in)
                if ok and 0 <= kh < 5:
                    cover[(r, kh, kw, ch)] = cover.get((r, kh, kw, ch), 0) + 1
    assert len(cover) == 4 * 5 * 5 * 3 and set(cover.values()) == {1}
    assert first == {0: 0, 1: 0, 2: 0, 3: 1}        # the two opening MMAs (accumulate = 0) touch every accumulator first
