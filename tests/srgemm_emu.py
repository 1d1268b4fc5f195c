"""CPU emulator of the srgemm kernel contract (include/pv_b200.h) — test infrastructure only.

It executes a ConvPlan's stage table exactly as the kernel does (slab row offsets, column
segments, packed weights, epilogue row maps) with fp32 accumulation over bf16 operands, so the
planning logic can be verified on CPU against torch.nn.functional.conv2d, and the CUDA kernel can
be verified against it on the GPU.
"""
import numpy as np
import torch


def _rowmap_index(layout, n, y, x):
    return layout.row_index(n, y, x)


def emulate(cp, x_rows, out, lout, scale, shift, relu, resid=None, lres=None, out_f32=False, q_rows=None,
            out_rows_f32=False):
    """x_rows: bf16 [rows, cols] CPU.  Writes into `out` (bf16 matrix of lout, or f32 [B,OH,OW])."""
    lin = cp.lin
    X = x_rows.float()
    Q = q_rows or lin.plane_rows
    N = cp.N
    D = torch.zeros(Q, N, dtype=torch.float32)
    q = torch.arange(Q)
    for st in cp.stages:
        wd = cp.widths[st.cls]
        Wp = cp.w_packed[st.cls].float()
        for t in range(st.n_taps):
            r = q + st.a_row_off + st.tap_rel[t]
            ok = (r >= 0) & (r < lin.rows)
            A = torch.zeros(Q, wd)
            A[ok] = X[r[ok], st.a_col:st.a_col + wd]
            Wt = Wp[st.b_row + t * N: st.b_row + (t + 1) * N]  # [N, wd]
            D += A @ Wt.t()
    sc = torch.zeros(N)
    sh = torch.zeros(N)
    sc[:cp.Cout] = scale.float()
    sh[:cp.Cout] = shift.float()
    Y = D * sc + sh
    qn = q.numpy()
    n = qn // lin.img
    rem = qn - n * lin.img
    y = rem // lin.Wq
    x = rem - y * lin.Wq
    valid = (y < cp.OH) & (x < cp.OW)
    n, y, x = n[valid], y[valid], x[valid]
    Yv = Y[torch.from_numpy(valid)]
    if resid is not None:
        rr = torch.from_numpy(_rowmap_index(lres, n, y, x).astype(np.int64))
        Yv = Yv + resid.float()[rr, :N]
    if relu:
        Yv = torch.clamp(Yv, min=0)
    if out_f32:
        out.view(-1)[torch.from_numpy((n * cp.OH * cp.OW + y * cp.OW + x).astype(np.int64))] = Yv[:, 0]
    else:
        dr = torch.from_numpy(_rowmap_index(lout, n, y, x).astype(np.int64))
        out[dr, :N] = Yv if out_rows_f32 else Yv.to(torch.bfloat16)
    return out
