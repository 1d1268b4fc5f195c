"""GPU probe: per-layer times of the CNN detector on 1080p frames (CUDA events around every launch
group), for one first-conv mode (env PV_DET_CONV1).  With --once it runs a single forward so that
an ncu capture of one kernel stays cheap.  Usage: python scripts/gpu_probe_det.py [--frames 8] [--once]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from pyannote_video_b200 import config, weights
from pyannote_video_b200.nets import DetectorNet
from pyannote_video_b200.synth import make_frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.frames
    det = DetectorNet(weights.make_detector(), 1080, 1920, 1, B, dev)
    frames = make_frames(B, 1080, 1920, seed=1, device=dev)
    det.detect(frames)
    det.check()
    if args.once:
        torch.cuda.synchronize()
        return
    names = list(det.layer_names)
    acc = {n: 0.0 for n in names}
    acc["pyramid"] = 0.0
    acc["forward_scores"] = 0.0
    for r in range(args.reps):
        evs = []
        for (op, img), n in zip(det.convs, names):
            def timed(q_rows=None, _run=op.run, _n=n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _run(q_rows)
                b.record()
                evs.append((_n, a, b))
            op._saved_run = op.run
            op.run = timed
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a0.record()
        det.build_plane(frames, B)
        a1.record()
        det.forward_scores(B)
        a2.record()
        torch.cuda.synchronize()
        for (op, img) in det.convs:
            op.run = op._saved_run
        for n, a, b in evs:
            acc[n] += a.elapsed_time(b)
        acc["pyramid"] += a0.elapsed_time(a1)
        acc["forward_scores"] += a1.elapsed_time(a2)
    out = {k: round(v / args.reps * 1000.0, 1) for k, v in acc.items()}
    if os.environ.get("PV_RS_DEBUG") and det.conv_impl == "rsconv":
        import ctypes as C
        from pyannote_video_b200 import _lib
        dbg = {}
        for (op, _), n in zip(det.convs[1:], names[1:]):
            if not hasattr(op, "_create") or op._create != "pv_rsconv_create":
                continue
            a = (C.c_longlong * 8)()
            _lib.check(_lib.lib().pv_rsconv_debug(op.h, a), "pv_rsconv_debug")
            rows = max(1, a[3])
            dbg[n] = dict(rows=a[3], wait_slot=round(a[0] / rows), wait_data=round(a[1] / rows), issue=round(a[2] / rows),
                          epi_wait=round(a[4] / rows), epi_work=round(a[5] / rows), total_per_row=round(a[6] / rows))
        out["rs_debug_cycles_per_input_row"] = dbg
    if os.environ.get("PV_C1_DEBUG") and det.conv1_mode == "fused":
        import ctypes as C
        from pyannote_video_b200 import _lib
        det.convs[0][0].run(B * det.convs[0][1])
        torch.cuda.synchronize()
        a = (C.c_longlong * 8)()
        _lib.check(_lib.lib().pv_conv1_debug(a), "pv_conv1_debug")
        t = max(1, a[3])
        out["c1_debug_cycles_per_tile"] = dict(tiles=a[3], mma_wait_acc=round(a[0] / t), mma_wait_px=round(a[1] / t), mma_issue=round(a[2] / t),
                                               conv_wait_raw=round(a[4] / t), conv_wait_slot=round(a[5] / t), conv_work=round(a[6] / t),
                                               epi_wait=round(a[7] / t))
    if os.environ.get("PV_C12_DEBUG") and det.conv1_mode == "c12":
        op = det.convs[0][0]
        d = op.debug()
        q, r = max(1, d["c1_quads"]), max(1, d["c2_rows"])
        out["c12_debug_cycles_per_quad"] = {k: round(v / q) for k, v in d.items() if k not in ("c1_quads", "c2_rows")}
        out["c12_debug_cycles_per_quad"].update(quads=d["c1_quads"], conv2_rows=d["c2_rows"], info=op.info())
    out["mode"] = det.conv1_mode
    out["frames"] = B
    out["unit"] = "us"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
