"""GPU probe: per-update PSR / position of the CUDA tracker bank next to the numpy oracle (translation filter only and with
the scale filter) on the frames of tests/test_tracker_gpu.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle.dsst import CorrelationTracker as OracleTracker
from pyannote_video_b200.geometry import DRect
from pyannote_video_b200.synth import make_frames
from pyannote_video_b200.tracker import TrackerBank

dev = torch.device("cuda:0")
frames = make_frames(9, 360, 640, seed=3, shift_per_frame=(2.0, 1.0))
rects = [(200.0, 100.0, 296.0, 196.0), (400.5, 150.25, 460.5, 230.0)]
for use_scale in (False, True):
    bank = TrackerBank(capacity=8, device=dev, use_scale=use_scale) if "use_scale" in TrackerBank.__init__.__code__.co_varnames else TrackerBank(capacity=8, device=dev)
    dfs = [bank.prepare_frame(f) for f in frames]
    hs = [bank.start(dfs[0], DRect(*r)) for r in rects]
    orc = []
    for r in rects:
        t = OracleTracker(use_scale=use_scale)
        t.start_track(frames[0].numpy(), r)
        orc.append(t)
    for i in range(1, 9):
        conf = bank.update(dfs[i], hs)
        for k, t in enumerate(orc):
            ref = t.update(frames[i].numpy())
            rp = np.array(t.get_position())
            g = bank.position(hs[k])
            gp = np.array([g.left(), g.top(), g.right(), g.bottom()])
            print("scale=%d i=%d k=%d psr gpu %.4f ref %.4f rel %.5f dpos %.5f" % (use_scale, i, k, conf[k], ref, (conf[k] - ref) / ref, np.abs(gp - rp).max()), flush=True)
