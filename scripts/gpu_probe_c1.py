"""conv1_fused ablation probe: kernel time with parts of the pipeline switched off (PV_C1_ABLATE bits: 1 no output
stores, 2 no epilogue math/stores, 4 converters store without converting, 8 only the two opening MMAs per tile, 16 converters do not store, 32 no
fence.proxy.async, 64 the epilogue does not read TMEM).
Results of ablated runs are wrong by construction; this only locates the stage that paces the kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pyannote_video_b200 import weights
from pyannote_video_b200.nets import DetectorNet
from pyannote_video_b200.synth import make_frames

dev = torch.device("cuda:0")
B = 8
det = DetectorNet(weights.make_detector(), 1080, 1920, 1, B, dev)
frames = make_frames(B, 1080, 1920, seed=1, device=dev)
det.build_plane(frames, B)
op, img = det.convs[0]
out = {}
for ab in (0, 14, 2, 126):
    os.environ["PV_C1_ABLATE"] = str(ab)
    for _ in range(2):
        op.run(B * img)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        op.run(B * img)
    b.record()
    torch.cuda.synchronize()
    out["ablate_%d" % ab] = round(a.elapsed_time(b) / 5 * 1000, 1)
os.environ["PV_C1_ABLATE"] = "0"
print(json.dumps(out))
