"""Generates tests/golden/path_golden.npz: outputs of the CPU oracle (oracle/*.py) for the face path
(pyramid plane -> detector scores -> decode; ERT landmarks -> chips -> embedding) on small seeded inputs
with the seeded synthetic models of pyannote_video_b200.weights.

The reference's own arithmetic lives in dlib 19.12, which is absent here together with its .dat weights
(SURVEY.md §8c), so these vectors pin OUR restatement against drift (a refactor of the oracle or of the shared
host logic — pyramid packing, chip geometry — that changes results shows up as a golden mismatch on CPU,
before any GPU run); they do not pin parity with dlib.  Both sides of the GPU parity tests are also checked
against them (tests/test_golden_cpu.py, tests/test_golden_gpu.py).

    python tests/golden/make_path_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from pyannote_video_b200 import weights as W                       # noqa: E402
from pyannote_video_b200.synth import make_frames, make_boxes      # noqa: E402
from oracle import nets as onets, pyramid as opyr, landmarks as olm  # noqa: E402

H, WD, F = 96, 128, 2
SEEDS = dict(frames=5, detector=2, shape=4, embedder=3, boxes=1)


def compute():
    frames = make_frames(F, H, WD, seed=SEEDS["frames"])
    det = W.make_detector(seed=SEEDS["detector"], score_bias=0.0)
    out = {}
    planes, scores = [], []
    for i in range(F):
        plane, geo = opyr.build_plane(frames[i].numpy(), 1)
        planes.append(plane)
        s = onets.detector_forward(det, torch.from_numpy(opyr.normalize_plane(plane))[None], bf16=True)[0]
        scores.append(s.numpy())
    out["plane_shape"] = np.asarray(planes[0].shape, np.int64)
    out["plane_sha256"] = np.frombuffer(hashlib.sha256(np.stack(planes).tobytes()).digest(), np.uint8)
    out["plane_rows_sample"] = np.stack(planes)[:, ::37, ::41].copy()           # a sparse sample, human-checkable
    out["scores"] = np.stack(scores).astype(np.float32)
    thr = float(np.quantile(out["scores"], 1 - 60.0 / out["scores"][0].size))
    out["decode_threshold"] = np.float32(thr)
    dec = []
    for i in range(F):
        d = opyr.decode(out["scores"][i], geo, det["window"], thr, det["iou_thresh"], det["covered_thresh"])
        dec.append(np.asarray([[r[0], r[1], r[2], r[3]] for r in d], np.int32).reshape(-1, 4))
    out["boxes0"], out["boxes1"] = dec
    sp = W.make_shape_predictor(seed=SEEDS["shape"], stages=4, trees=40)
    boxes, fidx = make_boxes(F, 2, H, WD, seed=SEEDS["boxes"], min_side=30, max_side=70)
    parts, chips = [], []
    for f in range(F):
        sel = (fidx == f).numpy()
        p = olm.ert_predict(sp, frames[f].numpy(), boxes[sel].numpy())
        parts.append(p)
        chips.append(olm.extract_chips(frames[f].numpy(), p))
    out["landmarks"] = np.concatenate(parts).astype(np.int32)
    chips = np.concatenate(chips)
    out["chips_sha256"] = np.frombuffer(hashlib.sha256(chips.tobytes()).digest(), np.uint8)
    emb = W.make_embedder(seed=SEEDS["embedder"])
    out["embedding_bf16"] = onets.embed_forward(emb, onets.normalize_rgb(chips), bf16=True).numpy().astype(np.float32)
    out["embedding_fp32"] = onets.embed_forward(emb, onets.normalize_rgb(chips), bf16=False).numpy().astype(np.float32)
    return out


if __name__ == "__main__":
    o = compute()
    np.savez_compressed(os.path.join(HERE, "path_golden.npz"), **o)
    print({k: (v.shape, str(v.dtype)) for k, v in o.items()})
