"""GPU: the CUDA path against the committed golden vectors (tests/golden/path_golden.npz): byte / integer
stages bit-exact, bf16 networks within the stated tolerances."""
import hashlib
import os

import numpy as np
import pytest
import torch

from pyannote_video_b200 import weights as W
from pyannote_video_b200.synth import make_frames, make_boxes

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("conv_impl", ["rsconv", "detconv", "srgemm"])
def test_cuda_path_matches_golden(cuda, conv_impl):
    from pyannote_video_b200.nets import DetectorNet, EmbedNet
    from pyannote_video_b200.ops import ShapePredictor, ChipExtractor
    g = np.load(os.path.join(HERE, "golden", "path_golden.npz"))
    H, Wd, F = 96, 128, 2
    frames = make_frames(F, H, Wd, seed=5)
    fd = frames.to(cuda)
    model = W.make_detector(seed=2, score_bias=0.0)
    det = DetectorNet(model, H, Wd, 1, max_batch=F, device=cuda, conv_impl=conv_impl)
    det.build_plane(fd, F)
    plane = det.plane[:F].cpu().numpy()
    assert tuple(g["plane_shape"]) == plane.shape[1:]
    assert np.array_equal(np.frombuffer(hashlib.sha256(plane.tobytes()).digest(), np.uint8), g["plane_sha256"])
    scores = det.forward_scores(F).cpu().numpy()
    det.check()
    assert np.abs(scores - g["scores"]).max() < 0.03 * max(1.0, float(np.abs(g["scores"]).max()))
    # decode the GOLDEN scores on the GPU: boxes must be identical
    det.scores[:F].copy_(torch.from_numpy(g["scores"]).to(cuda))
    m2 = dict(model)
    m2["adjust_threshold"] = float(g["decode_threshold"])
    det.model = m2
    boxes, _, counts = det.decode(F)
    boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
    for i, key in enumerate(("boxes0", "boxes1")):
        assert counts[i] == len(g[key]) and np.array_equal(boxes[i, :counts[i]], g[key])
    sp = ShapePredictor(W.make_shape_predictor(seed=4, stages=4, trees=40), cuda)
    bx, fidx = make_boxes(F, 2, H, Wd, seed=1, min_side=30, max_side=70)
    parts = sp.predict(fd, bx.to(cuda), fidx.to(cuda))
    assert np.array_equal(parts.cpu().numpy(), g["landmarks"])
    net = EmbedNet(W.make_embedder(seed=3), max_batch=4, device=cuda)
    M = bx.shape[0]
    ChipExtractor(cuda).extract(fd, parts, fidx.to(cuda), net.chips)
    chips = net.chips[:M, :, :, :3].cpu().numpy()
    assert np.array_equal(np.frombuffer(hashlib.sha256(np.ascontiguousarray(chips).tobytes()).digest(), np.uint8), g["chips_sha256"])
    emb = net.forward_chips(M).cpu().numpy()
    net.check()
    rel_bf = np.linalg.norm(emb - g["embedding_bf16"]) / np.linalg.norm(g["embedding_bf16"])
    rel_32 = np.linalg.norm(emb - g["embedding_fp32"]) / np.linalg.norm(g["embedding_fp32"])
    assert rel_bf < 2e-2 and rel_32 < 6e-2, (rel_bf, rel_32)
