"""GPU: the edge cases of the path — frames without detections, candidate overflow, empty and degenerate clustering
inputs, constant frames, boxes hanging over the frame border, partial batches."""
import numpy as np
import pytest
import torch

from pyannote_video_b200 import weights as W
from pyannote_video_b200.synth import make_frames

pytestmark = pytest.mark.gpu


def _face(cuda, **kw):
    from pyannote_video_b200.face import Face
    return Face(landmarks=W.make_shape_predictor(seed=4, stages=3, trees=20), embedding=W.make_embedder(seed=3),
                device=cuda, max_frames=2, max_faces=8, **kw)


def test_frame_without_detections_yields_nothing(cuda):
    det = W.make_detector(seed=2, score_bias=-50.0)           # nothing can pass the threshold
    face = _face(cuda, detector=det)
    rgb = make_frames(1, 96, 128, seed=3)[0].numpy()
    assert list(face.iterfaces(rgb)) == []
    assert list(face(rgb, return_landmarks=True, return_embedding=True)) == []
    boxes, fidx, scores = face.detect_batch(rgb)
    assert boxes.shape == (0, 4) and fidx.numel() == 0 and scores.numel() == 0


def test_candidate_overflow_is_reported_and_degrades_to_the_best_candidates(cuda):
    det = W.make_detector(seed=2, score_bias=50.0)            # every cell is a candidate
    face = _face(cuda, detector=det)
    rgb = make_frames(1, 160, 200, seed=3)[0].numpy()
    net = face._detector_for(160, 200)
    assert net.OH * net.OW > net.MAX_CAND
    _, _, counts = net.detect(torch.from_numpy(rgb)[None].to(cuda))
    assert int(counts[0]) < 0                                  # include/pv_b200.h: out_counts < 0 reports the overflow
    # the public API keeps the MAX_CAND best cells of the frame and warns instead of aborting the video
    with pytest.warns(UserWarning):
        boxes, fidx, scores = face.detect_batch(rgb)
    assert boxes.shape[0] > 0 and boxes.shape[0] <= net.MAX_DET
    from oracle import pyramid as opyr
    sc = net.scores[0].cpu().numpy()
    kth = float(np.sort(sc.reshape(-1))[-net.MAX_CAND])
    ref = opyr.decode(sc, net.geo, det["window"], kth, det["iou_thresh"], det["covered_thresh"])
    assert [tuple(b) for b in boxes.cpu().tolist()] == [r[:4] for r in ref][:boxes.shape[0]]


def test_constant_frames_are_finite_and_deterministic(cuda):
    from pyannote_video_b200.nets import DetectorNet
    model = W.make_detector(seed=2, score_bias=0.0)
    net = DetectorNet(model, 96, 128, 1, max_batch=2, device=cuda)
    frames = torch.stack([torch.zeros(96, 128, 3, dtype=torch.uint8), torch.full((96, 128, 3), 255, dtype=torch.uint8)]).to(cuda)
    net.build_plane(frames, 2)
    a = net.forward_scores(2).clone()
    net.check()
    assert torch.isfinite(a).all()
    # inside a tile a constant image gives a constant plane; the padding stays exactly zero
    x0, y0, w, h = net.geo.rects[0]
    assert int(net.plane[1, y0:y0 + h, x0:x0 + w, :3].min()) == 255 and int(net.plane[1, :y0].max()) == 0
    net.build_plane(frames, 2)
    assert torch.equal(net.forward_scores(2), a)


def test_landmarks_and_embedding_of_boxes_over_the_border(cuda):
    face = _face(cuda)
    rgb = make_frames(1, 96, 128, seed=5)[0].numpy()
    boxes = [[-30, -20, 40, 50], [100, 60, 170, 130], [10, 10, 11, 11]]      # off the top-left, off the bottom-right, 2x2
    parts = face.landmarks_batch(rgb, boxes, [0, 0, 0])
    assert parts.shape == (3, 68, 2)
    emb = face.embed_batch(rgb, parts, [0, 0, 0])
    assert emb.shape == (3, 128) and torch.isfinite(emb).all()
    # same call with a single box gives the same row (no dependence on batch mates)
    emb1 = face.embed_batch(rgb, parts[1:2], [0])
    assert torch.equal(emb1[0], emb[1])


def test_clustering_degenerate_inputs(cuda):
    from pyannote_video_b200.clustering import cluster
    t, l = cluster(np.zeros((0, 128), np.float32), np.zeros(0, np.int64), device=cuda)
    assert len(t) == 0 and len(l) == 0
    one = np.random.default_rng(0).standard_normal((3, 128)).astype(np.float32)
    t, l = cluster(one, np.array([7, 7, 7]), device=cuda)                    # a single track of three embeddings
    assert t.tolist() == [7] and l.tolist() == [7]
    x = np.stack([one[0], one[0], one[0] + 10.0])                             # two identical embeddings and a far one
    t, l = cluster(x, np.array([0, 1, 2]), threshold=0.6, device=cuda)
    assert t.tolist() == [0, 1, 2] and l.tolist() == [0, 0, 2]
