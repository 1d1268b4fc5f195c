"""GPU: the reference-facing API on the POSITIVE path (detections present), against the oracle composite:
`Face.__call__`, `iterfaces` / `get_landmarks` / `get_embedding`, `process_batch`, `extract_batch`, `FaceTracking`
with a real `Face`, and the CLI chain track -> extract -> cluster on a `synthetic:` video
(pyannote/video/face/face.py:64-132, pyannote/video/face/tracking.py:36-78, scripts/pyannote-face.py:239-314).
"""
import numpy as np
import pytest
import torch

from pyannote_video_b200 import weights as W
from pyannote_video_b200.synth import make_frames

pytestmark = pytest.mark.gpu

H, WD = 180, 240


def _models(frames, n_keep=30):
    """seeded models + a detector threshold that leaves a few dozen candidate cells per frame"""
    from oracle import cpu_ref
    det = W.make_detector(seed=2, score_bias=0.0)
    cpu_ref.set_threads(0)
    o = cpu_ref.Detector(det, bf16=True)
    plane, geo = o.build_plane(frames[0], 1)
    sc = o.scores(plane)
    det["adjust_threshold"] = float(np.quantile(sc, 1 - float(n_keep) / sc.size))
    return det, W.make_shape_predictor(seed=4, stages=5, trees=60), W.make_embedder(seed=3)


def _face(cuda, det, sp, emb, **kw):
    from pyannote_video_b200.face import Face
    return Face(landmarks=sp, embedding=emb, detector=det, device=cuda, max_frames=4, max_faces=64, **kw)


def _oracle_frame(det, sp, emb, rgb, gpu_scores):
    """oracle composite on one frame: decode of the CUDA scores (decode parity is exact; score parity is tested
    separately) -> ERT -> chips -> bf16-faithful embedding"""
    from oracle import nets as onets, pyramid as opyr, landmarks as olm
    geo = opyr.placement_for(rgb.shape[0], rgb.shape[1], 1)
    boxes = opyr.decode(gpu_scores, geo, det["window"], det["adjust_threshold"], det["iou_thresh"], det["covered_thresh"])
    rects = np.asarray([b[:4] for b in boxes], np.int64).reshape(-1, 4)
    parts = olm.ert_predict(sp, rgb, rects)
    chips = olm.extract_chips(rgb, parts)
    emb_ref = onets.embed_forward(emb, onets.normalize_rgb(chips), bf16=True).numpy()
    return rects, parts, emb_ref


def test_face_call_positive_path_matches_oracle(cuda):
    frames = make_frames(2, H, WD, seed=31).numpy()
    det, sp, emb = _models(frames)
    face = _face(cuda, det, sp, emb)
    for rgb in frames:
        out = list(face(rgb, return_landmarks=True, return_embedding=True))
        net = face._detector_for(H, WD)
        scores = net.scores[0].cpu().numpy()
        rects, parts, emb_ref = _oracle_frame(det, sp, emb, rgb, scores)
        assert len(out) == len(rects) and len(out) >= 3
        for k, (f, lm, e) in enumerate(out):
            assert (f.left(), f.top(), f.right(), f.bottom()) == tuple(rects[k])
            assert lm.num_parts == 68 and lm.rect is f
            assert [(p.x, p.y) for p in lm.parts()] == [tuple(v) for v in parts[k].tolist()]
            e = np.asarray(list(e), np.float64)
            assert e.shape == (128, )
            assert np.linalg.norm(e - emb_ref[k]) <= 2e-2 * np.linalg.norm(emb_ref[k])
        # the per-object methods give the same answers as the generator
        faces = list(face.iterfaces(rgb))
        assert [(f.left(), f.top(), f.right(), f.bottom()) for f in faces] == [tuple(r) for r in rects]
        lm0 = face.get_landmarks(rgb, faces[0])
        assert [(p.x, p.y) for p in lm0.parts()] == [tuple(v) for v in parts[0].tolist()]
        e0 = np.asarray(face.get_embedding(rgb, lm0))
        assert np.allclose(e0, np.asarray(list(out[0][2])), atol=1e-6)
        # plain call yields the rectangles only
        assert [type(x).__name__ for x in face(rgb)] == ["Rect"] * len(rects)


def test_process_batch_and_extract_batch_match_oracle(cuda):
    frames = make_frames(6, H, WD, seed=32).numpy()          # 6 frames through max_frames = 4: two chunks
    det, sp, emb = _models(frames)
    face = _face(cuda, det, sp, emb)
    boxes, fidx, scores, parts, e = face.process_batch(frames)
    boxes, fidx, parts, e = boxes.cpu().numpy(), fidx.cpu().numpy(), parts.cpu().numpy(), e.cpu().numpy()
    assert (np.diff(fidx) >= 0).all() and set(fidx.tolist()) == set(range(6))
    net = face._detector_for(H, WD)
    for f in range(6):
        net.detect(torch.from_numpy(frames[f:f + 1]).to(cuda))
        rects, ref_parts, ref_emb = _oracle_frame(det, sp, emb, frames[f], net.scores[0].cpu().numpy())
        sel = fidx == f
        assert np.array_equal(boxes[sel], rects)
        assert np.array_equal(parts[sel], ref_parts)
        assert np.linalg.norm(e[sel] - ref_emb) <= 2e-2 * np.linalg.norm(ref_emb)
    # extract_batch: given boxes (as `extract` does) + padded detections, no host sync
    res = face.extract_batch(frames[:4], boxes[fidx < 4], fidx[fidx < 4])
    assert np.array_equal(res["landmarks"].cpu().numpy(), parts[fidx < 4])
    assert np.allclose(res["embeddings"].cpu().numpy(), e[fidx < 4], atol=1e-6)
    cnt = res["det_counts"].cpu().numpy()
    assert [int(c) for c in cnt] == [int((fidx == f).sum()) for f in range(4)]
    for f in range(4):
        assert np.array_equal(res["det_boxes"][f, :cnt[f]].cpu().numpy(), boxes[fidx == f])
    # upload(): pinned host -> staging ring on the copy stream
    fr_d, bx_d, fi_d, ready = face.upload(torch.from_numpy(frames[:4]).pin_memory(), torch.from_numpy(boxes[fidx < 4]),
                                          torch.from_numpy(fidx[fidx < 4]), slot=1)
    torch.cuda.current_stream().wait_event(ready)
    res2 = face.extract_batch(fr_d, bx_d, fi_d, detect=False)
    assert np.array_equal(res2["landmarks"].cpu().numpy(), parts[fidx < 4])


def test_face_requires_a_detector_model_and_valid_upsample(cuda):
    from pyannote_video_b200.face import Face
    f = Face(device=cuda)                                   # constructible like the reference's Face()
    with pytest.raises(RuntimeError, match="no detector model"):
        list(f.iterfaces(np.zeros((64, 64, 3), np.uint8)))
    with pytest.raises(ValueError):
        Face(detector="synthetic", upsample=2, device=cuda)
    with pytest.raises(RuntimeError):
        Face(detector=W.make_embedder(seed=1), device=cuda)


class _Seg(object):
    def __init__(self, start, end):
        self.start, self.end = start, end


def _video(frames):
    from pyannote_video_b200.cli import ArrayVideo
    return ArrayVideo(frames, 25.0)


def test_face_tracking_with_a_real_face_matches_oracle_pipeline(cuda):
    """FaceTracking(face=Face(...)) on a translating synthetic shot vs the same control loop driven by the oracle:
    oracle decode of the CUDA scores as detections, the numpy DSST oracle as tracker"""
    from oracle import pyramid as opyr
    from oracle.dsst import CorrelationTracker as OracleTracker
    from pyannote_video_b200.geometry import DRect
    from pyannote_video_b200.tracking import FaceTracking, TrackingByDetection, PerObjectBank
    frames = make_frames(8, H, WD, seed=33, shift_per_frame=(2.0, 1.0)).numpy()
    det, sp, emb = _models(frames, n_keep=12)
    face = _face(cuda, det, sp, emb)
    shots = [_Seg(0.0, 10.0)]
    got = list(FaceTracking(face=face, track_min_confidence=3.0, track_max_gap=0.0, detect_every=0.12)(_video(frames), shots))
    assert len(got) >= 2

    net = face._detector_for(H, WD)
    geo = opyr.placement_for(H, WD, 1)

    def oracle_detect(frame):
        rgb = frame if isinstance(frame, np.ndarray) else frame.cpu().numpy()
        net.detect(torch.from_numpy(rgb[None]).to(cuda))
        sc = net.scores[0].cpu().numpy()
        return [b[:4] for b in opyr.decode(sc, geo, det["window"], det["adjust_threshold"], det["iou_thresh"],
                                           det["covered_thresh"])]

    class OT(OracleTracker):
        def start_track(self, frame, rect):
            OracleTracker.start_track(self, np.asarray(frame), (rect.left(), rect.top(), rect.right(), rect.bottom()))

        def update(self, frame):
            return OracleTracker.update(self, np.asarray(frame))

        def get_position(self):
            return DRect(*self.position)

    ref = list(TrackingByDetection(oracle_detect, detect_smallest=36, track_min_confidence=3.0, track_max_gap=0.0,
                                   detect_every=0.12, tracker_bank=PerObjectBank(OT))(_video(frames), shots))
    assert len(got) == len(ref)
    for tg, tr in zip(got, ref):
        assert [(t, s) for t, _, s in tg] == [(t, s) for t, _, s in tr]
        for (_, bg, _), (_, br, _) in zip(tg, tr):
            assert np.allclose(bg, br, atol=1.01 / H)          # integer-rounded boxes: tracker tolerance may flip one pixel


def test_cli_track_extract_cluster_chain(cuda, tmp_path):
    """the verbs of scripts/pyannote-face.py on a synthetic: video, end to end on the GPU, files checked against the API"""
    from pyannote_video_b200 import cli
    from oracle import hac as ohac
    spec = "synthetic:%dx%d:10:34" % (WD, H)
    frames = cli.open_video(spec).frames
    det, sp, emb = _models(frames, n_keep=10)
    W.save_model(str(tmp_path / "det.npz"), det)
    W.save_model(str(tmp_path / "sp.npz"), sp)
    W.save_model(str(tmp_path / "emb.npz"), emb)
    shots = tmp_path / "shots.json"
    shots.write_text('[{"start": 0.0, "end": 0.2}, {"start": 0.2, "end": 0.4}]')
    trk, lmk, embf, lab = (str(tmp_path / n) for n in ("t.track.txt", "t.landmarks.txt", "t.embedding.txt", "t.labels.txt"))
    assert cli.main(["--detector", str(tmp_path / "det.npz"), "track", "--min-confidence", "3", "--max-gap", "0", spec,
                     str(shots), trk]) == 0
    rows = cli.read_track_file(trk)
    assert len(rows) > 10 and all(-0.5 <= r[2] <= 1.5 and -0.5 <= r[3] <= 1.5 for r in rows)   # trackers may drift past the border
    assert all(r[6].split("+")[0] in ("forward", "backward", "detection") for r in rows)
    assert cli.main(["extract", spec, trk, str(tmp_path / "sp.npz"), str(tmp_path / "emb.npz"), lmk, embf]) == 0
    lm_rows = [l.split() for l in open(lmk)]
    em_rows = [l.split() for l in open(embf)]
    assert len(lm_rows) == len(em_rows) == len(rows) and len(lm_rows[0]) == 2 + 136 and len(em_rows[0]) == 2 + 128
    # the files hold what the API computes for the same boxes (scripts/pyannote-face.py:290-311 formats)
    face = _face(cuda, det, sp, emb)
    t0, ident = float(lm_rows[0][0]), int(lm_rows[0][1])
    r0 = [r for r in rows if abs(r[0] - t0) < 1e-6 and r[1] == ident][0]
    box = [int(r0[2] * WD), int(r0[3] * H), int(r0[4] * WD), int(r0[5] * H)]
    fi = int(round(t0 * 25.0))
    parts = face.landmarks_batch(frames[fi], [box], [0])
    e = face.embed_batch(frames[fi], parts, [0]).cpu().numpy()[0]
    assert [("%.5f" % (x / WD), "%.5f" % (y / H)) for x, y in parts[0].cpu().tolist()] == \
        list(zip(lm_rows[0][2::2], lm_rows[0][3::2]))
    assert ["%.5f" % v for v in e] == em_rows[0][2:]
    # cluster verb == greedy oracle on the parsed file (tracks seen at one timestamp are dropped, clustering.py:78-79)
    assert cli.main(["cluster", "--threshold", "0.6", embf, lab]) == 0
    raw = np.loadtxt(embf, ndmin=2)
    labels = dict((int(a), int(b)) for a, b in (l.split() for l in open(lab)))
    multi = [int(t) for t in np.unique(raw[:, 1]) if np.ptp(raw[raw[:, 1] == t, 0]) > 0]
    assert sorted(labels) == sorted(multi)
    keep = np.isin(raw[:, 1].astype(int), multi)
    ref = ohac.greedy_hac(raw[keep, 2:], raw[keep, 1].astype(int), threshold=0.6)
    assert ohac.partition_of(labels) == ohac.partition_of(ref)
