"""The C++ CPU restatement (oracle/cpu/*.cpp, the timed CPU baseline of bench.py) against the numpy / torch
oracle (oracle/*.py): byte and integer stages bit-exact, convolution stacks within float rounding, tracker
within 1e-6 (double-precision spectra on both sides).  Two independently written restatements of the same
recalled dlib algorithms agreeing is what this pins; neither is dlib (parity unpinned)."""
import numpy as np
import torch

from oracle import cpu_ref, nets as onets, pyramid as opyr, landmarks as olm, dsst as odsst
from pyannote_video_b200 import weights as W
from pyannote_video_b200.synth import make_frames, make_boxes


def test_library_builds_and_threads():
    cpu_ref.lib()
    assert cpu_ref.set_threads(1) == 1
    assert cpu_ref.set_threads(0) >= 1
    assert cpu_ref.host_cores() >= 1


def test_plane_bit_exact_and_scores_close():
    frames = make_frames(2, 90, 130, seed=11).numpy()
    model = W.make_detector(seed=2, score_bias=0.0)
    det = cpu_ref.Detector(model)
    for up in (1, 0):
        for i in range(2):
            plane, geo = det.build_plane(frames[i], up)
            ref_plane, ref_geo = opyr.build_plane(frames[i], up)
            assert np.array_equal(plane, ref_plane)
            s = det.scores(plane)
            ref = onets.detector_forward(model, torch.from_numpy(opyr.normalize_plane(ref_plane))[None])[0].numpy()
            assert s.shape == ref.shape
            assert np.abs(s - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
            thr = float(np.quantile(ref, 1 - 40.0 / ref.size))
            a = det.decode(ref, geo, threshold=thr)
            b = opyr.decode(ref, ref_geo, model["window"], thr, model["iou_thresh"], model["covered_thresh"])
            assert len(a) > 3 and [x[:4] for x in a] == [tuple(x[:4]) for x in b]


def test_scores_bf16_mode_matches_bf16_oracle():
    frames = make_frames(1, 80, 96, seed=3).numpy()
    model = W.make_detector(seed=2, score_bias=0.0)
    det = cpu_ref.Detector(model, bf16=True)
    plane, geo = det.build_plane(frames[0], 1)
    s = det.scores(plane)
    ref = onets.detector_forward(model, torch.from_numpy(opyr.normalize_plane(plane))[None], bf16=True)[0].numpy()
    assert np.abs(s - ref).max() < 2e-2 * max(1.0, np.abs(ref).max())      # bf16 roundings flip on 1-ulp fp32 differences


def test_landmarks_and_chips_bit_exact():
    H, Wd = 120, 160
    frames = make_frames(2, H, Wd, seed=5).numpy()
    sp = W.make_shape_predictor(seed=4, stages=5, trees=50)
    boxes, fidx = make_boxes(2, 3, H, Wd, seed=1, min_side=30, max_side=90)
    cp = cpu_ref.ShapePredictor(sp)
    for f in range(2):
        sel = (fidx == f).numpy()
        a = cp.predict(frames[f], boxes[sel].numpy())
        b = olm.ert_predict(sp, frames[f], boxes[sel].numpy())
        assert np.array_equal(a, b)
        assert np.array_equal(cpu_ref.extract_chips(frames[f], a), olm.extract_chips(frames[f], b))


def test_embedding_close():
    chips = make_frames(3, 150, 150, seed=9).numpy()
    model = W.make_embedder(seed=3)
    ref = onets.embed_forward(model, onets.normalize_rgb(chips)).numpy()
    out = cpu_ref.Embedder(model).forward(chips)
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-4
    ref16 = onets.embed_forward(model, onets.normalize_rgb(chips), bf16=True).numpy()
    out16 = cpu_ref.Embedder(model, bf16=True).forward(chips)
    assert np.linalg.norm(out16 - ref16) / np.linalg.norm(ref16) < 1e-2


def test_tracker_matches_numpy_oracle():
    H, Wd = 200, 260
    frames = make_frames(4, H, Wd, seed=2, shift_per_frame=(2.0, 1.0)).numpy()
    rect = (100.0, 70.0, 160.0, 130.0)
    ref = odsst.CorrelationTracker(use_scale=True)
    ref.start_track(frames[0], rect)
    bank = cpu_ref.TrackerBank(2, use_scale=True)
    bank.start(frames[0], [1], [rect])
    for t in range(1, 4):
        psr_ref = ref.update(frames[t])
        psr = bank.update(frames[t], [1])[0]
        assert abs(psr - psr_ref) < 1e-6 * max(1.0, abs(psr_ref))
        assert np.allclose(bank.position(1), ref.get_position(), atol=1e-6)
    assert abs((bank.position(1)[0] - rect[0]) + 6.0) < 0.5      # canvas moves +2 px / frame -> content moves -2
