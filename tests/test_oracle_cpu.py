"""CPU: weight-free known-answer tests that pin the oracle restatements (the reference ships no
tests or golden vectors — SURVEY.md §4 — so these are the pins we can have without dlib), plus
host-side logic and the C-ABI export check."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pyannote_video_b200 import weights as W
from pyannote_video_b200.geometry import Rect, DRect, match_overlap
from pyannote_video_b200.pyrgeom import pyramid_geometry, det_cell_to_plane
from oracle import nets as onets, pyramid as opyr, landmarks as olm, hac as ohac

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "pv_b200.h")).read()
    names = sorted(set(re.findall(r"\b(pv_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "pyannote_video_b200", "libpvb200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.pv_version() >= 100


def test_no_cpu_fallback_without_device():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyannote_video_b200.face import Face
    with pytest.raises(RuntimeError):
        Face()
    from pyannote_video_b200.clustering import cluster
    with pytest.raises(RuntimeError):
        cluster(np.zeros((4, 128), np.float32), np.arange(4))


# ---------------------------------------------------------------- networks
def test_detector_delta_kernels_shift_input():
    """identity affine + delta kernels: the stack reduces to strided sub-sampling of the input."""
    m = W.make_detector(seed=0)
    for i, c in enumerate(m["convs"]):
        cout, cin, k, s = W.DET_CONVS[i]
        c["w"][:] = 0
        for o in range(min(cout, cin)):
            c["w"][o, o, k // 2, k // 2] = 1.0
        c["b"][:] = 0
        c["gamma"][:] = 1
        c["beta"][:] = 0
    x = torch.rand(1, 3, 61, 77)
    y = onets.detector_forward(m, x)
    # centre taps: conv_i output(p) = input(s*p - pad + k//2); three stride-2 layers: 8p+14
    cx, cy = det_cell_to_plane(np.arange(y.shape[2]), np.arange(y.shape[1]))
    assert torch.allclose(y[0], x[0, 0][cy][:, cx], atol=1e-6)
    assert y.shape[1:] == (onets.detector_out_size(61), onets.detector_out_size(77))


def test_embedder_shapes_and_zero_extension():
    m = W.make_embedder(seed=1)
    x = torch.randn(2, 3, 150, 150) * 0.2
    out, taps = onets.embed_forward(m, x, return_taps=True)
    assert out.shape == (2, 128)
    sizes = [taps[k].shape[-1] for k in ["conv1", "pool1", "block2", "block3", "block7", "block10", "block13"]]
    assert sizes == [72, 35, 35, 17, 8, 4, 2]
    assert len(m["blocks"]) == 14 and 1 + 2 * len(m["blocks"]) == 29
    a, b = torch.ones(1, 2, 3, 3), torch.ones(1, 4, 4, 4)
    z = onets._zero_extend_add(a, b)
    assert z.shape == (1, 4, 4, 4) and z[0, 0, 0, 0] == 2 and z[0, 3, 0, 0] == 1 and z[0, 0, 3, 3] == 1


# ---------------------------------------------------------------- pyramid / decode
def test_resize_known_answers():
    img = np.full((9, 7, 3), 77, np.uint8)
    assert (opyr.resize_bilinear_u8(img, 5, 4) == 77).all()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (6, 8, 3), dtype=np.uint8)
    assert np.array_equal(opyr.resize_bilinear_u8(img, 6, 8), img)          # identity
    up = opyr.resize_bilinear_u8(img, 11, 15)                               # exact 2x-1: even samples hit pixels
    assert np.array_equal(up[::2, ::2], img)
    mid = (img[0, 0].astype(np.float32) + img[0, 1]) / 2
    assert np.array_equal(up[0, 1], np.floor(mid + 0.5).astype(np.uint8))


def test_pyramid_geometry_is_a_packing():
    for up in (0, 1):
        g = pyramid_geometry(1080, 1920, up)
        assert g.total_level_pixels() == (27110258 if up else 6772594)     # SURVEY.md §8d figures
        occ = np.zeros((g.plane_h, g.plane_w), np.uint8)
        for (x0, y0, w, h) in g.rects:
            assert x0 >= 11 and y0 >= 11 and x0 + w <= g.plane_w - 11 and y0 + h <= g.plane_h - 11
            assert occ[max(0, y0 - 10):y0 + h + 10, max(0, x0 - 10):x0 + w + 10].sum() == 0, "tiles closer than the padding"
            occ[y0:y0 + h, x0:x0 + w] = 1
        assert g.plane_w * g.plane_h < 1.25 * g.total_level_pixels()


def test_decode_nms_known_answer():
    geo = pyramid_geometry(200, 300, 0)
    oh = onets.detector_out_size(geo.plane_h)
    ow = onets.detector_out_size(geo.plane_w)
    s = np.full((oh, ow), -1.0, np.float32)
    s[5, 5] = 2.0
    s[5, 6] = 1.5      # 8 px away: IoU of two 40x40 boxes = 32*40/(2*1600-1280) = 0.67 > 0.4 -> suppressed
    s[5, 20] = 1.0     # 120 px away: kept
    out = opyr.decode(s, geo, 40, 0.0, 0.4, 1.0)
    assert len(out) == 2 and out[0][4] == 2.0 and out[1][4] == 1.0
    px, py = det_cell_to_plane(5, 5)
    l, t, r, b = out[0][:4]
    assert (l, t, r, b) == (px - 20 - 11, py - 20 - 11, px + 19 - 11, py + 19 - 11)  # level 0 at offset (11,11)


# ---------------------------------------------------------------- landmarks / chips
def test_similarity_fit_recovers_transform():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((3, 20, 2)).astype(np.float32)
    th, sc, t = 0.3, 1.7, np.array([5.0, -2.0], np.float32)
    R = sc * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], np.float32)
    b = (a @ R.T + t).astype(np.float32)
    m00, m01, m10, m11, tx, ty = olm.similarity_fit(a, b)
    assert np.allclose([m00[0], m01[0], m10[0], m11[0]], R.reshape(-1), atol=1e-4)
    assert np.allclose([tx[0], ty[0]], t, atol=1e-4)


def test_ert_zero_leaves_returns_scaled_mean_shape_and_single_tree():
    model = W.make_shape_predictor(seed=0, stages=2, trees=3, pool=10)
    model["leaf_values"][:] = 0
    rgb = np.zeros((100, 120, 3), np.uint8)
    rect = np.array([[10, 20, 70, 80]])
    parts = olm.ert_predict(model, rgb, rect)
    init = model["initial_shape"].reshape(-1, 2)
    exp = np.floor(np.stack([10 + init[:, 0] * 60, 20 + init[:, 1] * 60], 1).astype(np.float32) + 0.5)
    assert np.array_equal(parts[0], exp.astype(np.int64))
    # one informative tree: black image -> all features 0 -> diff 0 > thresh only where thresh < 0
    model["split_thresh"][:] = 1.0      # never go left  => node path 0 -> 2 -> 6 -> 14 -> leaf 30-15 = 15
    model["leaf_values"][0, 0, 15, :] = 0.1
    parts2 = olm.ert_predict(model, rgb, rect)
    exp2 = np.floor(np.stack([10 + (init[:, 0] + np.float32(0.1)) * 60, 20 + (init[:, 1] + np.float32(0.1)) * 60], 1).astype(np.float32) + 0.5)
    assert np.array_equal(parts2[0], exp2.astype(np.int64))


def test_chip_of_aligned_face_is_a_crop():
    """landmarks placed exactly at the chip-space targets scaled by 2 and shifted: the chip is a 2x
    sub-sampling of the image."""
    from pyannote_video_b200.ops import chip_from_points
    frm = chip_from_points()
    parts = np.zeros((1, 68, 2), np.int64)
    parts[0] = np.round(frm * 2 + np.array([40, 30]))
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (400, 420, 3), dtype=np.uint8)
    m00, m01, m10, m11, tx, ty = olm.chip_transform(parts)
    assert abs(m00[0] - 2) < 0.02 and abs(m10[0]) < 0.02 and abs(tx[0] - 40) < 1 and abs(ty[0] - 30) < 1
    chips = olm.extract_chips(img, parts)
    assert chips.shape == (1, 150, 150, 3)


# ---------------------------------------------------------------- rectangles / clustering
def test_rect_semantics_and_match():
    a, b = Rect(0, 0, 9, 9), Rect(5, 5, 14, 14)
    assert a.area() == 100 and a.intersect(b).area() == 25 and Rect(3, 3, 2, 2).area() == 0
    da, db = DRect(0, 0, 10, 10), DRect(5, 5, 15, 15)
    assert da.area() == 100.0 and da.intersect(db).area() == 25.0
    assert match_overlap(da, db, 0.2) == 25.0 and match_overlap(da, db, 0.3) == 0.0
    assert match_overlap(da, DRect(20, 20, 30, 30), 0.0) == 0.0


def test_hac_known_cut_and_scipy_crosscheck():
    X = np.zeros((6, 128))
    X[:, 0] = [0.0, 0.1, 0.25, 2.0, 2.2, 5.0]
    lab = ohac.greedy_hac(X, np.arange(6), threshold=0.6)
    assert ohac.partition_of(lab) == {frozenset({0, 1, 2}), frozenset({3, 4}), frozenset({5})}
    # tracks as forced initial groups: {0,3} is one track although its embeddings are far apart
    lab2 = ohac.greedy_hac(X, np.array([0, 1, 2, 0, 4, 5]), threshold=0.6)
    assert lab2[0] == lab2[0] and len(set(lab2.values())) >= 3
    from scipy.cluster.hierarchy import linkage, fcluster
    rng = np.random.default_rng(5)
    cent = rng.standard_normal((5, 128)) * 0.5
    Y = np.concatenate([c + 0.02 * rng.standard_normal((7, 128)) for c in cent])
    ours = ohac.partition_of(ohac.greedy_hac(Y, np.arange(len(Y)), threshold=0.6))
    Z = fcluster(linkage(Y, "average"), t=0.6, criterion="distance")
    ref = ohac.partition_of({i: int(z) for i, z in enumerate(Z)})
    assert ours == ref


# ---------------------------------------------------------------- correlation tracker (DSST)
def test_dsst_oracle_follows_translation_and_zoom():
    """known answers for the tracker restatement: a canvas moving by (-2,-1) px/frame is followed with
    PSR >> 10; a 6 % magnification about the box centre is recovered by the scale filter."""
    import cv2
    from oracle.dsst import CorrelationTracker
    from pyannote_video_b200.synth import make_frames
    fr = make_frames(5, 240, 320, seed=3, shift_per_frame=(2.0, 1.0)).numpy()
    tr = CorrelationTracker()
    tr.start_track(fr[0], (100.0, 60.0, 180.0, 140.0))
    psrs = [tr.update(fr[i]) for i in range(1, 5)]
    p = tr.get_position()
    assert min(psrs) > 10
    assert abs((p[0] - 100.0) - (-2.0 * 4)) < 1.0 and abs((p[1] - 60.0) - (-1.0 * 4)) < 1.0
    z = 1.06
    big = cv2.resize(fr[0], None, fx=z, fy=z, interpolation=cv2.INTER_LINEAR)
    cx, cy = 140.0 * z, 100.0 * z
    ox, oy = int(round(cx - 140.0)), int(round(cy - 100.0))
    f1 = np.ascontiguousarray(big[oy:oy + 240, ox:ox + 320])
    t2 = CorrelationTracker()
    t2.start_track(fr[0], (100.0, 60.0, 180.0, 140.0))
    t2.update(f1)
    q = t2.get_position()
    assert abs((q[2] - q[0]) / 80.0 - z) < 0.02


def test_fhog_known_answers():
    from oracle.dsst import fhog_cell1, fhog_cell4
    flat = np.full((64, 64, 3), 90, np.uint8)
    assert np.abs(fhog_cell1(flat)).max() == 0          # no gradient -> no features
    ramp = np.zeros((64, 64, 3), np.uint8)
    ramp[:] = (np.arange(64) * 3)[None, :, None]         # horizontal ramp: gradient along +x only
    f = fhog_cell1(ramp)
    inner = f[:, 8:56, 8:56]
    on = [c for c in range(18) if inner[c].max() > 0]
    assert on == [0]                                     # orientation bin 0 (cos 0 = 1), contrast-sensitive
    assert inner[18].max() > 0 and inner[19:27].max() == 0
    f4 = fhog_cell4(ramp[:23, :23])
    assert f4.shape == (31, 4, 4) and f4[0].min() > 0 and f4[1:18].max() == 0


def test_rect_overlap_c_entry_point_matches_the_python_rule():
    """pv_rect_overlap (host function of the C ABI) == geometry.match_overlap == TrackingByDetection._match
    (pyannote/video/tracking.py:129-134) on random and degenerate rectangles"""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "pyannote_video_b200", "libpvb200.so"))
    lib.pv_rect_overlap.restype = ctypes.c_double
    rng = np.random.default_rng(0)
    cases = [((0, 0, 10, 10), (5, 5, 15, 15)), ((0, 0, 10, 10), (10, 0, 20, 10)), ((0, 0, 10, 10), (2, 2, 8, 8)),
             ((0, 0, 4, 4), (0, 0, 4, 4))]
    for _ in range(200):
        a = rng.uniform(0, 100, 2)
        b = rng.uniform(0, 100, 2)
        cases.append(((a[0], a[1], a[0] + rng.uniform(1, 60), a[1] + rng.uniform(1, 60)),
                      (b[0], b[1], b[0] + rng.uniform(1, 60), b[1] + rng.uniform(1, 60))))
    for ra, rb in cases:
        for ratio in (0.3, 0.5):
            A = (ctypes.c_double * 4)(*ra)
            B = (ctypes.c_double * 4)(*rb)
            got = lib.pv_rect_overlap(A, B, ctypes.c_double(ratio))
            ref = match_overlap(DRect(*ra), DRect(*rb), ratio)
            assert abs(got - ref) < 1e-9, (ra, rb, ratio, got, ref)
