"""CPU: the oracle of dlib's HOG frontal detector (oracle/hog.py) — weight-free known answers — and the host-side feature
plane layout of pyannote_video_b200/hog.py."""
import numpy as np

from oracle import hog


def test_fhog_vectorised_equals_the_loop_definition():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (45, 70, 3), dtype=np.uint8)
    assert np.array_equal(hog.fhog_hist(img), hog.fhog_hist_reference(img))


def test_fhog_known_answers():
    # a constant image has no gradients: every feature is zero
    flat = np.full((64, 80, 3), 77, np.uint8)
    f = hog.fhog_features(hog.fhog_hist(flat))
    assert f.shape == (6, 8, 31) and not f.any()
    # a horizontal ramp: only the orientation along +x (bin 0) and its unsigned twin (18) carry energy; far from the
    # border all four block norms are equal, so the clipped value is min(h / (2 h), 0.2) = 0.2 per block
    ramp = np.tile((np.arange(96) * 2).astype(np.uint8)[None, :, None], (64, 1, 3))
    f = hog.fhog_features(hog.fhog_hist(ramp))
    mid = f[2:-2, 2:-2]
    assert np.allclose(mid[..., 0], 0.4) and np.allclose(mid[..., 18], 0.4)
    assert not mid[..., 1:18].any() and not mid[..., 19:27].any()
    assert np.allclose(mid[..., 27:], 0.2357 * 0.2, atol=1e-6)


def test_score_map_and_box_mapping():
    rng = np.random.default_rng(1)
    feat = rng.random((7, 9, 31)).astype(np.float32)
    filt = rng.standard_normal((2, 31, 10, 10)).astype(np.float32)
    sc = hog.score_maps(feat, filt)
    assert sc.shape == (2, 7, 9)
    # entry (y, x): the filter's top-left cell sits at data cell (y - 4, x - 4)
    y, x = 5, 2
    want = np.zeros(2)
    for kh in range(10):
        for kw in range(10):
            yy, xx = y - 4 + kh, x - 4 + kw
            if 0 <= yy < 7 and 0 <= xx < 9:
                want += filt[:, :, kh, kw] @ feat[yy, xx]
    assert np.allclose(sc[:, y, x], want, rtol=1e-5, atol=1e-5)
    # fhog_to_image: padded feature cell 4 (the first data cell) is image cell 1 -> pixel 8 + 1 + 4
    assert hog.fhog_to_image(4, 4) == (13, 13)
    l, t, r, b = hog.level_box(9, 9)              # filter centred on padded cell 9
    assert (r - l, b - t) == (56, 56) and l == (9 - 4 + 1 - 4) * 8 + 5
    assert hog.rect_up((10, 10, 66, 66), 0, True) == (5, 5, 33, 33)
    assert hog.rect_up((10, 10, 66, 66), 1, False) == (12, 12, 80, 80)      # 10 * 1.2 + 0.3 = 12.3, 66 * 1.2 + 0.3 = 79.5 -> 80


def test_nms_keeps_the_best_of_overlapping_boxes():
    assert hog.box_overlap((0, 0, 9, 9), (1, 1, 10, 10))            # 81 / 121 > 0.5
    assert not hog.box_overlap((0, 0, 9, 9), (8, 8, 20, 20))


def test_feature_plane_layout():
    from pyannote_video_b200.hog import hog_geometry, GAP
    from pyannote_video_b200.pyrgeom import pyramid_geometry
    geo = pyramid_geometry(270, 480, 1)
    g, used = hog_geometry(geo)
    assert g.n_levels == len(used) > 3
    rows = []
    for k in range(g.n_levels):
        L = g.lv[k]
        w, h = geo.sizes[used[k]]
        assert (L.w, L.h) == (w, h) and min(w, h) >= 80
        assert (L.cx, L.cy) == (int(w / 8.0 + 0.5), int(h / 8.0 + 0.5))
        rows.append((L.fy0, L.fy0 + L.cy - 2))
        assert L.fx0 == GAP and L.fx0 + L.cx - 2 + GAP <= g.FW
    for (a0, a1), (b0, b1) in zip(rows, rows[1:]):
        assert b0 - a1 >= GAP                       # tiles at least one filter apart
    assert rows[-1][1] + GAP <= g.FH
    assert g.total_feat == sum((g.lv[k].cx - 2) * (g.lv[k].cy - 2) for k in range(g.n_levels))


def test_oracle_matches_committed_golden():
    """regression pin of the restatement (tests/golden/make_hog_golden.py)"""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_hog_golden as mk
    g = np.load(os.path.join(here, "golden", "hog_golden.npz"))
    img, filt = mk.inputs()
    feat = hog.fhog_features(hog.fhog_hist(img))
    assert np.array_equal(feat, g["feat"])
    sc = hog.score_maps(feat, filt)
    assert np.allclose(sc, g["scores"], rtol=0, atol=1e-5)
    boxes, scores, which = hog.detect_levels([img], filt, [float(g["thr"])] * filt.shape[0], upsampled=True)
    assert np.array_equal(np.asarray(boxes, np.int32), g["boxes"]) and list(which) == g["which"].tolist()
    assert np.allclose(scores, g["det_scores"], atol=1e-5)
