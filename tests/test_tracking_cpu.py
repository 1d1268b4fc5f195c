"""CPU: our TrackingByDetection control loop against golden tracks produced by the REFERENCE's own
pyannote/video/tracking.py (tests/golden/make_tracking_golden.py) on the same scripted scenarios —
identical tracks, boxes, statuses and yield order."""
import json
import os

import numpy as np
import pytest

import tracking_scenario as sc
from pyannote_video_b200.geometry import DRect
from pyannote_video_b200.hungarian import hungarian
from pyannote_video_b200.tracking import TrackingByDetection, PerObjectBank, get_segment_generator

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "tracking_golden.json")))


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_tracks_equal_reference_golden(idx):
    c = GOLD[idx]["case"]
    video, segs = sc.make_scenario(c["seed"])
    bank = PerObjectBank(sc.make_fake_tracker_class(DRect))
    tracking = TrackingByDetection(sc.make_detect_func(c["seed"]), detect_smallest=36, detect_min_size=c["min_size"],
                                   detect_every=c["every"], track_min_confidence=c["min_conf"],
                                   track_min_overlap_ratio=c["overlap"], track_max_gap=c["gap"], tracker_bank=bank)
    got = [[[t, list(box), status] for t, box, status in track] for track in tracking(video, segs)]
    ref = GOLD[idx]["tracks"]
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g == r
    assert video.frame_size == video.size           # restored (or untouched) like the reference


def test_hungarian_matches_scipy():
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    for n in (1, 2, 4, 7, 11):
        for _ in range(20):
            c = rng.random((n, n))
            pairs = hungarian(c.tolist())
            r, cc = linear_sum_assignment(c)
            assert pairs == list(zip(r.tolist(), cc.tolist()))


def test_hungarian_with_ties_is_optimal_and_deterministic():
    """cost matrices with ties (equal overlaps, the zero padding of `_associate`).  The real `munkres`
    package is absent, so its tie ORDER cannot be pinned; what can be: the assignment is a permutation, its total cost is
    the optimum (scipy), it is the same on every call, and the pairs `_associate` keeps (overlap > 0 inside the real
    block, tracking.py:176-178) have the same total overlap whichever optimal assignment is taken."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(1)
    for n in (2, 3, 5, 8):
        for _ in range(25):
            c = rng.integers(0, 3, size=(n, n)).astype(float)       # many ties
            k = int(rng.integers(0, n))
            c[k:, :] = c.max()                                       # zero-overlap padding rows (cost = max - 0)
            pairs = hungarian(c.tolist())
            assert pairs == hungarian(c.tolist())
            assert sorted(r for r, _ in pairs) == list(range(n)) and sorted(q for _, q in pairs) == list(range(n))
            r, cc = linear_sum_assignment(c)
            assert sum(c[i, j] for i, j in pairs) == c[r, cc].sum()


def test_segment_generator_protocol():
    g = get_segment_generator([sc.Seg(0, 1.0), sc.Seg(1.0, 2.0)])
    g.send(None)
    assert g.send(0.5) is None
    assert g.send(1.0) == 1.0        # first frame with t >= end
    assert g.send(1.04) is None
    assert g.send(2.5) == 2.0
    with pytest.raises(StopIteration):
        g.send(2.6)
