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


@pytest.mark.parametrize("control", ["python", "native"])
@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_tracks_equal_reference_golden(idx, control):
    """`control="native"`: association + link graph + _fix + _fill_gaps in C++ (csrc/control.cu, host code of libpvb200.so)"""
    c = GOLD[idx]["case"]
    video, segs = sc.make_scenario(c["seed"])
    bank = PerObjectBank(sc.make_fake_tracker_class(DRect))
    tracking = TrackingByDetection(sc.make_detect_func(c["seed"]), detect_smallest=36, detect_min_size=c["min_size"],
                                   detect_every=c["every"], track_min_confidence=c["min_conf"],
                                   track_min_overlap_ratio=c["overlap"], track_max_gap=c["gap"], tracker_bank=bank,
                                   control=control)
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


def test_native_control_equals_python_on_random_scenarios():
    """more objects, crowded scenes (ties and conflicts in the association), gaps, every-N detection: the C++ control
    path must reproduce the Python one track for track (times, integer boxes, status strings, yield order)"""
    import random
    rnd = random.Random(7)
    for seed in range(100, 140):
        video, segs = sc.make_scenario(seed, n_frames=rnd.choice([50, 70, 90]), n_objects=rnd.choice([2, 4, 6, 9]),
                                       shots=rnd.choice([(0.6, 1.3, 3.7), (3.7, ), (0.4, 0.9, 1.5, 3.7)]))
        kw = dict(detect_smallest=36, detect_min_size=0.0, detect_every=rnd.choice([0.0, 0.08, 0.2]),
                  track_min_confidence=rnd.choice([5.0, 10.0]), track_min_overlap_ratio=rnd.choice([0.1, 0.3, 0.6]),
                  track_max_gap=rnd.choice([0.0, 0.2, 1.0]))
        miss = rnd.choice([0.1, 0.4])
        out = {}
        for control in ("python", "native"):
            bank = PerObjectBank(sc.make_fake_tracker_class(DRect))
            tr = TrackingByDetection(sc.make_detect_func(seed, miss=miss), tracker_bank=bank, control=control, **kw)
            out[control] = [[(t, tuple(box), status) for t, box, status in track] for track in tr(video, segs)]
        assert out["python"] == out["native"], (seed, kw)
        assert len(out["python"]) > 0


def test_native_associate_matches_python_with_ties():
    import ctypes as C
    from pyannote_video_b200 import _lib
    rng = np.random.default_rng(3)
    for _ in range(200):
        nt, nd = int(rng.integers(1, 7)), int(rng.integers(1, 7))
        # boxes on a coarse grid: many equal overlaps
        pos = [DRect(*(np.array([x, y, x + 40, y + 40], float))) for x, y in rng.integers(0, 4, size=(nt, 2)) * 20]
        det = [tuple(int(v) for v in (x, y, x + 40, y + 40)) for x, y in rng.integers(0, 4, size=(nd, 2)) * 20]
        a = TrackingByDetection(lambda f: [], control="python", track_min_overlap_ratio=0.3)._associate(pos, det)
        b = TrackingByDetection(lambda f: [], control="native", track_min_overlap_ratio=0.3)._associate(pos, det)
        assert a == b and list(a) == list(b)


def test_segment_generator_protocol():
    g = get_segment_generator([sc.Seg(0, 1.0), sc.Seg(1.0, 2.0)])
    g.send(None)
    assert g.send(0.5) is None
    assert g.send(1.0) == 1.0        # first frame with t >= end
    assert g.send(1.04) is None
    assert g.send(2.5) == 2.0
    with pytest.raises(StopIteration):
        g.send(2.6)
