"""Generates tests/golden/tracking_golden.json by running the REFERENCE's own control loop
(/root/reference/pyannote/video/tracking.py, imported unmodified from where it lies) on the
deterministic scenarios of tests/tracking_scenario.py.

The reference module imports `dlib`, `munkres` and `networkx`; networkx is installed, the other two
are not, so they are stubbed: dlib.drectangle -> our DRect shim, dlib.correlation_tracker -> the
scripted FakeTracker, munkres.Munkres -> our Hungarian (checked against scipy in the tests; the
optimum is unique for these inputs).  Run in the build container only:

    python tests/golden/make_tracking_golden.py
"""
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pyannote_video_b200.geometry import DRect            # noqa: E402
from pyannote_video_b200.hungarian import Munkres         # noqa: E402
import tracking_scenario as sc                            # noqa: E402

dlib = types.ModuleType("dlib")
dlib.drectangle = DRect
dlib.correlation_tracker = sc.make_fake_tracker_class(DRect)
munkres = types.ModuleType("munkres")
munkres.Munkres = Munkres
sys.modules["dlib"] = dlib
sys.modules["munkres"] = munkres

spec = importlib.util.spec_from_file_location("ref_tracking", "/root/reference/pyannote/video/tracking.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [
    dict(seed=1, every=0.0, min_conf=10.0, overlap=0.3, gap=0.0, min_size=0.0),
    dict(seed=2, every=0.2, min_conf=10.0, overlap=0.5, gap=1.0, min_size=0.0),
    dict(seed=3, every=0.12, min_conf=10.0, overlap=0.3, gap=0.5, min_size=0.2),
    dict(seed=4, every=0.4, min_conf=12.0, overlap=0.5, gap=1.0, min_size=0.0),
]

out = []
for c in CASES:
    video, segs = sc.make_scenario(c["seed"])
    tracking = ref.TrackingByDetection(sc.make_detect_func(c["seed"]), detect_smallest=36,
                                       detect_min_size=c["min_size"], detect_every=c["every"],
                                       track_min_confidence=c["min_conf"], track_min_overlap_ratio=c["overlap"],
                                       track_max_gap=c["gap"])
    tracks = [[[t, list(box), status] for t, box, status in track] for track in tracking(video, segs)]
    out.append(dict(case=c, tracks=tracks))
    print("case", c["seed"], "tracks", len(tracks), "rows", sum(len(t) for t in tracks))

with open(os.path.join(HERE, "tracking_golden.json"), "w") as f:
    json.dump(out, f)
print("written")
