"""Tracking by detection — drop-in for `pyannote.video.tracking.TrackingByDetection`
(pyannote/video/tracking.py:68-434) and `pyannote.video.face.tracking.FaceTracking`
(pyannote/video/face/tracking.py:36-78).

Same constructor arguments, same call protocol (`tracking(video, segmentation)` is a generator of
tracks `[(t, (left, top, right, bottom) / frame size, status), ...]`), same semantics, re-built
around a *bank* of correlation trackers so that all live trackers of a frame are advanced by ONE
batched kernel launch instead of one dlib call each:

  per frame (both passes):  update every tracker -> drop those below `track_min_confidence`
  -> Hungarian association with the frame's detections (overlap must cover `track_min_overlap_ratio`
  of both boxes) -> matched trackers are linked to the detection and retired -> unmatched trackers
  are linked to their new position -> a fresh tracker starts on EVERY detection.

Tracks are the connected components of the link graph after a forward and a backward pass over
the shot; positions seen at the same time are averaged (`_fix`) and tracks separated by less than
`track_max_gap` whose end/start boxes overlap are concatenated (`_fill_gaps`).  The graph is a
union-find over insertion-ordered nodes (networkx is not needed); component order follows node
insertion order like networkx's, so tie-breaking between simultaneous tracks is unchanged.
"""
from __future__ import division

import itertools

import numpy as np

from .geometry import DRect, match_overlap
from .hungarian import Munkres

FORWARD = 'forward'
BACKWARD = 'backward'
DETECTION = 'detection'
ERROR = 'error'

_STATUS_ORDER = {FORWARD: 1, DETECTION: 2, BACKWARD: 3}


def get_segment_generator(segmentation):
    """Time-driven segment generator: send(t) returns a segment's end once, on the first t >= end
    (reference: pyannote/video/tracking.py:44-58)."""
    t = yield
    for segment in segmentation:
        end = segment.end
        while end > t:
            t = yield
        t = yield end


def get_min_max_t(track):
    """Get track start and end times"""
    times = [t for t, _, _ in track]
    return (min(times), max(times))


class PerObjectBank(object):
    """Adapts a dlib-style tracker class (start_track / update / get_position) to the bank
    interface; used with third-party trackers and in the CPU tests."""

    def __init__(self, tracker_class):
        self.tracker_class = tracker_class
        self.trackers = {}
        self._next = 0

    def reset(self):
        self.trackers = {}

    def start(self, frame, rect):
        tr = self.tracker_class()
        tr.start_track(frame, rect)
        self._next += 1
        self.trackers[self._next] = tr
        return self._next

    def update(self, frame, handles):
        return [self.trackers[h].update(frame) for h in handles]

    def position(self, handle):
        return self.trackers[handle].get_position()

    def release(self, handle):
        del self.trackers[handle]


class _LinkGraph(object):
    """insertion-ordered nodes + union-find"""

    def __init__(self):
        self.index = {}
        self.nodes = []
        self.parent = []

    def add(self, node):
        i = self.index.get(node)
        if i is None:
            i = len(self.nodes)
            self.index[node] = i
            self.nodes.append(node)
            self.parent.append(i)
        return i

    def find(self, i):
        root = i
        while self.parent[root] != root:
            root = self.parent[root]
        while self.parent[i] != root:
            self.parent[i], i = root, self.parent[i]
        return root

    def link(self, a, b):
        ra, rb = self.find(self.add(a)), self.find(self.add(b))
        if ra != rb:
            self.parent[max(ra, rb)] = min(ra, rb)

    def components(self):
        groups = {}
        for i, node in enumerate(self.nodes):
            groups.setdefault(self.find(i), []).append(node)
        return [groups[r] for r in sorted(groups)]   # root = smallest index = first inserted node


_STATUS_CODE = {FORWARD: 0, DETECTION: 1, BACKWARD: 2}


class _NativeGraph(object):
    """the shot's link graph in C++ (csrc/control.cu): same add / link protocol as _LinkGraph; `tracks()` runs the
    connected components, _fix, _fill_gaps and the final sort natively"""

    def __init__(self):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib
        self.h = C.c_void_p()
        _lib.check(_lib.lib().pv_ctl_shot_create(C.byref(self.h)), "pv_ctl_shot_create")

    def _box(self, box):
        return (self._C.c_double * 4)(*[float(v) for v in box])

    def add(self, node):
        t, box, status = node
        self._lib.check(self._lib.lib().pv_ctl_shot_add(self.h, self._C.c_double(t), self._box(box), _STATUS_CODE[status]),
                        "pv_ctl_shot_add")

    def link(self, a, b):
        C = self._C
        self._lib.check(self._lib.lib().pv_ctl_shot_link(self.h, C.c_double(a[0]), self._box(a[1]), _STATUS_CODE[a[2]],
                                                         C.c_double(b[0]), self._box(b[1]), _STATUS_CODE[b[2]]),
                        "pv_ctl_shot_link")

    def tracks(self, min_overlap_ratio, max_gap):
        C, L = self._C, self._lib.lib()
        nt, nr = C.c_int(), C.c_int()
        self._lib.check(L.pv_ctl_shot_finish(self.h, C.c_double(min_overlap_ratio), C.c_double(max_gap), C.byref(nt), C.byref(nr)),
                        "pv_ctl_shot_finish")
        lens = (C.c_int * max(nt.value, 1))()
        ts = (C.c_double * max(nr.value, 1))()
        boxes = (C.c_longlong * (4 * max(nr.value, 1)))()
        counts = (C.c_int * (4 * max(nr.value, 1)))()
        self._lib.check(L.pv_ctl_shot_tracks(self.h, lens, ts, boxes, counts), "pv_ctl_shot_tracks")
        out, r = [], 0
        for k in range(nt.value):
            track = []
            for _ in range(lens[k]):
                nf, nd, nb, err = counts[4 * r:4 * r + 4]
                status = "+".join([FORWARD] * nf + [DETECTION] * nd + [BACKWARD] * nb)
                if err:
                    status = "error({0})".format(status)
                track.append((ts[r], (boxes[4 * r], boxes[4 * r + 1], boxes[4 * r + 2], boxes[4 * r + 3]), status))
                r += 1
            out.append(track)
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._lib.lib().pv_ctl_shot_destroy(self.h)
                self.h = None
        except Exception:
            pass


class TrackingByDetection(object):
    """(Forward/backward) tracking by detection

    Parameters
    ----------
    detect_func : func
        Detection function. Takes a video frame and returns an iterable of detections as
        (left, top, right, bottom) tuples.
    detect_smallest : int, optional
        Smallest object (height, in pixels) that `detect_func` can detect.
    detect_min_size : float, optional
        Approximate size (in video height ratio) of the smallest object that should be detected.
    detect_every : float, optional
        When provided, `detect_func` is applied every `detect_every` seconds.
    track_min_confidence : float, optional
        Kill trackers whose confidence goes below this value. Defaults to 10.
    track_min_overlap_ratio : float, optional
        Do not associate trackers and detections if their overlap ratio goes below this value.
    track_max_gap : float, optional
        Bridge gaps with duration shorter than this value.
    tracker_bank : object, optional
        Correlation-tracker bank (start/update/position/release/reset).  Defaults to the CUDA
        `TrackerBank` of this package.
    control : "python" | "native", optional
        Where the per-frame association (overlap matrix + Munkres) and the per-shot link graph (components, `_fix`,
        `_fill_gaps`) run: in this file, or in C++ (`csrc/control.cu`, `pv_ctl_*`).  Both give identical tracks
        (tests/test_tracking_cpu.py: the reference's golden scenarios and random ones); default: `PV_TRACK_CONTROL`
        or "python".
    """

    def __init__(self, detect_func, detect_smallest=1, detect_min_size=0., detect_every=0.,
                 track_min_confidence=10., track_min_overlap_ratio=0.3, track_max_gap=0.,
                 tracker_bank=None, prepare_frame=None, control=None):
        super(TrackingByDetection, self).__init__()
        self.detect_func = detect_func
        self.detect_smallest = detect_smallest
        self.detect_min_size = detect_min_size
        self.detect_every = detect_every
        self.track_min_confidence = track_min_confidence
        self.track_min_overlap_ratio = track_min_overlap_ratio
        self.track_max_gap = track_max_gap
        self._hungarian = Munkres()
        self._bank = tracker_bank
        self._prepare_frame = prepare_frame
        import os
        self.control = control or os.environ.get("PV_TRACK_CONTROL", "python")
        if self.control not in ("python", "native"):
            raise ValueError("control must be 'python' or 'native'")

    # ------------------------------------------------------------------ helpers
    def _get_bank(self):
        """resolve the tracker bank and the frame-upload hook (called at the top of __call__, before the first
        frame is cached, so that the whole shot — not only its tail — is cached as device tensors)"""
        if self._bank is None:
            from .tracker import TrackerBank
            self._bank = TrackerBank()
        if self._prepare_frame is None:
            self._prepare_frame = getattr(self._bank, "prepare_frame", None)
        return self._bank

    def _match(self, rectangle1, rectangle2):
        return match_overlap(rectangle1, rectangle2, self.track_min_overlap_ratio)

    def _associate(self, positions, detections):
        """positions: list of DRect (live trackers, in tracker order); detections: list of boxes.
        Returns {detection index: tracker index}."""
        n_trackers, n_detections = len(positions), len(detections)
        if n_trackers < 1 or n_detections < 1:
            return dict()
        if self.control == "native":
            import ctypes as C
            from . import _lib
            pos = (C.c_double * (4 * n_trackers))(*[v for p in positions for v in (p.left(), p.top(), p.right(), p.bottom())])
            det = (C.c_double * (4 * n_detections))(*[float(v) for d in detections for v in d])
            out = (C.c_int * n_detections)()
            _lib.check(_lib.lib().pv_ctl_associate(pos, n_trackers, det, n_detections, C.c_double(self.track_min_overlap_ratio), out),
                       "pv_ctl_associate")
            # same insertion order as the Python path (pairs sorted by tracker index)
            return {d: t for t, d in sorted((out[d], d) for d in range(n_detections) if out[d] >= 0)}
        n = max(n_trackers, n_detections)
        overlap_area = np.zeros((n, n))
        for t, position in enumerate(positions):
            for d, detection in enumerate(detections):
                overlap_area[t, d] = self._match(position, DRect(*detection))
        match = {}
        for t, d in self._hungarian.compute(np.max(overlap_area) - overlap_area):
            if t >= n_trackers or d >= n_detections:
                continue
            if overlap_area[t, d] > 0.:
                match[d] = t
        return match

    # ------------------------------------------------------------------ one pass over the shot
    def _track(self, direction=FORWARD):
        if direction == FORWARD:
            frame_cache = self._frame_cache
        elif direction == BACKWARD:
            frame_cache = reversed(self._frame_cache)
        else:
            raise NotImplementedError()
        bank = self._get_bank()
        bank.reset()
        live = []          # [handle, previous node, confidence] in tracker-creation order
        graph = self._graph
        for t, frame in frame_cache:
            # update trackers & end those with low confidence (one batched launch)
            if live:
                confidences = bank.update(frame, [h for h, _, _ in live])
                survivors = []
                for entry, confidence in zip(live, confidences):
                    entry[2] = confidence
                    if confidence < self.track_min_confidence:
                        bank.release(entry[0])
                    else:
                        survivors.append(entry)
                live = survivors
            detections = self._detections.get(t, [])
            positions = [bank.position(h) for h, _, _ in live]
            match = self._associate(positions, detections)
            # matched trackers: link previous position to the detection, retire the tracker
            matched = set()
            for d, ti in match.items():
                graph.link(live[ti][1], (t, detections[d], DETECTION))
                matched.add(ti)
            for ti in matched:
                bank.release(live[ti][0])
            # unmatched trackers: link previous position to the current one
            remaining = []
            for ti, entry in enumerate(live):
                if ti in matched:
                    continue
                p = positions[ti]
                current = (t, (p.left(), p.top(), p.right(), p.bottom()), direction)
                graph.link(entry[1], current)
                entry[1] = current
                remaining.append(entry)
            live = remaining
            # a new tracker starts on every detection
            for detection in detections:
                handle = bank.start(frame, DRect(*detection))
                live.append([handle, (t, detection, DETECTION), None])
        for h, _, _ in live:
            bank.release(h)

    def _fix(self, track):
        """merge forward/backward/detection positions seen at the same time"""
        fixed_track = []
        for t, group in itertools.groupby(sorted(track), key=lambda x: x[0]):
            group = list(group)
            error = False
            for (_, pos1, _), (_, pos2, _) in itertools.combinations(group, 2):
                if self._match(DRect(*pos1), DRect(*pos2)) == 0:
                    error = True
                    break
            status = "+".join(sorted((s for _, _, s in group), key=lambda s: _STATUS_ORDER[s]))
            if error:
                status = "error({0})".format(status)
            pos = tuple(int(round(v)) for v in np.mean(np.vstack([p for _, p, _ in group]), axis=0))
            fixed_track.append((t, pos, status))
        return fixed_track

    def _fill_gaps(self, tracks):
        tracks = sorted(tracks, key=get_min_max_t)
        n = len(tracks)
        parent = list(range(n))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i

        for i, j in itertools.combinations(range(n), 2):
            ti = tracks[i][-1][0]
            tj = tracks[j][0][0]
            if (tj < ti) or (tj - ti > self.track_max_gap):
                continue
            if self._match(DRect(*tracks[i][-1][1]), DRect(*tracks[j][0][1])):
                ri, rj = find(i), find(j)
                if ri != rj:
                    parent[max(ri, rj)] = min(ri, rj)
        groups = {}
        for i in range(n):
            groups.setdefault(find(i), []).append(i)
        return [[item for k in sorted(groups[r]) for item in tracks[k]] for r in sorted(groups)]

    def _forward_backward(self):
        self._track(direction=FORWARD)
        self._track(direction=BACKWARD)
        if self.control == "native":
            for track in self._graph.tracks(self.track_min_overlap_ratio, self.track_max_gap):
                yield track
            return
        tracks = [self._fix(track) for track in self._graph.components()]
        tracks = self._fill_gaps(tracks)
        for track in sorted(tracks, key=get_min_max_t):
            yield track

    def _reset(self):
        self._frame_cache = []
        self._graph = _NativeGraph() if self.control == "native" else _LinkGraph()
        self._detections = {}

    def _normalize_track(self, track, frame_width, frame_height):
        return [(t, (left / frame_width, top / frame_height, right / frame_width, bottom / frame_height), status)
                for (t, (left, top, right, bottom), status) in track]

    def __call__(self, video, segmentation):
        """
        Parameters
        ----------
        video : Video-like (iteration yields (t, frame); `.frame_rate`, `.size`, `.frame_size`)
        segmentation : iterable of segments with `.end` (shots)
        """
        if self.detect_every > 0.0:
            every_x_frames = int(self.detect_every * video.frame_rate)
        else:
            every_x_frames = 1

        width, height = video.size
        ratio = 1.0
        if self.detect_min_size > 0.0:
            ratio = min(1.0, self.detect_smallest / (self.detect_min_size * height))

        old_frame_width, old_frame_height = video.frame_size
        frame_width = int(width * ratio)
        frame_height = int(height * ratio)
        video.frame_size = (frame_width, frame_height)

        segment_generator = get_segment_generator(segmentation)
        segment_generator.send(None)
        self._get_bank()
        self._reset()

        for i, (t, frame) in enumerate(video):
            if segment_generator.send(t):
                for track in self._forward_backward():
                    yield self._normalize_track(track, frame_width, frame_height)
                self._reset()
            if self._prepare_frame is not None:
                frame = self._prepare_frame(frame)
            self._frame_cache.append((t, frame))
            if i % every_x_frames == 0:
                for detection in self.detect_func(frame):
                    detection = tuple(detection)
                    self._graph.add((t, detection, DETECTION))
                    dets = self._detections.setdefault(t, [])
                    if detection not in dets:
                        dets.append(detection)

        for track in self._forward_backward():
            yield self._normalize_track(track, frame_width, frame_height)

        if self.detect_min_size > 0.0:
            video.frame_size = (old_frame_width, old_frame_height)


def get_face_detect(face):
    """Create function for face detection"""
    def face_detect(frame):
        for f in face.iterfaces(frame):
            yield (f.left(), f.top(), f.right(), f.bottom())
    return face_detect


class FaceTracking(TrackingByDetection):
    """Face tracking (same parameters as pyannote/video/face/tracking.py:45-78)"""

    def __init__(self, detect_min_size=0., detect_every=0., track_min_confidence=10., track_min_overlap_ratio=0.3,
                 track_max_gap=0., face=None, tracker_bank=None, control=None):
        from .face import Face, DLIB_SMALLEST_FACE
        face = face if face is not None else Face()
        super(FaceTracking, self).__init__(
            detect_func=get_face_detect(face), detect_smallest=DLIB_SMALLEST_FACE, detect_min_size=detect_min_size,
            detect_every=detect_every, track_min_confidence=track_min_confidence,
            track_min_overlap_ratio=track_min_overlap_ratio, track_max_gap=track_max_gap, tracker_bank=tracker_bank,
            control=control)
