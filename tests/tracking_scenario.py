"""Deterministic synthetic tracking scenario shared by the golden generator (which runs the
REFERENCE's TrackingByDetection control loop) and the tests (which run ours)."""
import math
import random


class Seg(object):
    def __init__(self, start, end):
        self.start, self.end = start, end


class Frame(object):
    def __init__(self, index, truths):
        self.index = index
        self.truths = truths    # {object id: (l,t,r,b) floats} visible objects


class FakeDRect(object):
    pass


class FakeVideo(object):
    def __init__(self, frames, fps=25.0, size=(640, 360)):
        self.frames = frames
        self.frame_rate = fps
        self.size = size
        self._frame_size = size

    @property
    def frame_size(self):
        return self._frame_size

    @frame_size.setter
    def frame_size(self, v):
        self._frame_size = tuple(v)

    def __iter__(self):
        for i, f in enumerate(self.frames):
            yield (i / self.frame_rate, f)


def make_scenario(seed, n_frames=60, n_objects=3, shots=(0.92, 1.64, 2.4)):
    rnd = random.Random(seed)
    objs = []
    for k in range(n_objects):
        objs.append(dict(x=rnd.uniform(50, 450), y=rnd.uniform(40, 200), s=rnd.uniform(40, 90),
                         vx=rnd.uniform(-3, 3), vy=rnd.uniform(-2, 2), a=rnd.randrange(0, 15),
                         b=rnd.randrange(35, n_frames)))
    frames = []
    for i in range(n_frames):
        truths = {}
        for k, o in enumerate(objs):
            if o["a"] <= i <= o["b"]:
                x = o["x"] + o["vx"] * i + 2.0 * math.sin(0.3 * i + k)
                y = o["y"] + o["vy"] * i
                truths[k] = (x, y, x + o["s"], y + o["s"])
        frames.append(Frame(i, truths))
    segs, start = [], 0.0
    for e in shots:
        segs.append(Seg(start, e))
        start = e
    return FakeVideo(frames), segs


def make_detect_func(seed, miss=0.25):
    def detect(frame):
        rnd = random.Random(seed * 1000 + frame.index)
        out = []
        for k in sorted(frame.truths):
            if rnd.random() < miss:
                continue
            l, t, r, b = frame.truths[k]
            out.append((int(l + rnd.uniform(-2, 2)), int(t + rnd.uniform(-2, 2)), int(r + rnd.uniform(-2, 2)),
                        int(b + rnd.uniform(-2, 2))))
        return out
    return detect


def make_fake_tracker_class(drect_class):
    """dlib.correlation_tracker stand-in: follows the visible object that overlaps it most, with a
    lag; confidence (PSR) drops below 10 when nothing is left to follow."""

    class FakeTracker(object):
        def start_track(self, frame, rect):
            self.pos = [rect.left(), rect.top(), rect.right(), rect.bottom()]

        def update(self, frame):
            best, best_ov = None, 0.0
            for k, (l, t, r, b) in sorted(frame.truths.items()):
                w = min(self.pos[2], r) - max(self.pos[0], l)
                h = min(self.pos[3], b) - max(self.pos[1], t)
                ov = max(w, 0.0) * max(h, 0.0)
                if ov > best_ov:
                    best, best_ov = (l, t, r, b), ov
            if best is None:
                return 4.0
            self.pos = [0.5 * p + 0.5 * q + 0.125 for p, q in zip(self.pos, best)]
            area = (self.pos[2] - self.pos[0]) * (self.pos[3] - self.pos[1])
            return 8.0 + 30.0 * best_ov / max(area, 1.0)

        def get_position(self):
            return drect_class(*self.pos)

    return FakeTracker
