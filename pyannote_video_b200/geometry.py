"""dlib-object duck types used by the reference pipeline (host side, plain Python).

The reference touches dlib objects directly:
  rectangle / drectangle : .left() .top() .right() .bottom() .area() .intersect()
                           (pyannote/video/face/tracking.py:41, pyannote/video/tracking.py:130-134,233-236)
  full_object_detection  : .parts() -> points with .x .y, .rect, .part(i), .num_parts
                           (pyannote/video/face/face.py:81-82, scripts/pyannote-face.py:301-302)
Area conventions [MEMORY of dlib 19.12, isolated here]: `rectangle` is inclusive
(width = right-left+1); `drectangle` is continuous (width = right-left).
"""


class Rect(object):
    """dlib.rectangle: integer, inclusive right/bottom."""
    __slots__ = ("l", "t", "r", "b")

    def __init__(self, left=0, top=0, right=-1, bottom=-1):
        self.l, self.t, self.r, self.b = int(left), int(top), int(right), int(bottom)

    def left(self):
        return self.l

    def top(self):
        return self.t

    def right(self):
        return self.r

    def bottom(self):
        return self.b

    def is_empty(self):
        return self.t > self.b or self.l > self.r

    def width(self):
        return 0 if self.is_empty() else self.r - self.l + 1

    def height(self):
        return 0 if self.is_empty() else self.b - self.t + 1

    def area(self):
        return self.width() * self.height()

    def intersect(self, other):
        return Rect(max(self.l, other.l), max(self.t, other.t), min(self.r, other.r), min(self.b, other.b))

    def contains(self, x, y):
        return self.l <= x <= self.r and self.t <= y <= self.b

    def __eq__(self, other):
        return isinstance(other, Rect) and (self.l, self.t, self.r, self.b) == (other.l, other.t, other.r, other.b)

    def __hash__(self):
        return hash((self.l, self.t, self.r, self.b))

    def __repr__(self):
        return "[(%d, %d) (%d, %d)]" % (self.l, self.t, self.r, self.b)


class DRect(object):
    """dlib.drectangle: floating point, continuous (area = (r-l)*(b-t), empty if r<l or b<t)."""
    __slots__ = ("l", "t", "r", "b")

    def __init__(self, left=0.0, top=0.0, right=-1.0, bottom=-1.0):
        self.l, self.t, self.r, self.b = float(left), float(top), float(right), float(bottom)

    def left(self):
        return self.l

    def top(self):
        return self.t

    def right(self):
        return self.r

    def bottom(self):
        return self.b

    def is_empty(self):
        return self.t > self.b or self.l > self.r

    def width(self):
        return 0.0 if self.is_empty() else self.r - self.l

    def height(self):
        return 0.0 if self.is_empty() else self.b - self.t

    def area(self):
        return self.width() * self.height()

    def intersect(self, other):
        return DRect(max(self.l, other.l), max(self.t, other.t), min(self.r, other.r), min(self.b, other.b))

    def __repr__(self):
        return "[(%g, %g) (%g, %g)]" % (self.l, self.t, self.r, self.b)


class Point(object):
    __slots__ = ("x", "y")

    def __init__(self, x, y):
        self.x, self.y = int(x), int(y)

    def __repr__(self):
        return "(%d, %d)" % (self.x, self.y)


class FullObjectDetection(object):
    """dlib.full_object_detection"""

    def __init__(self, rect, parts):
        self.rect = rect
        self._parts = [p if isinstance(p, Point) else Point(p[0], p[1]) for p in parts]

    @property
    def num_parts(self):
        return len(self._parts)

    def part(self, i):
        return self._parts[i]

    def parts(self):
        return list(self._parts)


def match_overlap(r1, r2, min_overlap_ratio):
    """TrackingByDetection._match (pyannote/video/tracking.py:129-134): intersection area, zeroed
    unless it covers at least `ratio` of BOTH rectangles."""
    overlap = r1.intersect(r2).area()
    if (overlap < min_overlap_ratio * r1.area()) or (overlap < min_overlap_ratio * r2.area()):
        overlap = 0.
    return overlap
