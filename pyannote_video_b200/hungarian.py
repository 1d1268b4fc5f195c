"""Minimum-cost assignment (Kuhn–Munkres / Hungarian algorithm, O(n^3)) on the host.

Replaces `munkres.Munkres().compute(cost)` used by TrackingByDetection._associate
(pyannote/video/tracking.py:35,121,172); `munkres` is not installed here and
scipy.optimize.linear_sum_assignment may break ties differently, so the potentials method is
implemented directly.  Matrices are tiny (trackers x detections of one frame).
"""


def hungarian(cost):
    """cost: square (or rectangular, rows <= cols after padding) list of lists.  Returns the list of
    (row, col) pairs of a minimum-cost perfect matching of the rows, sorted by row."""
    n = len(cost)
    if n == 0:
        return []
    m = len(cost[0])
    assert n <= m, "pad the matrix so that rows <= cols"
    INF = float("inf")
    u = [0.0] * (n + 1)
    v = [0.0] * (m + 1)
    p = [0] * (m + 1)       # p[j] = row matched to column j (1-based), 0 = free
    way = [0] * (m + 1)
    for i in range(1, n + 1):
        p[0] = i
        j0 = 0
        minv = [INF] * (m + 1)
        used = [False] * (m + 1)
        while True:
            used[j0] = True
            i0 = p[j0]
            delta = INF
            j1 = 0
            row = cost[i0 - 1]
            for j in range(1, m + 1):
                if not used[j]:
                    cur = row[j - 1] - u[i0] - v[j]
                    if cur < minv[j]:
                        minv[j] = cur
                        way[j] = j0
                    if minv[j] < delta:
                        delta = minv[j]
                        j1 = j
            for j in range(m + 1):
                if used[j]:
                    u[p[j]] += delta
                    v[j] -= delta
                else:
                    minv[j] -= delta
            j0 = j1
            if p[j0] == 0:
                break
        while True:
            j1 = way[j0]
            p[j0] = p[j1]
            j0 = j1
            if j0 == 0:
                break
    pairs = [(p[j] - 1, j - 1) for j in range(1, m + 1) if p[j] != 0]
    pairs.sort()
    return pairs


class Munkres(object):
    """munkres.Munkres duck type"""

    def compute(self, cost_matrix):
        cost = [list(map(float, row)) for row in cost_matrix]
        return hungarian(cost)
