"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's face clustering:
  * `_Model.compute_similarity_matrix` — -squareform(pdist(X, 'euclidean')) then, for every pair of
    clusters, the mean of the block (pyannote/video/face/clustering.py:92-114)
  * `compute_merged_model` / `compute_similarity` — merged cluster = concatenated index sets,
    similarity = mean of the precomputed block (clustering.py:89-90,116-119)
  * pyannote.algorithms HierarchicalAgglomerativeClustering with DistanceThreshold(threshold,
    force=False): merge the most similar pair until the best average distance exceeds the threshold
    (clustering.py:138-148; SURVEY.md App. A.7 — pyannote.algorithms 0.8 is absent, parity unpinned;
    scipy's pdist IS present and is used here exactly as the reference uses it).
"""
import numpy as np
from scipy.spatial.distance import pdist, squareform


def greedy_hac(X, track_of_row, threshold=0.6, metric="euclidean", strict=False):
    """X float64 [N,128]; track_of_row int [N].  Returns dict track_id -> cluster label, where the
    label is the smallest track id of the cluster.  One-pair-at-a-time greedy loop (O(T^3): small T only)."""
    X = np.asarray(X, np.float64)
    track_of_row = np.asarray(track_of_row)
    D = squareform(pdist(X, metric=metric))
    tracks = sorted(set(track_of_row.tolist()))
    members = {t: np.where(track_of_row == t)[0] for t in tracks}
    clusters = {t: [t] for t in tracks}
    def dist(a, b):
        return float(np.mean(D[members[a]][:, members[b]]))
    while len(members) > 1:
        keys = sorted(members)
        best = None
        for i, a in enumerate(keys):
            for b in keys[i + 1:]:
                d = dist(a, b)
                if best is None or d < best[0]:
                    best = (d, a, b)
        d, a, b = best
        if (d >= threshold) if strict else (d > threshold):
            break
        members[a] = np.hstack([members[a], members[b]])
        clusters[a] = clusters[a] + clusters[b]
        del members[b], clusters[b]
    out = {}
    for a, ts in clusters.items():
        lab = min(ts)
        for t in ts:
            out[t] = lab
    return out


def partition_of(labels):
    """dict item -> label  ->  canonical set of frozensets (comparison up to label permutation)"""
    groups = {}
    for k, v in labels.items():
        groups.setdefault(v, set()).add(k)
    return set(frozenset(g) for g in groups.values())
