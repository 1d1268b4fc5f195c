"""Face clustering — drop-in for `pyannote.video.face.clustering.FaceClustering`
(pyannote/video/face/clustering.py:49-148) running on the sm_100a kernels of csrc/cluster.cu.

Reference semantics kept: embeddings are read from `embedding.txt` (`time track d0..d127`), sorted
by (track, time); tracks observed at a single timestamp are not clustered (clustering.py:78-79);
initial clusters are tracks; linkage = mean of all pairwise EUCLIDEAN embedding distances between
two clusters; merging stops once the closest pair is farther than `threshold` (default 0.6).
`metric='cosine'` is the north_star's variant of the same kernels.  pyannote.core is not available
here, so `__call__` returns a plain dict {track: cluster_label} (label = smallest track id in the
cluster) instead of an `Annotation`; `FaceClustering.annotation(...)` renders the same information
as (segment, track, label) triples.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def cluster(emb, track_id, threshold=0.6, metric="euclidean", strict=False, device=None, max_rounds=10000,
            return_stats=False):
    """Threshold-stopped average-linkage clustering of tracks.

    emb      float32 [N,128] (tensor on device, or array); track_id int [N]
    returns  (tracks int64 [T] sorted unique track ids, labels int64 [T]) — label = smallest track id
             of the track's cluster.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("pyannote_video_b200.cluster needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device) if device is not None else (emb.device if isinstance(emb, torch.Tensor) and emb.is_cuda
                                                           else torch.device("cuda", torch.cuda.current_device()))
    X = torch.as_tensor(emb, dtype=torch.float32).to(dev).contiguous()
    track_id = np.asarray(track_id.cpu() if isinstance(track_id, torch.Tensor) else track_id).astype(np.int64)
    N = X.shape[0]
    if N == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    order = np.argsort(track_id, kind="stable")
    tracks, start = np.unique(track_id[order], return_index=True)
    T = len(tracks)
    L = _lib.lib()
    st = _lib.stream_ptr()
    m = 0 if metric == "euclidean" else 1
    if metric not in ("euclidean", "cosine"):
        raise ValueError("metric must be 'euclidean' or 'cosine'")
    D = torch.empty(N, N, dtype=torch.float32, device=dev)
    _lib.check(L.pv_pdist(_lib.ptr(X), C.c_int64(N), 128, m, _lib.ptr(D), st), "pv_pdist")
    # members of current clusters, in terms of the rows/cols of the current matrix
    sizes = np.diff(np.append(start, N)).astype(np.int64)            # embeddings per track
    groups = [[int(t)] for t in range(T)]                              # track indices per cluster

    def contract(S, tin, member_lists):
        tout = len(member_lists)
        offs = np.zeros(tout + 1, np.int32)
        offs[1:] = np.cumsum([len(g) for g in member_lists])
        memb = np.concatenate([np.asarray(g, np.int32) for g in member_lists])
        offs_d, memb_d = _i32(offs, dev), _i32(memb, dev)
        R = torch.empty(tout, tin, dtype=torch.float32, device=dev)
        _lib.check(L.pv_pool_rows(_lib.ptr(S), C.c_int64(tin), _lib.ptr(offs_d), _lib.ptr(memb_d), _lib.ptr(R),
                                  C.c_int64(tout), st), "pv_pool_rows")
        S2 = torch.empty(tout, tout, dtype=torch.float32, device=dev)
        _lib.check(L.pv_pool_cols(_lib.ptr(R), C.c_int64(tin), _lib.ptr(offs_d), _lib.ptr(memb_d), _lib.ptr(S2),
                                  C.c_int64(tout), st), "pv_pool_cols")
        return S2

    if T == N:
        S = D if np.array_equal(order, np.arange(N)) else D[torch.from_numpy(order).to(dev)][:, torch.from_numpy(order).to(dev)].contiguous()
    else:
        S = contract(D, N, [order[start[t]:start[t] + sizes[t]].tolist() for t in range(T)])
        del D
    csize = sizes.astype(np.float64)                                   # embeddings per cluster
    rounds = 0
    while len(groups) > 1 and rounds < max_rounds:
        t = len(groups)
        sz = torch.from_numpy(csize.astype(np.float32)).to(dev)
        nn = torch.empty(t, dtype=torch.int32, device=dev)
        nnd = torch.empty(t, dtype=torch.float32, device=dev)
        _lib.check(L.pv_row_argmin(_lib.ptr(S), C.c_int64(t), _lib.ptr(sz), _lib.ptr(nn), _lib.ptr(nnd), st),
                   "pv_row_argmin")
        nn_h, nnd_h = nn.cpu().numpy(), nnd.cpu().numpy()
        idx = np.arange(t)
        ok = (nnd_h < threshold) if strict else (nnd_h <= threshold)
        recip = (nn_h[nn_h] == idx) & ok & (idx < nn_h)
        pairs = idx[recip]
        if len(pairs) == 0:
            break
        merged_into = np.full(t, -1, np.int64)
        merged_into[nn_h[pairs]] = pairs
        new_members, new_groups, new_sizes = [], [], []
        for a in range(t):
            if merged_into[a] >= 0:
                continue
            if recip[a]:
                b = int(nn_h[a])
                new_members.append([a, b])
                new_groups.append(groups[a] + groups[b])
                new_sizes.append(csize[a] + csize[b])
            else:
                new_members.append([a])
                new_groups.append(groups[a])
                new_sizes.append(csize[a])
        S = contract(S, t, new_members)
        groups, csize = new_groups, np.asarray(new_sizes, np.float64)
        rounds += 1
    labels = np.zeros(T, np.int64)
    for g in groups:
        lab = tracks[min(g)]
        for ti in g:
            labels[ti] = lab
    if return_stats:
        return tracks, labels, dict(rounds=rounds, n_clusters=len(groups))
    return tracks, labels


class _Model(object):
    """Average Euclidean distance between face embeddings (reference: clustering.py:49-119)"""

    def preprocess(self, embedding):
        """Read `embedding.txt`; returns (starting_point, data): starting_point maps track ->
        (start, end) for tracks seen at more than one timestamp, data is a dict of arrays
        time/track/X sorted by (track, time)."""
        raw = np.loadtxt(embedding, ndmin=2, dtype=np.float64)
        if raw.shape[1] != 130:
            raise ValueError("embedding file must have 130 columns (time track d0..d127)")
        order = np.lexsort((raw[:, 0], raw[:, 1]))
        raw = raw[order]
        time, track, X = raw[:, 0], raw[:, 1].astype(np.int64), raw[:, 2:]
        starting_point = {}
        for t in np.unique(track):
            ts = time[track == t]
            if ts.max() > ts.min():            # empty Segment (single timestamp) is skipped
                starting_point[int(t)] = (float(ts.min()), float(ts.max()))
        return starting_point, dict(time=time, track=track, X=X)


class FaceClustering(object):
    """Face clustering

    Parameters
    ----------
    threshold : float, optional
        Defaults to 0.6.

    Usage
    -----
    >>> clustering = FaceClustering()
    >>> starting_point, features = clustering.model.preprocess(embedding)
    >>> result = clustering(starting_point, features=features)
    """

    def __init__(self, threshold=0.6, force=False, logger=None, metric="euclidean"):
        if force:
            raise NotImplementedError("force=True is not used by the reference pipeline")
        self.threshold = threshold
        self.metric = metric
        self.logger = logger
        self.model = _Model()

    def __call__(self, starting_point, features=None):
        keep = np.isin(features["track"], np.asarray(sorted(starting_point), np.int64))
        X = features["X"][keep].astype(np.float32)
        tr = features["track"][keep]
        tracks, labels = cluster(X, tr, threshold=self.threshold, metric=self.metric)
        return {int(t): int(l) for t, l in zip(tracks, labels)}

    @staticmethod
    def annotation(starting_point, result):
        return [(starting_point[t], t, result[t]) for t in sorted(result)]
