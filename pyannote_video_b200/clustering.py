"""Face clustering — drop-in for `pyannote.video.face.clustering.FaceClustering`
(pyannote/video/face/clustering.py:49-148) running on the sm_100a kernels of csrc/cluster.cu.

Reference semantics kept: embeddings are read from `embedding.txt` (`time track d0..d127`), sorted
by (track, time); tracks observed at a single timestamp are not clustered (clustering.py:78-79);
initial clusters are tracks; linkage = mean of all pairwise EUCLIDEAN embedding distances between
two clusters; merging stops once the closest pair is farther than `threshold` (default 0.6).
`metric='cosine'` is the north_star's variant of the same kernels.  `preprocess` returns the starting
point as an `Annotation` (one segment per track, track name = label = track id) and `__call__` returns an
`Annotation` with the merged labels (label = smallest track id of the cluster), like the reference;
pyannote.core is absent here, `annotation.py` provides the duck types (`.to_dict()` gives {track: label}).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .annotation import Annotation, Segment


def _i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def pairwise_distances(X, metric="euclidean", impl="tcgen05"):
    """D float32 [N,N] on X's device: scipy pdist+squareform (clustering.py:101).  impl 'tcgen05' = csrc/gram.cu (bf16 x 3
    split Gram on the tensor cores, fused distance epilogue), 'fp32' = the CUDA-core difference kernel (kept as the
    cross-check: it is exact for duplicate points, the Gram form has |dD| <= ~2e-7 / D)."""
    if metric not in ("euclidean", "cosine"):
        raise ValueError("metric must be 'euclidean' or 'cosine'")
    m = 0 if metric == "euclidean" else 1
    L = _lib.lib()
    st = _lib.stream_ptr()
    N = X.shape[0]
    dev = X.device
    D = torch.empty(N, N, dtype=torch.float32, device=dev)
    if impl == "fp32":
        _lib.check(L.pv_pdist(_lib.ptr(X), C.c_int64(N), 128, m, _lib.ptr(D), st), "pv_pdist")
        return D
    L.pv_gram_npad.restype = C.c_int64
    npad = int(L.pv_gram_npad(C.c_int64(N)))
    xs = torch.empty(3 * npad, 128, dtype=torch.bfloat16, device=dev)
    norms = torch.empty(npad, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(L.pv_gram_dist(_lib.ptr(X), C.c_int64(N), 128, m, _lib.ptr(D), _lib.ptr(xs), _lib.ptr(norms), _lib.ptr(err), st),
               "pv_gram_dist")
    return D


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
    D = pairwise_distances(X, metric)
    # members of current clusters, in terms of the rows/cols of the current matrix
    sizes = np.diff(np.append(start, N)).astype(np.int64)            # embeddings per track

    def contract_csr(S, tin, offs, memb):
        tout = len(offs) - 1
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
        # rows are grouped by track through `order`: the CSR of the first contraction is (start, order) as they stand
        S = contract_csr(D, N, np.append(start, N).astype(np.int32), order.astype(np.int32))
        del D
    # ---- agglomeration rounds, entirely on the device: the host only learns how many clusters are left ----
    t = T
    sz = torch.from_numpy(sizes.astype(np.float32)).to(dev)           # embeddings per cluster
    cl = torch.arange(T, dtype=torch.int32, device=dev)               # cluster index of every track
    rounds = 0
    S_store = (S._base if S._base is not None else S).reshape(-1)     # flat storage behind the current matrix
    spare = None
    while t > 1 and rounds < max_rounds:
        nn = torch.empty(t, dtype=torch.int32, device=dev)
        nnd = torch.empty(t, dtype=torch.float32, device=dev)
        _lib.check(L.pv_row_argmin(_lib.ptr(S), C.c_int64(t), _lib.ptr(sz), _lib.ptr(nn), _lib.ptr(nnd), st),
                   "pv_row_argmin")
        keep = torch.empty(t, dtype=torch.int32, device=dev)
        partner = torch.empty(t, dtype=torch.int32, device=dev)
        _lib.check(L.pv_hac_plan(_lib.ptr(nn), _lib.ptr(nnd), C.c_int64(t), C.c_float(threshold), int(bool(strict)),
                                 _lib.ptr(keep), _lib.ptr(partner), st), "pv_hac_plan")
        incl = torch.cumsum(keep, 0, dtype=torch.int32)
        tout = int(incl[-1].item())                                   # the round's only device -> host read
        if tout == t:
            break
        newidx = (incl - keep).contiguous()
        m0 = torch.empty(tout, dtype=torch.int32, device=dev)
        m1 = torch.empty(tout, dtype=torch.int32, device=dev)
        sz2 = torch.empty(tout, dtype=torch.float32, device=dev)
        mp = torch.empty(t, dtype=torch.int32, device=dev)
        _lib.check(L.pv_hac_members(_lib.ptr(keep), _lib.ptr(partner), _lib.ptr(newidx), _lib.ptr(nn), _lib.ptr(sz),
                                    C.c_int64(t), _lib.ptr(m0), _lib.ptr(m1), _lib.ptr(sz2), _lib.ptr(mp), st),
                   "pv_hac_members")
        # the contracted matrix goes into the OTHER of two flat buffers (the distance matrix's storage and one more block
        # of the first contraction's size): no allocation and no free of multi-gigabyte blocks inside the loop
        if spare is None:
            spare = torch.empty(tout * tout, dtype=torch.float32, device=dev)
        S2 = spare[:tout * tout].view(tout, tout)
        _lib.check(L.pv_hac_contract(_lib.ptr(S), C.c_int64(t), _lib.ptr(m0), _lib.ptr(m1), _lib.ptr(S2), C.c_int64(tout), st),
                   "pv_hac_contract")
        _lib.check(L.pv_hac_relabel(_lib.ptr(cl), C.c_int64(T), _lib.ptr(mp), st), "pv_hac_relabel")
        spare, S_store = S_store, spare
        S, sz, t = S2, sz2, tout
        del S2
        rounds += 1
    # label of a cluster = its smallest track id (tracks are sorted: the smallest track INDEX)
    first = torch.full((t, ), T, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, cl.long(), torch.arange(T, device=dev), reduce="amin")
    labels = tracks[first[cl.long()].cpu().numpy()]
    groups = range(t)
    if return_stats:
        return tracks, labels, dict(rounds=rounds, n_clusters=len(groups))
    return tracks, labels


class _Model(object):
    """Average Euclidean distance between face embeddings (reference: clustering.py:49-119)"""

    def preprocess(self, embedding):
        """Read `embedding.txt` (`time track d0..d127`); returns (starting_point, data): starting_point is an
        Annotation with one segment [first, last timestamp] per track (tracks seen at a single timestamp give an
        empty segment and are dropped, clustering.py:76-80), data a dict of arrays time/track/X sorted by
        (track, time) (the reference's DataFrame)."""
        raw = np.loadtxt(embedding, ndmin=2, dtype=np.float64)
        if raw.shape[1] != 130:
            raise ValueError("embedding file must have 130 columns (time track d0..d127)")
        order = np.lexsort((raw[:, 0], raw[:, 1]))
        raw = raw[order]
        time, track, X = raw[:, 0], raw[:, 1].astype(np.int64), raw[:, 2:]
        starting_point = Annotation(modality="face")
        for t in np.unique(track):
            ts = time[track == t]
            starting_point[Segment(float(ts.min()), float(ts.max())), int(t)] = int(t)
        return starting_point, dict(time=time, track=track, X=X)


class FaceClustering(object):
    """Face clustering

    Parameters
    ----------
    threshold : float, optional
        Defaults to 0.6.
    force : bool, optional
        Passed by the reference to pyannote.algorithms' `DistanceThreshold(threshold, force)`
        (clustering.py:140-141).  With `constraint=None` (clustering.py:143-144) no merge is ever vetoed, so the
        flag cannot change the partition at the threshold [MEMORY of pyannote.algorithms 0.8 — unverified]; it is
        accepted and recorded.

    Usage
    -----
    >>> clustering = FaceClustering()
    >>> starting_point, features = clustering.model.preprocess(embedding)
    >>> result = clustering(starting_point, features=features)
    """

    def __init__(self, threshold=0.6, force=False, logger=None, metric="euclidean"):
        self.threshold = threshold
        self.force = bool(force)
        self.metric = metric
        self.logger = logger
        self.model = _Model()

    def __call__(self, starting_point, features=None):
        if isinstance(starting_point, Annotation):
            wanted = sorted(t for _, t in starting_point.itertracks())
        else:
            wanted = sorted(starting_point)
        result = Annotation(modality="face") if not isinstance(starting_point, Annotation) else \
            Annotation(starting_point.uri, starting_point.modality)
        if not wanted:
            return result
        keep = np.isin(features["track"], np.asarray(wanted, np.int64))
        X = features["X"][keep].astype(np.float32)
        tr = features["track"][keep]
        tracks, labels = cluster(X, tr, threshold=self.threshold, metric=self.metric)
        lab = {int(t): int(l) for t, l in zip(tracks, labels)}
        if isinstance(starting_point, Annotation):
            for segment, track in starting_point.itertracks():
                result[segment, track] = lab[int(track)]
        else:
            for track in wanted:
                result[Segment(*starting_point[track]), track] = lab[int(track)]
        return result
