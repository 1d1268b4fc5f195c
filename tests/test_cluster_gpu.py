"""GPU parity: pairwise distances and threshold-stopped average-linkage clustering against scipy
(`pdist`, exactly what the reference calls, pyannote/video/face/clustering.py:101) and the greedy
oracle.  Partitions are compared up to label permutation."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import hac as ohac

pytestmark = pytest.mark.gpu


def _data(seed, n_ids=9, tracks_per_id=4, emb_per_track=(1, 6), sigma=0.02, scale=0.5):
    rng = np.random.default_rng(seed)
    cent = rng.standard_normal((n_ids, 128)) * scale / np.sqrt(128) * 8
    X, tr = [], []
    t = 0
    for c in cent:
        for _ in range(tracks_per_id):
            k = int(rng.integers(emb_per_track[0], emb_per_track[1] + 1))
            X.append(c + sigma * rng.standard_normal((k, 128)))
            tr += [t] * k
            t += 1
    X = np.concatenate(X)
    perm = rng.permutation(len(X))
    return X[perm].astype(np.float32), np.asarray(tr)[perm] * 3 + 7   # non-contiguous ids, shuffled rows


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_pdist_matches_scipy(cuda, metric):
    from scipy.spatial.distance import pdist, squareform
    from pyannote_video_b200 import _lib
    X, _ = _data(0)
    Xd = torch.from_numpy(X).to(cuda)
    n = X.shape[0]
    D = torch.empty(n, n, dtype=torch.float32, device=cuda)
    _lib.check(_lib.lib().pv_pdist(_lib.ptr(Xd), C.c_int64(n), 128, 0 if metric == "euclidean" else 1, _lib.ptr(D),
                                   _lib.stream_ptr()))
    ref = squareform(pdist(X.astype(np.float64), metric=metric))
    assert np.allclose(D.cpu().numpy(), ref, atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
@pytest.mark.parametrize("n", [300, 1000, 1343])
def test_tcgen05_gram_distances_match_scipy(cuda, metric, n):
    """csrc/gram.cu (bf16 x 3 split Gram on the tensor cores + fused distance epilogue) against scipy's float64 pdist —
    the reference's own call — and against the fp32 CUDA-core kernel; n not a multiple of 128 / 8 exercises the tails"""
    from scipy.spatial.distance import pdist, squareform
    from pyannote_video_b200.clustering import pairwise_distances
    rng = np.random.default_rng(n)
    cent = rng.standard_normal((12, 128)) * 0.35
    X = (cent[rng.integers(0, 12, n)] + 0.03 * rng.standard_normal((n, 128))).astype(np.float32)
    Xd = torch.from_numpy(X).to(cuda)
    D = pairwise_distances(Xd, metric).cpu().numpy()
    ref = squareform(pdist(X.astype(np.float64), metric=metric))
    assert D.shape == ref.shape and (np.diag(D) == 0).all()
    err = np.abs(D - ref)
    print("gram %s n=%d: max |dD| %.2e (D range %.3f..%.3f)" % (metric, n, err.max(), ref[ref > 0].min(), ref.max()))
    assert err.max() < 2e-5
    D32 = pairwise_distances(Xd, metric, impl="fp32").cpu().numpy()
    assert np.abs(D - D32).max() < 2e-5
    assert np.abs(D - D.T).max() < 2e-5          # D[i][j] and D[j][i] accumulate the six products in different orders


@pytest.mark.parametrize("seed,metric,thr", [(1, "euclidean", 0.6), (2, "euclidean", 0.6), (3, "cosine", 0.05)])
def test_cluster_matches_greedy_oracle(cuda, seed, metric, thr):
    from pyannote_video_b200.clustering import cluster
    X, tr = _data(seed)
    tracks, labels, stats = cluster(X, tr, threshold=thr, metric=metric, device=cuda, return_stats=True)
    ref = ohac.greedy_hac(X.astype(np.float64), tr, threshold=thr, metric=metric)
    got = {int(t): int(l) for t, l in zip(tracks, labels)}
    assert ohac.partition_of(got) == ohac.partition_of(ref)
    assert got == ref                                  # labels too: smallest track id of the cluster
    assert 1 < stats["n_clusters"] < len(tracks)


def test_cluster_singletons_large(cuda):
    """size-independent property at a larger size: every merge the GPU made respects the cut, and
    the partition equals scipy's average-linkage cut (independent oracle)."""
    from scipy.cluster.hierarchy import linkage, fcluster
    from pyannote_video_b200.clustering import cluster
    rng = np.random.default_rng(9)
    cent = rng.standard_normal((40, 128)) * 0.35
    X = np.concatenate([c + 0.02 * rng.standard_normal((25, 128)) for c in cent]).astype(np.float32)
    tracks, labels = cluster(X, np.arange(len(X)), threshold=0.6, device=cuda)
    Z = fcluster(linkage(X.astype(np.float64), "average"), t=0.6, criterion="distance")
    assert ohac.partition_of(dict(zip(tracks.tolist(), labels.tolist()))) == ohac.partition_of(dict(enumerate(Z.tolist())))


def test_face_clustering_file_roundtrip(cuda, tmp_path):
    from pyannote_video_b200.clustering import FaceClustering
    X, tr = _data(4, n_ids=4, tracks_per_id=3, emb_per_track=(2, 4))
    path = tmp_path / "emb.txt"
    with open(path, "w") as f:
        for i, (x, t) in enumerate(zip(X, tr)):
            f.write("%.3f %d" % (0.04 * i, t) + "".join(" %.5f" % v for v in x) + "\n")
    clustering = FaceClustering(threshold=0.6)
    starting_point, features = clustering.model.preprocess(str(path))
    result = clustering(starting_point, features=features)
    ref = ohac.greedy_hac(features["X"], features["track"], threshold=0.6)
    from pyannote_video_b200.annotation import Annotation
    assert isinstance(starting_point, Annotation) and isinstance(result, Annotation)
    kept = set(t for _, t in starting_point.itertracks())
    assert ohac.partition_of(result.to_dict()) == ohac.partition_of({k: v for k, v in ref.items() if k in kept})
    # the notebook's loop (doc/getting_started.ipynb cell 21): itertracks(yield_label=True), label = a surviving track id
    for segment, track, label in result.itertracks(yield_label=True):
        assert label in kept and segment.end > segment.start
    assert FaceClustering(threshold=0.6, force=True)(starting_point, features=features) == result


def test_whole_stage_c_entry_point_matches_the_python_loop(cuda):
    """pv_hac_threshold (one C call: tcgen05 Gram -> rounds of argmin / plan / contract on the device) gives the partition
    of the Python-driven loop and of scipy's average linkage cut"""
    from scipy.cluster.hierarchy import linkage, fcluster
    from pyannote_video_b200 import _lib
    from pyannote_video_b200.clustering import cluster
    rng = np.random.default_rng(5)
    cent = rng.standard_normal((30, 128)) * 0.35
    X = np.concatenate([c + 0.02 * rng.standard_normal((20, 128)) for c in cent]).astype(np.float32)
    X = X[rng.permutation(len(X))]
    n = len(X)
    Xd = torch.from_numpy(X).to(cuda)
    labels = torch.empty(n, dtype=torch.int32, device=cuda)
    rounds = C.c_int(0)
    _lib.check(_lib.lib().pv_hac_threshold(_lib.ptr(Xd), C.c_int64(n), 128, 0, C.c_float(0.6), 0, _lib.ptr(labels),
                                           C.byref(rounds), _lib.stream_ptr()), "pv_hac_threshold")
    got = dict(enumerate(labels.cpu().tolist()))
    tracks, lab = cluster(X, np.arange(n), threshold=0.6, device=cuda)
    assert got == dict(zip(tracks.tolist(), lab.tolist())) and rounds.value > 0
    Z = fcluster(linkage(X.astype(np.float64), "average"), t=0.6, criterion="distance")
    assert ohac.partition_of(got) == ohac.partition_of(dict(enumerate(Z.tolist())))
