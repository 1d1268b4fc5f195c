"""GPU: size-independent properties at (or near) BASELINE.json's full configuration sizes, where the
CPU oracle would take too long:
  C3  256 concurrent trackers on 1080p frames with a known global translation
  C5  clustering of 100k x 128-d embeddings planted around well-separated centroids
  C2  a 1080p frame through the whole detector (sizes, determinism, batch independence)
"""
import numpy as np
import pytest
import torch

from pyannote_video_b200.geometry import DRect
from pyannote_video_b200.synth import make_frames

pytestmark = pytest.mark.gpu


def test_c3_256_trackers_follow_known_translation(cuda):
    from pyannote_video_b200.tracker import TrackerBank
    n_frames, shift = 40, (1.0, 0.5)
    frames = make_frames(n_frames, 1080, 1920, seed=11, device=cuda, shift_per_frame=shift)
    bank = TrackerBank(capacity=256, device=cuda)
    rects = [(100.0 + 110 * (k % 16), 60.0 + 62 * (k // 16), 196.0 + 110 * (k % 16), 156.0 + 62 * (k // 16)) for k in range(256)]
    handles = [bank.start(frames[0], DRect(*r)) for r in rects]
    psr_min = 1e9
    for i in range(1, n_frames):
        conf = bank.update(frames[i], handles)
        psr_min = min(psr_min, min(conf))
    pos = np.asarray([[bank.position(h).left(), bank.position(h).top()] for h in handles])
    start = np.asarray([[r[0], r[1]] for r in rects])
    # the crop window moves by +shift per frame, so image content (and the trackers) move by -shift
    expect = start - np.asarray(shift) * (n_frames - 1)
    err = np.abs(pos - expect).max()
    assert err < 1.5, err
    assert psr_min > 10, psr_min


def test_c5_100k_embeddings_recover_planted_clusters(cuda):
    from pyannote_video_b200.clustering import cluster
    g = torch.Generator(device="cpu").manual_seed(3)
    n_cent, per = 2000, 50
    cent = torch.randn(n_cent, 128, generator=g)
    cent = cent / cent.norm(dim=1, keepdim=True) * 0.9          # pairwise centroid distance ~ 1.27 >> 0.6
    X = (cent[:, None, :] + 0.02 * torch.randn(n_cent, per, 128, generator=g)).reshape(-1, 128)
    perm = torch.randperm(X.shape[0], generator=g)
    truth = torch.arange(n_cent).repeat_interleave(per)[perm].numpy()
    X = X[perm].contiguous()
    tracks, labels, stats = cluster(X.to(cuda), np.arange(X.shape[0]), threshold=0.6, device=cuda, return_stats=True)
    assert stats["n_clusters"] == n_cent
    # same partition as the planted one: each label maps to exactly one centroid and vice versa
    pairs = set(zip(labels.tolist(), truth.tolist()))
    assert len(pairs) == n_cent
    # idempotence: clustering one representative per cluster merges nothing
    reps = np.unique(labels)
    t2, l2 = cluster(X[torch.from_numpy(reps)].to(cuda), reps, threshold=0.6, device=cuda)
    assert len(np.unique(l2)) == n_cent


def test_c2_full_1080p_detector_is_deterministic_and_batch_independent(cuda):
    from pyannote_video_b200 import weights as W
    from pyannote_video_b200.nets import DetectorNet
    model = W.make_detector(seed=2, score_bias=0.0)
    frames = make_frames(2, 1080, 1920, seed=0, device=cuda)
    net = DetectorNet(model, 1080, 1920, 1, max_batch=2, device=cuda)
    net.build_plane(frames, 2)
    s2 = net.forward_scores(2).clone()
    net.check()
    assert s2.shape[1:] == (net.OH, net.OW) and torch.isfinite(s2).all()
    # a frame's scores do not depend on its batch mates (no cross-image leakage through row offsets)
    net.build_plane(frames[1:2], 1)
    s1 = net.forward_scores(1).clone()
    assert torch.equal(s1[0], s2[1])
    # plane borders (padding) are zero in the conv1 input and stay zero
    assert int(net.plane[0, :11].max()) == 0 and int(net.plane[0, :, :11].max()) == 0


def test_c2_full_1080p_three_conv_implementations_agree(cuda):
    """at full size the CPU oracle is too slow, but the three independent tensor-core formulations of the detector
    convs (row streaming, 2-D tiles, 1-D shifted rows) must produce the same score map up to fp32 summation order"""
    from pyannote_video_b200 import weights as W
    from pyannote_video_b200.nets import DetectorNet
    model = W.make_detector(seed=2, score_bias=0.0)
    frames = make_frames(1, 1080, 1920, seed=3, device=cuda)
    ref = None
    for impl in ("rsconv", "detconv", "srgemm"):
        net = DetectorNet(model, 1080, 1920, 1, max_batch=1, device=cuda, conv_impl=impl)
        net.build_plane(frames, 1)
        s = net.forward_scores(1).clone()
        net.check()
        assert torch.isfinite(s).all()
        if ref is None:
            ref = s
        else:
            d = float((s - ref).abs().max())
            assert d < 0.02 * max(1.0, float(ref.abs().max())), (impl, d)
        del net
        torch.cuda.empty_cache()


def test_c4_one_4k_frame_through_the_detector(cuda):
    """BASELINE.json configs[3] frame size (3840x2160, upsample 1: a 6852 x 16760 plane, 108 M pyramid pixels):
    index arithmetic at the largest size, determinism, zero borders"""
    from pyannote_video_b200 import weights as W
    from pyannote_video_b200.nets import DetectorNet
    model = W.make_detector(seed=2)          # score bias -3: few candidates (with bias 0 decode reports overflow as a negative count)
    frames = make_frames(1, 2160, 3840, seed=1, device=cuda)
    net = DetectorNet(model, 2160, 3840, 1, max_batch=1, device=cuda)
    net.build_plane(frames, 1)
    a = net.forward_scores(1).clone()
    net.check()
    net.build_plane(frames, 1)
    b = net.forward_scores(1).clone()
    assert torch.isfinite(a).all() and torch.equal(a, b)
    assert a.shape[1:] == (net.OH, net.OW)
    boxes, scores, counts = net.decode(1)
    assert 0 <= int(counts[0]) <= net.MAX_DET
    assert int(net.plane[0, :11].max()) == 0 and int(net.plane[0, -11:].max()) == 0
