"""GPU: csrc/hog.cu + the 10x10 rsconv instance (dlib's HOG frontal detector, the detector the reference really calls,
pyannote/video/face/face.py:54,66) against oracle/hog.py on the same pyramid plane: features bit-exact (the bf16 the
kernel stores == the oracle's float32 rounded to bf16), score maps within bf16-product tolerance, detections identical."""
import numpy as np
import pytest
import torch

from oracle import hog as OH
from pyannote_video_b200 import weights as Wt
from pyannote_video_b200.hog import HogDetectorNet
from pyannote_video_b200.synth import make_frames

pytestmark = pytest.mark.gpu


def _levels(net, b):
    plane = net.plane[b].cpu().numpy()
    out = []
    for k in range(net.hgeo.n_levels):
        L = net.hgeo.lv[k]
        out.append(np.ascontiguousarray(plane[L.y0:L.y0 + L.h, L.x0:L.x0 + L.w, :3]))
    return out


@pytest.mark.parametrize("H,W,upsample", [(270, 480, 1), (200, 333, 0)])
def test_hog_features_scores_and_detections_match_oracle(cuda, H, W, upsample):
    B = 2
    model = Wt.make_hog_detector(seed=5, threshold=1e9)          # thresholds are set from the oracle's scores below
    net = HogDetectorNet(model, H, W, upsample, B, cuda)
    frames = make_frames(B, H, W, seed=7, device=cuda)
    net.build_plane(frames, B)
    net.forward_scores(B)
    net.check()
    filt = np.asarray(model["filters"], np.float32)
    all_scores = []
    for b in range(B):
        for k, img in enumerate(_levels(net, b)):
            feat = OH.fhog_features(OH.fhog_hist(img))
            want = torch.from_numpy(feat).to(torch.bfloat16)
            got = net.level_features(b, k).cpu()
            assert got.shape == want.shape
            assert torch.equal(got.view(torch.int16), want.view(torch.int16)), (b, k)
            sc = OH.score_maps(want.float().numpy(), filt)                   # oracle scores of the SAME bf16 features
            gs = net.level_scores(b, k).cpu().numpy()
            assert gs.shape == sc.shape
            assert np.abs(gs - sc).max() <= 2e-3 + 1e-3 * np.abs(sc).max(), (b, k, float(np.abs(gs - sc).max()))
            all_scores.append(sc.reshape(-1))
    # a threshold that ~30 windows pass, moved away from any score by more than the comparison tolerance
    s = np.sort(np.concatenate(all_scores))[::-1]
    thr = float(s[30])
    while np.any(np.abs(s - thr) < 5e-3):
        thr += 2e-3
    net.thr.fill_(thr)
    boxes, scores, counts = net.decode(B)
    counts = counts.cpu().numpy()
    assert (counts > 0).any()
    for b in range(B):
        # oracle detections from the oracle's own features / scores of the plane's levels
        lv_imgs = _levels(net, b)
        full = [None] * (max(net.levels) + 1)
        ob, os_, ow = [], [], []
        cands = []
        for k, img in enumerate(lv_imgs):
            feat = torch.from_numpy(OH.fhog_features(OH.fhog_hist(img))).to(torch.bfloat16).float().numpy()
            sc = OH.score_maps(feat, filt)
            for d in range(filt.shape[0]):
                ys, xs = np.nonzero(sc[d] >= np.float32(thr))
                for y, x in zip(ys, xs):
                    box = OH.rect_up(OH.level_box(int(y) + 5, int(x) + 5), net.levels[k], bool(upsample))
                    cands.append((float(sc[d, y, x]), k, d, int(y), int(x), box))
        cands.sort(key=lambda t: (-t[0], t[1], t[2], t[3], t[4]))
        for c in cands:
            if any(OH.box_overlap(c[5], kb) for kb in ob):
                continue
            ob.append(c[5]); os_.append(c[0]); ow.append(c[2])
        n = int(counts[b])
        assert n == len(ob), (b, n, len(ob))
        assert boxes[b, :n].cpu().numpy().tolist() == [list(x) for x in ob]
        assert net.out_which[b, :n].cpu().numpy().tolist() == ow
        assert np.allclose(scores[b, :n].cpu().numpy(), np.asarray(os_, np.float32), atol=5e-3)


def test_face_with_a_hog_detector_model(cuda):
    """`Face(detector=<hog model>)`: the reference's detector family behind the reference's API"""
    from pyannote_video_b200.face import Face
    model = Wt.make_hog_detector(seed=5, threshold=1e9)
    face = Face(detector=model, upsample=1, device=cuda, max_frames=2)
    frames = make_frames(2, 270, 480, seed=7, device=cuda)
    det = face._detector_for(270, 480)
    det.build_plane(frames, 2)
    sc = det.forward_scores(2)
    thr = float(torch.sort(sc[..., :det.D].reshape(-1), descending=True).values[40])
    det.thr.fill_(thr)
    boxes, fidx, scores = face.detect_batch(frames)
    assert boxes.shape[0] == fidx.shape[0] == scores.shape[0] > 0
    assert bool((scores >= thr - 1e-6).all())
    rects = list(face.iterfaces(frames[0].cpu().numpy()))
    assert len(rects) == int((fidx == 0).sum())
    r0 = rects[0]
    assert [r0.left(), r0.top(), r0.right(), r0.bottom()] == boxes[fidx == 0][0].cpu().numpy().tolist()
