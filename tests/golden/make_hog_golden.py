"""Generates tests/golden/hog_golden.npz: outputs of the HOG-detector oracle (oracle/hog.py) on a seeded image with seeded
filters — a regression pin of the restatement itself (dlib and its trained filters are absent: parity with dlib is
unpinned).  Run from the repository root:  python tests/golden/make_hog_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hog  # noqa: E402


def inputs():
    rng = np.random.default_rng(11)
    base = rng.random((30, 40, 3))
    img = np.kron(base, np.ones((6, 6, 1)))                       # 180 x 240, blocky structure
    img = (0.7 * img + 0.3 * rng.random(img.shape)) * 255.0
    img = img.astype(np.uint8)
    filt = (rng.standard_normal((3, 31, 10, 10)) * 0.05).astype(np.float32)
    return img, filt


def main():
    img, filt = inputs()
    levels = [img]
    feat = hog.fhog_features(hog.fhog_hist(img))
    sc = hog.score_maps(feat, filt)
    thr = float(np.sort(sc.reshape(-1))[-12])
    boxes, scores, which = hog.detect_levels(levels, filt, [thr] * filt.shape[0], upsampled=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hog_golden.npz"), feat=feat, scores=sc, thr=np.float32(thr),
                        boxes=np.asarray(boxes, np.int32), det_scores=np.asarray(scores, np.float32),
                        which=np.asarray(which, np.int32))
    print("features", feat.shape, "detections", len(boxes), "threshold", thr)


if __name__ == "__main__":
    main()
