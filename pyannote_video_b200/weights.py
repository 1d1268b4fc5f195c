"""Model containers + seeded synthetic weights for the three dlib models on the path.

The reference loads dlib `.dat` files (pyannote/video/face/face.py:57-62; README.md:27-30):
  * mmod_human_face_detector.dat            (CNN/MMOD detector, per BASELINE.json north_star)
  * shape_predictor_68_face_landmarks.dat   (ERT cascade)
  * dlib_face_recognition_resnet_model_v1.dat (29-conv ResNet, 128-d)
None of them (nor dlib) exists in the build environment, so models are stored in our own
container (`.npz`: flat arrays + a `kind` tag) and generated with fixed seeds.  Architectures
follow SURVEY.md App. A.1/A.3/A.4 [MEMORY of dlib 19.12 — unverified].
"""
import math

import numpy as np

# --- architecture constants (each is a restatement of a dlib detail; audit list in DESIGN.md) ---
DET_CONVS = [  # (cout, cin, k, stride)   dlib: con5d<16>, con5d<32>, con5d<32>, con5<45> x3, con<1,9,9,1,1>
    (16, 3, 5, 2), (32, 16, 5, 2), (32, 32, 5, 2), (45, 32, 5, 1), (45, 45, 5, 1), (45, 45, 5, 1), (1, 45, 9, 1)]
DET_WINDOW = 40            # MMOD detector window (pixels, at the matching pyramid level)
DET_IOU_THRESH = 0.4       # overlaps_nms: test_box_overlap(iou_thresh, percent_covered_thresh)
DET_COVERED_THRESH = 1.0
PIXEL_MEAN = (122.782, 117.001, 104.298)
PIXEL_SCALE = 1.0 / 256.0

EMB_LEVELS = [  # (channels, n_plain_residual_blocks, has_down_block) applied in this order
    (32, 3, False), (64, 3, True), (128, 2, True), (256, 2, True), (256, 0, True)]
EMB_CHIP = 150
EMB_CHIP_PADDING = 0.25
EMB_DIM = 128

ERT_STAGES = 15
ERT_TREES = 500
ERT_DEPTH = 4
ERT_POINTS = 68
ERT_POOL = 400


def conv_pad(k, stride):
    """dlib con_/pool default padding: stride != 1 ? 0 : k/2   [MEMORY]"""
    return 0 if stride != 1 else k // 2


def _conv(rng, cout, cin, k, gamma=(0.8, 1.2), gain=2.0):
    std = math.sqrt(gain / (cin * k * k))
    return dict(
        w=(rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32),
        b=(rng.standard_normal(cout) * 0.05).astype(np.float32),
        gamma=rng.uniform(gamma[0], gamma[1], cout).astype(np.float32),
        beta=(rng.standard_normal(cout) * 0.05).astype(np.float32),
    )


def make_detector(seed=2, score_bias=-3.0):
    """Synthetic MMOD detector.  The last conv has bias only (no affine)."""
    rng = np.random.default_rng(seed)
    convs = []
    for i, (cout, cin, k, s) in enumerate(DET_CONVS):
        c = _conv(rng, cout, cin, k, gain=2.0 if i < len(DET_CONVS) - 1 else 1.0)
        if i == len(DET_CONVS) - 1:
            c["gamma"] = np.ones(cout, np.float32)
            c["beta"] = np.zeros(cout, np.float32)
            c["b"] = np.full(cout, score_bias, np.float32)
        convs.append(c)
    return dict(kind="mmod_detector", convs=convs, window=DET_WINDOW, iou_thresh=DET_IOU_THRESH,
                covered_thresh=DET_COVERED_THRESH, adjust_threshold=0.0)


def embed_block_list():
    """[(type, channels_in, channels_out)] in execution order."""
    blocks = []
    cin = 32
    for ch, n_plain, down in EMB_LEVELS:
        if down:
            blocks.append(("ares_down", cin, ch))
            cin = ch
        for _ in range(n_plain):
            blocks.append(("ares", cin, ch))
    return blocks


def make_embedder(seed=3):
    rng = np.random.default_rng(seed)
    conv1 = _conv(rng, 32, 3, 7)
    blocks = []
    for typ, cin, ch in embed_block_list():
        a = _conv(rng, ch, cin, 3)
        b = _conv(rng, ch, ch, 3, gamma=(0.3, 0.5), gain=1.0)
        blocks.append(dict(type=typ, a=a, b=b))
    fc = (rng.standard_normal((EMB_DIM, 256)) * math.sqrt(1.0 / 256)).astype(np.float32)
    return dict(kind="resnet_v1_embedder", conv1=conv1, blocks=blocks, fc=fc)


def mean_shape():
    """A synthetic 68-point mean shape in the unit square (jaw arc, brows, nose, eyes, mouth)."""
    pts = []
    for i in range(17):  # jaw
        a = math.pi * (0.05 + 0.9 * i / 16)
        pts.append((0.5 - 0.45 * math.cos(a), 0.35 + 0.6 * math.sin(a)))
    for i in range(5):
        pts.append((0.15 + 0.06 * i, 0.28 - 0.02 * math.sin(math.pi * i / 4)))
    for i in range(5):
        pts.append((0.61 + 0.06 * i, 0.28 - 0.02 * math.sin(math.pi * i / 4)))
    for i in range(4):
        pts.append((0.5, 0.35 + 0.07 * i))
    for i in range(5):
        pts.append((0.40 + 0.05 * i, 0.62 + 0.01 * math.sin(math.pi * i / 4)))
    for cx in (0.27, 0.73):
        for i in range(6):
            a = 2 * math.pi * i / 6
            pts.append((cx + 0.07 * math.cos(a), 0.38 + 0.03 * math.sin(a)))
    for i in range(12):
        a = 2 * math.pi * i / 12
        pts.append((0.5 + 0.16 * math.cos(a), 0.78 + 0.07 * math.sin(a)))
    for i in range(8):
        a = 2 * math.pi * i / 8
        pts.append((0.5 + 0.10 * math.cos(a), 0.78 + 0.03 * math.sin(a)))
    assert len(pts) == ERT_POINTS
    return np.asarray(pts, np.float32)


def make_shape_predictor(seed=4, stages=ERT_STAGES, trees=ERT_TREES, pool=ERT_POOL, leaf_sigma=0.0015):
    rng = np.random.default_rng(seed)
    n_split = (1 << ERT_DEPTH) - 1
    n_leaf = 1 << ERT_DEPTH
    return dict(
        kind="ert_shape_predictor",
        initial_shape=mean_shape().reshape(-1).copy(),                       # [136] x0,y0,x1,y1...
        anchor_idx=rng.integers(0, ERT_POINTS, (stages, pool)).astype(np.int32),
        deltas=(rng.standard_normal((stages, pool, 2)) * 0.08).astype(np.float32),
        split_idx1=rng.integers(0, pool, (stages, trees, n_split)).astype(np.int32),
        split_idx2=rng.integers(0, pool, (stages, trees, n_split)).astype(np.int32),
        split_thresh=(rng.standard_normal((stages, trees, n_split)) * 25.0).astype(np.float32),
        leaf_values=(rng.standard_normal((stages, trees, n_leaf, 2 * ERT_POINTS)) * leaf_sigma).astype(np.float32),
    )


# get_face_chip_details (dlib image_transforms/interpolation.h): mean_face_shape_x/y, 51 constants for
# landmarks 17..67 [MEMORY of dlib 19.12].  The oracle keeps its own copy (oracle/constants.py).
MEAN_FACE_X = (
    0.000213256, 0.0752622, 0.18113, 0.29077, 0.393397, 0.586856, 0.689483, 0.799124,
    0.904991, 0.98004, 0.490127, 0.490127, 0.490127, 0.490127, 0.36688, 0.426036,
    0.490127, 0.554217, 0.613373, 0.121737, 0.187122, 0.265825, 0.334606, 0.260918,
    0.182743, 0.645647, 0.714428, 0.793132, 0.858516, 0.79751, 0.719335, 0.254149,
    0.340985, 0.428858, 0.490127, 0.551395, 0.639268, 0.726104, 0.642159, 0.556721,
    0.490127, 0.423532, 0.338094, 0.290379, 0.428096, 0.490127, 0.552157, 0.689874,
    0.553364, 0.490127, 0.42689)
MEAN_FACE_Y = (
    0.106454, 0.038915, 0.0187482, 0.0344891, 0.0773906, 0.0773906, 0.0344891,
    0.0187482, 0.038915, 0.106454, 0.203352, 0.307009, 0.409805, 0.515625, 0.587326,
    0.609345, 0.628106, 0.609345, 0.587326, 0.216423, 0.178758, 0.179852, 0.231733,
    0.245099, 0.244077, 0.231733, 0.179852, 0.178758, 0.216423, 0.244077, 0.245099,
    0.780233, 0.745405, 0.727388, 0.742578, 0.727388, 0.745405, 0.780233, 0.864805,
    0.902192, 0.909281, 0.902192, 0.864805, 0.784792, 0.778746, 0.785343, 0.778746,
    0.784792, 0.824182, 0.831803, 0.824182)
# landmarks that take part in the chip alignment: 17..67 minus eyebrows (17..26) and lower lip (55..59, 65..67)
CHIP_POINTS = tuple(i for i in range(17, 68)
                    if not (17 <= i <= 26) and not (55 <= i <= 59) and not (65 <= i <= 67))


def chip_mean_face():
    return np.stack([np.asarray(MEAN_FACE_X, np.float32), np.asarray(MEAN_FACE_Y, np.float32)], axis=1)


# ---------------------------------------------------------------------------------------------
# container I/O (.npz)
# ---------------------------------------------------------------------------------------------
def _flatten(prefix, obj, out):
    if isinstance(obj, dict):
        for k, v in obj.items():
            _flatten(prefix + k + "/", v, out)
    elif isinstance(obj, (list, tuple)):
        out[prefix + "__len__"] = np.asarray(len(obj))
        for i, v in enumerate(obj):
            _flatten(prefix + str(i) + "/", v, out)
    else:
        out[prefix[:-1]] = np.asarray(obj)


def save_model(path, model):
    flat = {}
    _flatten("", model, flat)
    with open(path, "wb") as f:
        np.savez(f, **flat)


def _unflatten(flat):
    root = {}
    for key, val in flat.items():
        parts = key.split("/")
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = val

    def fix(d):
        if not isinstance(d, dict):
            if isinstance(d, np.ndarray) and d.dtype.kind in "US":
                return str(d)
            if isinstance(d, np.ndarray) and d.ndim == 0:
                return d.item()
            return d
        if "__len__" in d:
            return [fix(d[str(i)]) for i in range(int(d["__len__"]))]
        return {k: fix(v) for k, v in d.items()}

    return fix(root)


def load_model(path, kind=None):
    with np.load(path, allow_pickle=False) as z:
        model = _unflatten({k: z[k] for k in z.files})
    if kind is not None and model.get("kind") != kind:
        raise RuntimeError("model file %s holds a '%s', expected '%s'" % (path, model.get("kind"), kind))
    return model


# ---------------------------------------------------------------------------------------------------------------
# HOG frontal detector (dlib.get_frontal_face_detector, pyannote/video/face/face.py:54).  dlib compiles its five
# trained filters into the library as a base64 blob — not obtainable here — so the tests use seeded random filters.
# ---------------------------------------------------------------------------------------------------------------
HOG_CELL = 8
HOG_FILTER = 10          # cells: 80 x 80 detection window, padding 1  [MEMORY]
HOG_NMS_IOU = 0.5        # [MEMORY] test_box_overlap of the serialised detector
HOG_NMS_COVERED = 1.0


def make_hog_detector(seed=5, n_filters=5, threshold=None):
    """seeded random HOG detector: {"kind": "hog_detector", "filters": f32 [D,31,10,10], "thresholds": f32 [D]}.
    The default thresholds are set so that a few windows per 1080p frame fire on the synthetic test frames."""
    rng = np.random.default_rng(seed)
    filt = (rng.standard_normal((n_filters, 31, HOG_FILTER, HOG_FILTER)) * 0.05).astype(np.float32)
    # round the filters to bf16: the CUDA path multiplies bf16 operands (fp32 accumulate), the oracle then sees the same weights
    import torch
    filt = torch.from_numpy(filt).to(torch.bfloat16).float().numpy()
    thr = np.full(n_filters, 2.3 if threshold is None else float(threshold), np.float32)
    return {"kind": "hog_detector", "filters": filt, "thresholds": thr, "iou_thresh": HOG_NMS_IOU, "covered_thresh": HOG_NMS_COVERED}
