"""ORACLE (test infrastructure — never imported by the product path).

The oracle's OWN copy of every dlib constant it restates (dlib 19.12, recalled — SURVEY.md App. A;
dlib itself is absent from the build environment: parity unpinned).  The product keeps its copy in
pyannote_video_b200/weights.py / pyrgeom.py / tracker.py; tests/test_oracle_cpu.py asserts the two
copies agree, so that neither side can silently drift and a wrong value on ONE side is visible.
"""
import numpy as np

# --- MMOD CNN face detector (mmod_human_face_detector.dat): con5d<16>, con5d<32>, con5d<32>, con5<45> x3, con<1,9,9,1,1>
DET_CONVS = [  # (cout, cin, k, stride)
    (16, 3, 5, 2), (32, 16, 5, 2), (32, 32, 5, 2), (45, 32, 5, 1), (45, 45, 5, 1), (45, 45, 5, 1), (1, 45, 9, 1)]
DET_WINDOW = 40
DET_IOU_THRESH = 0.4
DET_COVERED_THRESH = 1.0
# input_rgb_image(_pyramid): (v - avg) / 256 per channel
PIXEL_MEAN = (122.782, 117.001, 104.298)
PIXEL_SCALE = 1.0 / 256.0

# --- input_rgb_image_pyramid<pyramid_down<6>>
PYR_N = 6
PYR_PAD = 10
PYR_OUTER_PAD = 11
PYR_MIN_SIDE = 5

# --- face_recognition_resnet_model_v1 (anet_type)
EMB_LEVELS = [(32, 3, False), (64, 3, True), (128, 2, True), (256, 2, True), (256, 0, True)]
EMB_CHIP = 150
EMB_CHIP_PADDING = 0.25
EMB_DIM = 128

# --- shape_predictor_68_face_landmarks
ERT_POINTS = 68

# get_face_chip_details (image_transforms/interpolation.h): 51 mean-face constants for landmarks 17..67
MEAN_FACE_X = [
    0.000213256, 0.0752622, 0.18113, 0.29077, 0.393397, 0.586856, 0.689483, 0.799124,
    0.904991, 0.98004, 0.490127, 0.490127, 0.490127, 0.490127, 0.36688, 0.426036,
    0.490127, 0.554217, 0.613373, 0.121737, 0.187122, 0.265825, 0.334606, 0.260918,
    0.182743, 0.645647, 0.714428, 0.793132, 0.858516, 0.79751, 0.719335, 0.254149,
    0.340985, 0.428858, 0.490127, 0.551395, 0.639268, 0.726104, 0.642159, 0.556721,
    0.490127, 0.423532, 0.338094, 0.290379, 0.428096, 0.490127, 0.552157, 0.689874,
    0.553364, 0.490127, 0.42689]
MEAN_FACE_Y = [
    0.106454, 0.038915, 0.0187482, 0.0344891, 0.0773906, 0.0773906, 0.0344891,
    0.0187482, 0.038915, 0.106454, 0.203352, 0.307009, 0.409805, 0.515625, 0.587326,
    0.609345, 0.628106, 0.609345, 0.587326, 0.216423, 0.178758, 0.179852, 0.231733,
    0.245099, 0.244077, 0.231733, 0.179852, 0.178758, 0.216423, 0.244077, 0.245099,
    0.780233, 0.745405, 0.727388, 0.742578, 0.727388, 0.745405, 0.780233, 0.864805,
    0.902192, 0.909281, 0.902192, 0.864805, 0.784792, 0.778746, 0.785343, 0.778746,
    0.784792, 0.824182, 0.831803, 0.824182]
assert len(MEAN_FACE_X) == 51 and len(MEAN_FACE_Y) == 51


def chip_points():
    """landmarks used for the alignment: 17..67 minus the eyebrows (17..26) and the lower lip
    (55..59 and 65..67) — 33 points"""
    out = []
    for i in range(17, 68):
        if (55 <= i <= 59) or (65 <= i <= 67):
            continue
        if 17 <= i <= 26:
            continue
        out.append(i)
    return out


def mean_face():
    return np.stack([np.asarray(MEAN_FACE_X, np.float32), np.asarray(MEAN_FACE_Y, np.float32)], axis=1)


def conv_pad(k, stride):
    """dlib con_ default padding template arguments: stride != 1 ? 0 : k/2"""
    if stride != 1:
        return 0
    return k // 2


# --- correlation_tracker defaults
TRK_FILTER = 64
TRK_PADDING = 1.4
TRK_LAMBDA = 0.001
TRK_NU = 0.025
TRK_N_SCALES = 32
TRK_SCALE_WINDOW = 23
TRK_SCALE_ALPHA = 1.020
