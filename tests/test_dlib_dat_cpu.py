"""dlib `.dat` streams (pyannote_video_b200/dlib_dat.py): primitive encodings against hand-computed
byte strings, and writer -> reader round trips of the three model kinds.  The round trips show self-consistency of a
format restated from memory of dlib's serialize.h — no real dlib file exists in the build environment to pin it."""
import math

import numpy as np
import pytest

from pyannote_video_b200 import dlib_dat as D
from pyannote_video_b200 import weights as W


def test_integer_and_float_encodings_known_answers():
    w = D.Writer()
    for v in (0, 1, 255, 256, -1, -300, 2 ** 40 + 5):
        w.int(v)
    raw = w.bytes()
    assert raw[:2] == b"\x01\x00" and raw[2:4] == b"\x01\x01" and raw[4:6] == b"\x01\xff"
    assert raw[6:9] == b"\x02\x00\x01" and raw[9:11] == b"\x81\x01" and raw[11:14] == b"\x82\x2c\x01"
    r = D.Reader(raw)
    assert [r.int() for _ in range(7)] == [0, 1, 255, 256, -1, -300, 2 ** 40 + 5]
    # float_details: 0.5 = frexp -> (0.5, 0) -> mantissa 2^23, exponent -24 -> low zero bytes shifted off: 0x80 * 2^-8
    w = D.Writer()
    w.float(0.5)
    assert w.bytes() == b"\x01\x80" + b"\x81\x08"
    w = D.Writer()
    vals = [0.0, 1.0, -2.75, 3.1415927410125732, 1e-20, -1e20, math.inf, -math.inf]
    for v in vals:
        w.float(v)
    w.float(math.nan)
    w.double(0.1)
    r = D.Reader(w.bytes())
    got = [r.float() for _ in vals]
    assert got == [float(np.float32(v)) if math.isfinite(v) else v for v in vals]
    assert math.isnan(r.float()) and r.float() == 0.1
    with pytest.raises(D.DatError):
        D.Reader(b"\x09\x00").int()


def test_tensor_matrix_string_bool():
    w = D.Writer()
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
    w.tensor(a)
    w.matrix(np.asarray([1.5, -2.0, 0.25]))
    w.string("con_4")
    w.bool(True), w.bool(False)
    r = D.Reader(w.bytes())
    assert np.array_equal(r.tensor(), a)
    assert np.array_equal(r.matrix().reshape(-1), [1.5, -2.0, 0.25])
    assert r.expect("con_4", "con_5") == "con_4" and r.bool() is True and r.bool() is False
    with pytest.raises(D.DatError, match="expected relu_"):
        D.Reader(D.Writer().bytes() + b"\x01\x03abc").expect("relu_")


def test_shape_predictor_roundtrip(tmp_path):
    m = W.make_shape_predictor(seed=4, stages=3, trees=7, pool=50)
    raw = D.dumps(m)
    p = tmp_path / "sp.dat"
    p.write_bytes(raw)
    back = D.load(str(p), "ert_shape_predictor")
    for k in ("initial_shape", "anchor_idx", "deltas", "split_idx1", "split_idx2", "split_thresh", "leaf_values"):
        assert np.array_equal(back[k], m[k]), k
    # 6 bytes per leaf float on average (mantissa 3-4 bytes + control, exponent 2): what makes the real file ~95 MB
    assert 4.0 < len(raw) / m["leaf_values"].size < 8.0


def test_embedder_and_detector_roundtrip():
    e = W.make_embedder(seed=3)
    back = D.loads(D.dumps(e), "resnet_v1_embedder")
    assert np.array_equal(back["fc"], e["fc"]) and len(back["blocks"]) == len(e["blocks"])
    for a, b in zip([e["conv1"]] + [c for blk in e["blocks"] for c in (blk["a"], blk["b"])],
                    [back["conv1"]] + [c for blk in back["blocks"] for c in (blk["a"], blk["b"])]):
        for k in ("w", "b", "gamma", "beta"):
            assert np.array_equal(a[k], b[k])
    assert [b["type"] for b in back["blocks"]] == [b["type"] for b in e["blocks"]]
    d = W.make_detector(seed=2)
    back = D.loads(D.dumps(d), "mmod_detector")
    assert back["window"] == 40 and back["iou_thresh"] == 0.4 and back["covered_thresh"] == 1.0
    for a, b in zip(d["convs"], back["convs"]):
        for k in ("w", "b", "gamma", "beta"):
            assert np.array_equal(a[k], b[k]), k
    # a truncated / foreign stream fails loudly with a byte offset
    with pytest.raises(D.DatError, match="at byte"):
        D.loads(D.dumps(d)[:5000], "mmod_detector")
    with pytest.raises(D.DatError):
        D.loads(D.dumps(e), "mmod_detector")


def test_arch_matches_the_layer_tables():
    convs = [s for s in D.anet_arch() if s[0] == "con"]
    assert len(convs) == 29 and convs[-1] == ("con", 32, 7, 2)
    assert [(s[1], s[2], s[3]) for s in reversed([s for s in D.mmod_arch() if s[0] == "con"])] == \
        [(c[0], c[2], c[3]) for c in W.DET_CONVS]
