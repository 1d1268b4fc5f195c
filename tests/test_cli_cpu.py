"""CPU: file formats and the track-file / frame re-synchronisation coroutine (reference
getFaceGenerator, scripts/pyannote-face.py:121-175; semantics listed in SURVEY.md App. B / E.3)."""
import numpy as np

from pyannote_video_b200 import cli


def _write_tracks(path):
    rows = [(0.04, 0, 0.1, 0.2, 0.3, 0.6, "detection"), (0.04, 1, 0.5, 0.1, 0.7, 0.5, "forward"),
            (0.12, 0, 0.11, 0.2, 0.31, 0.6, "forward+backward"), (0.2, 1, 0.5, 0.1, 0.7, 0.5, "backward")]
    with open(path, "w") as f:
        for t, i, l, tp, r, b, s in rows:
            f.write(cli.FACE_TEMPLATE.format(t=t, identifier=i, left=l, top=tp, right=r, bottom=b, status=s))
    return rows


def test_face_template_format():
    line = cli.FACE_TEMPLATE.format(t=1.23456, identifier=7, left=0.12345, top=0.5, right=0.75, bottom=1.0, status="forward")
    assert line == "1.235 7 0.123 0.500 0.750 1.000 forward\n"


def test_face_generator_resync_and_quirk(tmp_path):
    path = tmp_path / "t.track.txt"
    _write_tracks(path)
    W, H = 200, 100
    for quirks, n_groups in ((False, 3), (True, 2)):
        gen = cli.face_generator(str(path), W, H, reference_quirks=quirks)
        gen.send(None)
        released = []
        for i in range(8):
            t = i * 0.04
            T, faces = gen.send(t)
            if faces:
                released.append((round(t, 2), T, [(i_, f.left(), f.top(), f.right(), f.bottom()) for i_, f, _ in faces]))
        assert len(released) == n_groups
        # a group is released on the first frame whose time >= its time, labelled with ITS time
        assert released[0][0] == 0.04 and released[0][1] == 0.04 and len(released[0][2]) == 2
        assert released[1][0] == 0.12 and released[1][1] == 0.12
        # int() truncation of normalised float32 coordinates times the frame size
        assert released[0][2][0] == (0, int(np.float32(0.1) * W), int(np.float32(0.2) * H), int(np.float32(0.3) * W),
                                     int(np.float32(0.6) * H))


def test_array_video_and_shots(tmp_path):
    frames = np.zeros((5, 20, 30, 3), np.uint8)
    v = cli.ArrayVideo(frames, 25.0)
    assert v.size == (30, 20) and v.frame_size == (30, 20)
    ts = [t for t, _ in v]
    assert np.allclose(ts, [0, 0.04, 0.08, 0.12, 0.16])
    v.frame_size = (15, 10)
    assert next(iter(v))[1].shape == (10, 15, 3)
    p = tmp_path / "shots.json"
    p.write_text('[{"start": 0, "end": 1.5}, {"start": 1.5, "end": 3}]')
    shots = cli.load_shots(str(p))
    assert [s.end for s in shots] == [1.5, 3.0]
