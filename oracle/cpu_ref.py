"""ORACLE (test infrastructure — never imported by the product path).

ctypes front end of the C++ CPU restatement of the face path (oracle/cpu/*.cpp, ABI in
oracle/cpu/pvcpu.h): the "restated dlib-style CPU baseline" of BASELINE.md §2.  One thread =
B-cpu-1 (the reference is one Python thread calling dlib), all cores = B-cpu-N.  Same algorithm
as the numpy/torch oracle (oracle/pyramid.py, nets.py, landmarks.py, dsst.py), which it is
cross-checked against in tests/test_cpu_ref_cpu.py (byte / integer stages bit-exact).

The shared library is built with -march=native on the machine that runs it: its file name is
keyed on the CPU feature flags, so a tree copied to another box rebuilds there (gcc only).
"""
import ctypes as C
import hashlib
import os
import subprocess
import threading

import numpy as np

from . import constants as K
from . import geometry as ogeo
from . import landmarks as olm

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _cpu_key():
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    src = b""
    for name in ("cpu/conv.cpp", "cpu/path.cpp", "cpu/pvcpu.h", "cpu/tensor.h", "Makefile"):
        with open(os.path.join(_HERE, name), "rb") as f:
            src += f.read()
    return hashlib.sha1(flags.encode() + src).hexdigest()[:10]


def lib_path():
    return os.path.join(_HERE, "_build", "libpvcpu_%s.so" % _cpu_key())


def ensure_built():
    path = lib_path()
    if not os.path.exists(path):
        # build under a private name, then rename: another process / thread never sees a half-written library
        tmp = "tmp%d_%s" % (os.getpid(), os.path.basename(path))
        r = subprocess.run(["make", "-s", "-C", _HERE, "LIB=" + tmp], capture_output=True, text=True)
        tmp_path = os.path.join(_HERE, "_build", tmp)
        if r.returncode != 0 or not os.path.exists(tmp_path):
            raise RuntimeError("building the C++ CPU oracle failed:\n" + r.stdout + r.stderr)
        os.replace(tmp_path, path)
    return path


def host_cores():
    """usable host cores: min(online CPUs, cgroup cpu.max quota)"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


_lib_lock = threading.Lock()


def lib():
    global _lib
    if _lib is None:
        with _lib_lock:
            if _lib is None:
                l = C.CDLL(ensure_built())
                l.pvc_net_create.restype = C.c_void_p
                l.pvc_trackers_create.restype = C.c_void_p
                _lib = l
    return _lib


def set_threads(n):
    return int(lib().pvc_set_threads(int(n)))


def _p(a, typ):
    return a.ctypes.data_as(C.POINTER(typ))


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def _affine(c, affine=True):
    g, b, be = _f(c["gamma"]), _f(c["b"]), _f(c["beta"])
    if affine:
        return g, (g * b + be).astype(np.float32)
    return np.ones_like(b), b


class Net(object):
    def __init__(self):
        self.h = C.c_void_p(lib().pvc_net_create())
        self.keep = []

    def add(self, c, stride, pad, relu, affine=True, bf16=False):
        w = _f(c["w"])
        cout, cin, k, _ = w.shape
        sc, sh = _affine(c, affine)
        lib().pvc_net_add_conv(self.h, cout, cin, k, stride, pad, _p(w, C.c_float), _p(sc, C.c_float), _p(sh, C.c_float),
                               int(relu), int(bf16))

    def __del__(self):
        try:
            lib().pvc_net_destroy(self.h)
        except Exception:
            pass


class Detector(object):
    """MMOD CNN detector: rgb u8 [H,W,3] -> plane -> scores -> boxes (oracle/pyramid.py + oracle/nets.py)"""

    def __init__(self, model, bf16=False):
        self.model, self.bf16 = model, bool(bf16)
        self.net = Net()
        n = len(model["convs"])
        for i, c in enumerate(model["convs"]):
            _, _, k, s = K.DET_CONVS[i]
            last = i == n - 1
            self.net.add(c, s, K.conv_pad(k, s), relu=not last, affine=not last, bf16=bf16)
        self.mean = _f(K.PIXEL_MEAN)

    def build_plane(self, rgb, upsample=1, geo=None):
        from . import pyramid as opyr
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W, _ = rgb.shape
        if geo is None:
            geo = opyr.placement_for(H, W, upsample)
        rects = _i(np.asarray(geo.rects).reshape(-1, 4))
        plane = np.empty((geo.plane_h, geo.plane_w, 4), np.uint8)
        rc = lib().pvc_build_plane(_p(rgb, C.c_uint8), H, W, int(upsample), _p(rects, C.c_int), len(geo.rects), geo.plane_h,
                                   geo.plane_w, _p(plane, C.c_uint8))
        assert rc == 0
        return plane, geo

    def scores(self, plane):
        Hp, Wp, _ = plane.shape
        oh = int(lib().pvc_detector_out_size(self.net.h, Hp))
        ow = int(lib().pvc_detector_out_size(self.net.h, Wp))
        out = np.empty((oh, ow), np.float32)
        a, b = C.c_int(), C.c_int()
        rc = lib().pvc_detector_forward(self.net.h, _p(plane, C.c_uint8), Hp, Wp, _p(self.mean, C.c_float),
                                        C.c_float(K.PIXEL_SCALE), int(self.bf16), _p(out, C.c_float), C.byref(a), C.byref(b))
        assert rc == 0 and (a.value, b.value) == (oh, ow)
        return out

    def decode(self, scores, geo, threshold=None, max_candidates=0, max_out=4096):
        m = self.model
        if not isinstance(geo, ogeo.Geometry):
            geo = ogeo.from_product(geo)
        rects = _i(np.asarray(geo.rects).reshape(-1, 4))
        fxy = _f([geo.level_factors(lv) for lv in range(len(geo.rects))])
        mul, add = ogeo.cell_to_plane(1) - ogeo.cell_to_plane(0), ogeo.cell_to_plane(0)
        boxes = np.empty((max_out, 4), np.int32)
        sc = np.empty(max_out, np.float32)
        scores = _f(scores)
        thr = float(m["adjust_threshold"]) if threshold is None else float(threshold)
        n = lib().pvc_decode(_p(scores, C.c_float), scores.shape[0], scores.shape[1], _p(rects, C.c_int), _p(fxy, C.c_float),
                             len(geo.rects), int(m["window"]), mul, add, C.c_float(thr), C.c_double(float(m["iou_thresh"])),
                             C.c_double(float(m["covered_thresh"])), int(max_candidates), max_out, _p(boxes, C.c_int),
                             _p(sc, C.c_float))
        assert n >= 0
        return [(int(b[0]), int(b[1]), int(b[2]), int(b[3]), float(s)) for b, s in zip(boxes[:n], sc[:n])]

    def detect(self, rgb, upsample=1):
        plane, geo = self.build_plane(rgb, upsample)
        return self.decode(self.scores(plane), geo, max_candidates=4096)


class ShapePredictor(object):
    def __init__(self, model):
        self.m = dict(initial_shape=_f(model["initial_shape"]), anchor_idx=_i(model["anchor_idx"]), deltas=_f(model["deltas"]),
                      split_idx1=_i(model["split_idx1"]), split_idx2=_i(model["split_idx2"]),
                      split_thresh=_f(model["split_thresh"]), leaf_values=_f(model["leaf_values"]))
        self.stages, self.trees, n_split = self.m["split_thresh"].shape
        self.pool = self.m["anchor_idx"].shape[1]
        assert n_split == 15

    def predict(self, rgb, rects):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        rects = _i(np.asarray(rects).reshape(-1, 4))
        M = rects.shape[0]
        out = np.empty((M, 68, 2), np.int64)
        m = self.m
        lib().pvc_ert_predict(_p(rgb, C.c_uint8), rgb.shape[0], rgb.shape[1], _p(m["initial_shape"], C.c_float),
                              _p(m["anchor_idx"], C.c_int), _p(m["deltas"], C.c_float), _p(m["split_idx1"], C.c_int),
                              _p(m["split_idx2"], C.c_int), _p(m["split_thresh"], C.c_float), _p(m["leaf_values"], C.c_float),
                              self.stages, self.trees, self.pool, _p(rects, C.c_int), M, _p(out, C.c_int64))
        return out


def extract_chips(rgb, parts, size=K.EMB_CHIP):
    rgb = np.ascontiguousarray(rgb, np.uint8)
    parts = np.ascontiguousarray(parts, np.int64).reshape(-1, 68, 2)
    M = parts.shape[0]
    frm = _f(olm.chip_from_points(size))
    idx = _i(olm.CHIP_POINTS)
    out = np.empty((M, size, size, 3), np.uint8)
    lib().pvc_extract_chips(_p(rgb, C.c_uint8), rgb.shape[0], rgb.shape[1], _p(parts, C.c_int64), M, _p(frm, C.c_float),
                            _p(idx, C.c_int), len(idx), size, _p(out, C.c_uint8))
    return out


class Embedder(object):
    def __init__(self, model, bf16=False):
        self.bf16 = bool(bf16)
        self.net = Net()
        self.net.add(model["conv1"], 2, 0, relu=True, bf16=bf16)
        down = []
        for blk in model["blocks"]:
            d = blk["type"] == "ares_down"
            down.append(1 if d else 0)
            self.net.add(blk["a"], 2 if d else 1, 0 if d else 1, relu=True, bf16=bf16)
            self.net.add(blk["b"], 1, 1, relu=False, bf16=bf16)
        self.down = _i(down)
        self.fc = _f(model["fc"])
        self.mean = _f(K.PIXEL_MEAN)

    def forward(self, chips):
        chips = np.ascontiguousarray(chips, np.uint8)
        M, S = chips.shape[0], chips.shape[1]
        out = np.empty((M, 128), np.float32)
        rc = lib().pvc_embed_forward(self.net.h, _p(self.down, C.c_int), len(self.down), _p(self.fc, C.c_float),
                                     _p(chips, C.c_uint8), M, S, _p(self.mean, C.c_float), C.c_float(K.PIXEL_SCALE),
                                     int(self.bf16), _p(out, C.c_float))
        assert rc == 0
        return out


class TrackerBank(object):
    def __init__(self, capacity, use_scale=True):
        self.h = C.c_void_p(lib().pvc_trackers_create(int(capacity), int(use_scale)))

    def start(self, rgb, ids, rects):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        ids = _i(ids)
        rects = np.ascontiguousarray(rects, np.float64).reshape(-1, 4)
        lib().pvc_trackers_start(self.h, _p(rgb, C.c_uint8), rgb.shape[0], rgb.shape[1], _p(ids, C.c_int), _p(rects, C.c_double),
                                 len(ids))

    def update(self, frames, ids, frame_idx=None):
        """frames u8 [H,W,3] or [F,H,W,3] with frame_idx per track"""
        frames = np.ascontiguousarray(frames, np.uint8)
        if frames.ndim == 3:
            frames = frames[None]
        ids = _i(ids)
        fi = _i(frame_idx) if frame_idx is not None else _i(np.zeros(len(ids)))
        psr = np.empty(len(ids), np.float64)
        lib().pvc_trackers_update(self.h, _p(frames, C.c_uint8), frames.shape[1], frames.shape[2], _p(fi, C.c_int),
                                  _p(ids, C.c_int), len(ids), _p(psr, C.c_double))
        return psr

    def position(self, i):
        out = np.empty(4, np.float64)
        lib().pvc_trackers_position(self.h, int(i), _p(out, C.c_double))
        return tuple(out.tolist())

    def __del__(self):
        try:
            lib().pvc_trackers_destroy(self.h)
        except Exception:
            pass
