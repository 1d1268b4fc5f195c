"""dlib `.dat` model files -> this package's model dicts.

The reference loads three dlib models (pyannote/video/face/face.py:57-62, README.md:27-30):
`shape_predictor_68_face_landmarks.dat`, `dlib_face_recognition_resnet_model_v1.dat` and — for the CNN detector
of BASELINE.json's north_star — `mmod_human_face_detector.dat`.  Neither dlib nor any of these files exists in
the build environment, so the stream format below is RESTATED FROM MEMORY of dlib 19.x's `serialize.h`,
`shape_predictor.h`, `dnn/core.h`, `dnn/layers.h`, `dnn/loss.h`, `dnn/input.h` [MEMORY — unverified against a real
file].  What is believed exact: the primitive encodings (ints, floats, vectors, matrices, strings, tensors) and the
whole `shape_predictor` stream.  The DNN streams are architecture-driven (the layer sequence of `anet_type` and of
the MMOD face net is spelled out below) and every layer reader checks the version string it finds, so a mismatch
fails loudly with the byte offset instead of producing garbage weights.  `dump_*` writes the same format and is
used by the round-trip tests; it also gives users a way to inspect what this module expects.

Primitive encodings (dlib/serialize.h):
  integer      one control byte (low nibble = n magnitude bytes, 0x80 = negative) + n little-endian bytes
  float/double `float_details`: int64 mantissa + int16 exponent (value = mantissa * 2^exponent; exponent 32000 /
               32001 / 32002 = +inf / -inf / nan), both as integers above
  bool         one byte '1' / '0'
  std::string  length (integer) + bytes;   std::vector<T>: length + items
  matrix<T>    -nr, -nc (negative marks the current format) + items row-major
  tensor       int version (2), num_samples, k, nr, nc (integers), then raw little-endian float32
"""
import io
import math
import struct

import numpy as np

from . import weights as W


class DatError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------------------------
class Reader(object):
    def __init__(self, data):
        self.b = memoryview(data)
        self.p = 0

    def fail(self, msg):
        raise DatError("dlib .dat: %s at byte %d" % (msg, self.p))

    def int(self):
        if self.p >= len(self.b):
            self.fail("unexpected end of file")
        c = self.b[self.p]
        n = c & 0x0F
        if n == 0 or n > 8 or (c & 0x70):
            self.fail("bad integer control byte 0x%02x" % c)
        v = int.from_bytes(self.b[self.p + 1:self.p + 1 + n], "little")
        self.p += 1 + n
        return -v if (c & 0x80) else v

    def float(self):
        m, e = self.int(), self.int()
        if e == 32000:
            return math.inf
        if e == 32001:
            return -math.inf
        if e == 32002:
            return math.nan
        return math.ldexp(m, e)

    def bool(self):
        if self.p >= len(self.b):
            self.fail("unexpected end of file")
        c = self.b[self.p]
        self.p += 1
        if c not in (0x30, 0x31):
            self.fail("bad bool byte 0x%02x" % c)
        return c == 0x31

    def string(self):
        n = self.int()
        if n < 0 or self.p + n > len(self.b):
            self.fail("string of length %d runs past the end of the file" % n)
        s = bytes(self.b[self.p:self.p + n]).decode("latin-1")
        self.p += n
        return s

    def expect(self, *names):
        s = self.string()
        if s not in names:
            self.fail("expected %s, found %r" % (" / ".join(names), s[:40]))
        return s

    def vector(self, item):
        return [item() for _ in range(self.int())]

    def matrix(self, item=None):
        nr, nc = self.int(), self.int()
        if nr > 0 or nc > 0:
            self.fail("old-format matrix (positive sizes) is not supported")
        nr, nc = -nr, -nc
        item = item or self.float
        return np.asarray([item() for _ in range(nr * nc)], np.float64).reshape(nr, nc)

    def tensor(self):
        v = self.int()
        if v != 2:
            self.fail("tensor version %d (expected 2)" % v)
        shape = [self.int() for _ in range(4)]
        n = int(np.prod(shape))
        if min(shape) < 0 or self.p + 4 * n > len(self.b):
            self.fail("tensor of shape %s runs past the end of the file" % (shape, ))
        a = np.frombuffer(self.b[self.p:self.p + 4 * n], "<f4").reshape(shape).copy()
        self.p += 4 * n
        return a

    def alias_tensor(self):
        v = self.int()
        if v != 1:
            self.fail("alias_tensor version %d" % v)
        return [self.int() for _ in range(4)]


class Writer(object):
    def __init__(self):
        self.o = io.BytesIO()

    def int(self, v):
        v = int(v)
        neg = 0x80 if v < 0 else 0
        v = abs(v)
        raw = v.to_bytes(8, "little").rstrip(b"\x00") or b"\x00"
        self.o.write(bytes([len(raw) | neg]) + raw)

    def float(self, x, digits=24):
        x = float(np.float32(x)) if digits == 24 else float(x)      # a C++ float is serialised: exactly 24 mantissa bits
        if math.isinf(x):
            self.int(0), self.int(32000 if x > 0 else 32001)
            return
        if math.isnan(x):
            self.int(0), self.int(32002)
            return
        fr, ex = math.frexp(x)
        m = int(fr * (1 << digits))
        e = ex - digits
        for _ in range(8):
            if m == 0 or (m & 0xFF):
                break
            m >>= 8
            e += 8
        self.int(m), self.int(e)

    def double(self, x):
        self.float(x, 53)

    def bool(self, v):
        self.o.write(b"1" if v else b"0")

    def string(self, s):
        raw = s.encode("latin-1")
        self.int(len(raw))
        self.o.write(raw)

    def vector(self, items, item):
        self.int(len(items))
        for it in items:
            item(it)

    def matrix(self, a):
        a = np.asarray(a, np.float32)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        self.int(-a.shape[0]), self.int(-a.shape[1])
        for v in a.reshape(-1):
            self.float(v)

    def tensor(self, a):
        a = np.ascontiguousarray(a, "<f4")
        shape = list(a.shape) + [1] * (4 - a.ndim) if a.size else [0, 0, 0, 0]
        self.int(2)
        for s in shape:
            self.int(s)
        self.o.write(a.tobytes())

    def alias_tensor(self, shape):
        self.int(1)
        for s in list(shape) + [1] * (4 - len(shape)):
            self.int(s)

    def bytes(self):
        return self.o.getvalue()


# ------------------------------------------------------------------------------------------------------------------
# shape_predictor (dlib/image_processing/shape_predictor.h)
# ------------------------------------------------------------------------------------------------------------------
def read_shape_predictor(r):
    if r.int() != 1:
        r.fail("shape_predictor version")
    initial = r.matrix().reshape(-1)

    def tree():
        splits = r.vector(lambda: (r.int(), r.int(), r.float()))
        leaves = r.vector(lambda: r.matrix().reshape(-1))
        return splits, leaves

    forests = r.vector(lambda: r.vector(tree))
    anchor = r.vector(lambda: r.vector(r.int))
    deltas = r.vector(lambda: r.vector(lambda: (r.float(), r.float())))
    S, T = len(forests), len(forests[0])
    n_split = len(forests[0][0][0])
    n_leaf = len(forests[0][0][1])
    if n_split + 1 != n_leaf:
        r.fail("regression tree with %d splits and %d leaves" % (n_split, n_leaf))
    P2 = initial.shape[0]
    m = dict(kind="ert_shape_predictor", initial_shape=initial.astype(np.float32),
             anchor_idx=np.asarray(anchor, np.int32), deltas=np.asarray(deltas, np.float32),
             split_idx1=np.zeros((S, T, n_split), np.int32), split_idx2=np.zeros((S, T, n_split), np.int32),
             split_thresh=np.zeros((S, T, n_split), np.float32), leaf_values=np.zeros((S, T, n_leaf, P2), np.float32))
    for s in range(S):
        for t in range(T):
            splits, leaves = forests[s][t]
            for k, (i1, i2, th) in enumerate(splits):
                m["split_idx1"][s, t, k], m["split_idx2"][s, t, k], m["split_thresh"][s, t, k] = i1, i2, th
            m["leaf_values"][s, t] = np.asarray(leaves, np.float32)
    return m


def dump_shape_predictor(m):
    w = Writer()
    w.int(1)
    w.matrix(m["initial_shape"])
    S, T, n_split = m["split_thresh"].shape
    w.int(S)
    for s in range(S):
        w.int(T)
        for t in range(T):
            w.int(n_split)
            for k in range(n_split):
                w.int(m["split_idx1"][s, t, k]), w.int(m["split_idx2"][s, t, k]), w.float(m["split_thresh"][s, t, k])
            w.vector(list(m["leaf_values"][s, t]), w.matrix)
    w.vector(list(m["anchor_idx"]), lambda row: w.vector(list(row), w.int))
    w.vector(list(m["deltas"]), lambda row: w.vector(list(row), lambda d: (w.float(d[0]), w.float(d[1]))))
    return w.bytes()


# ------------------------------------------------------------------------------------------------------------------
# DNN streams (dlib/dnn/core.h): add_layer writes  version, SUBNETWORK, details, 3 bools, x_grad, cached_output,
# params_grad  — so all wrapper versions come first (outermost to innermost), then the input layer, then the layer
# details in execution order.  Tag / skip layers write only their version.
# ------------------------------------------------------------------------------------------------------------------
def _block(n, stride):
    # block<N,BN,stride,SUBNET> = BN<con<N,3,3,1,1,relu<BN<con<N,3,3,stride,stride,SUBNET>>>>>   (outermost first)
    return [("affine", ), ("con", n, 3, 1), ("relu", ), ("affine", ), ("con", n, 3, stride)]


def _ares(n):
    return [("relu", ), ("add_prev", )] + _block(n, 1) + [("tag", )]


def _ares_down(n):
    return [("relu", ), ("add_prev", ), ("avg_pool", ), ("skip", ), ("tag", )] + _block(n, 2) + [("tag", )]


def anet_arch():
    """face_recognition_resnet_model_v1's anet_type, outermost layer first (dlib examples/dnn_face_recognition_ex.cpp)"""
    a = [("fc_no_bias", 128), ("avg_pool_everything", )]
    a += _ares_down(256)                                        # alevel0
    a += _ares(256) + _ares(256) + _ares_down(256)              # alevel1
    a += _ares(128) + _ares(128) + _ares_down(128)              # alevel2
    a += _ares(64) + _ares(64) + _ares(64) + _ares_down(64)     # alevel3
    a += _ares(32) + _ares(32) + _ares(32)                      # alevel4
    a += [("max_pool", ), ("relu", ), ("affine", ), ("con", 32, 7, 2)]
    return a


def mmod_arch():
    """mmod_human_face_detector's net_type (dlib examples/dnn_mmod_face_detection_ex.cpp), outermost first"""
    a = [("con", 1, 9, 1)]
    for _ in range(3):
        a += [("relu", ), ("affine", ), ("con", 45, 5, 1)]
    for n in (32, 32, 16):
        a += [("relu", ), ("affine", ), ("con", n, 5, 2)]
    return a


def _read_layer_details(r, spec):
    kind = spec[0]
    if kind == "con":
        r.expect("con_4", "con_5")
        params = r.tensor().reshape(-1)
        nf, nr, nc, sy, sx, py, px = [r.int() for _ in range(7)]
        fs, bs = r.alias_tensor(), r.alias_tensor()
        for _ in range(4):
            r.float()
        n_w = int(np.prod(fs))
        return dict(w=params[:n_w].reshape(fs), b=params[n_w:n_w + int(np.prod(bs))].copy(), stride=sy, pad=py)
    if kind == "affine":
        v = r.expect("affine_", "bn_con2")
        params = r.tensor().reshape(-1)
        g, b = r.alias_tensor(), r.alias_tensor()
        n = int(np.prod(g))
        gamma, beta = params[:n].copy(), params[n:2 * n].copy()
        if v == "affine_":
            r.int()                                            # mode
        else:                                                  # a bn_ layer saved from training: fold the running statistics
            r.tensor(), r.tensor()                             # means, invstds
            rm, rv = r.tensor().reshape(-1), r.tensor().reshape(-1)
            r.int(), r.int()                                   # num_updates, running_stats_window_size
            for _ in range(4):
                r.float()
            eps = r.float()
            gamma = gamma / np.sqrt(rv + np.float32(eps))
            beta = beta - gamma * rm
        return dict(gamma=gamma, beta=beta)
    if kind == "relu":
        r.expect("relu_")
    elif kind == "add_prev":
        r.expect("add_prev_")
    elif kind in ("max_pool", "avg_pool", "avg_pool_everything"):
        r.expect("max_pool_2" if kind == "max_pool" else "avg_pool_2")
        return dict(geom=[r.int() for _ in range(6)])
    elif kind == "fc_no_bias":
        r.expect("fc_2")
        n_out, n_in = r.int(), r.int()
        params = r.tensor().reshape(-1)
        r.alias_tensor(), r.alias_tensor()
        r.int()
        for _ in range(4):
            r.float()
        return dict(w=params[:n_in * n_out].reshape(n_in, n_out))
    return {}


def _read_net(r, arch, read_input):
    for spec in arch:                                          # wrapper versions, outermost first
        v = r.int()
        if spec[0] in ("tag", "skip"):
            if v != 1:
                r.fail("tag/skip layer version %d" % v)
        elif v not in (2, 3):
            r.fail("add_layer version %d" % v)
    inp = read_input(r)
    layers = []
    for i, spec in enumerate(reversed(arch)):                  # details, innermost (first executed) first
        if spec[0] in ("tag", "skip"):
            continue
        d = _read_layer_details(r, spec)
        r.bool(), r.bool(), r.bool()
        r.tensor(), r.tensor(), r.tensor()
        layers.append((spec, d))
    return inp, layers


def _conv_dict(con, aff):
    cout = con["w"].shape[0]
    one, zero = np.ones(cout, np.float32), np.zeros(cout, np.float32)
    return dict(w=con["w"].astype(np.float32), b=con["b"].astype(np.float32),
                gamma=(aff["gamma"] if aff else one).astype(np.float32), beta=(aff["beta"] if aff else zero).astype(np.float32))


def read_embedder(r):
    if r.int() != 1:
        r.fail("add_loss_layer version")
    v = r.expect("loss_metric_", "loss_metric_2")
    if v == "loss_metric_2":
        r.float(), r.float()

    def read_input(r):
        r.expect("input_rgb_image_sized")
        avg = [r.float() for _ in range(3)]
        r.int(), r.int()
        return avg
    avg, layers = _read_net(r, anet_arch(), read_input)
    if not np.allclose(avg, W.PIXEL_MEAN, atol=1e-3):
        r.fail("unexpected input averages %s" % (avg, ))
    convs = []
    it = iter(layers)
    fc = None
    pend = None
    for spec, d in it:
        if spec[0] == "con":
            pend = d
        elif spec[0] == "affine":
            convs.append(_conv_dict(pend, d))
        elif spec[0] == "fc_no_bias":
            fc = d["w"].T.copy()
    blocks = []
    types = [t for t, _, _ in W.embed_block_list()]
    for i, t in enumerate(types):
        blocks.append(dict(type=t, a=convs[1 + 2 * i], b=convs[2 + 2 * i]))
    return dict(kind="resnet_v1_embedder", conv1=convs[0], blocks=blocks, fc=fc.astype(np.float32))


def read_detector(r):
    if r.int() != 1:
        r.fail("add_loss_layer version")
    r.expect("loss_mmod_", "loss_mmod_2", "loss_mmod_3")
    # mmod_options
    windows = r.vector(lambda: (r.int(), r.int(), r.string()))
    r.float(), r.float(), r.float()                            # loss_per_false_alarm, loss_per_missed_target, truth_match_iou_threshold
    nms = (r.float(), r.float())                               # overlaps_nms: iou_thresh, percent_covered_thresh
    r.float(), r.float()                                       # overlaps_ignore

    def read_input(r):
        r.expect("input_rgb_image_pyramid", "input_rgb_image_pyramid2")
        avg = [r.float() for _ in range(3)]
        r.int(), r.int()                                       # pyramid_padding, pyramid_outer_padding
        return avg
    avg, layers = _read_net(r, mmod_arch(), read_input)
    convs, pend = [], None
    for spec, d in layers:
        if spec[0] == "con":
            if pend is not None:
                convs.append(_conv_dict(pend, None))
            pend = d
        elif spec[0] == "affine":
            convs.append(_conv_dict(pend, d))
            pend = None
    if pend is not None:
        convs.append(_conv_dict(pend, None))
    win = max(windows[0][0], windows[0][1]) if windows else W.DET_WINDOW
    return dict(kind="mmod_detector", convs=convs, window=int(win), iou_thresh=float(nms[0]), covered_thresh=float(nms[1]),
                adjust_threshold=0.0)


# ---- writers (same streams; used by the round-trip tests) ----
def _dump_layer_details(w, spec, d):
    kind = spec[0]
    if kind == "con":
        w.string("con_4")
        wt, b = np.asarray(d["w"], np.float32), np.asarray(d["b"], np.float32)
        w.tensor(np.concatenate([wt.reshape(-1), b.reshape(-1)]))
        for v in (wt.shape[0], wt.shape[2], wt.shape[3], spec[3], spec[3], W.conv_pad(spec[2], spec[3]), W.conv_pad(spec[2], spec[3])):
            w.int(v)
        w.alias_tensor(wt.shape), w.alias_tensor([1, wt.shape[0]])
        for _ in range(4):
            w.double(1.0)
    elif kind == "affine":
        w.string("affine_")
        g, b = np.asarray(d["gamma"], np.float32), np.asarray(d["beta"], np.float32)
        w.tensor(np.concatenate([g, b]))
        w.alias_tensor([1, g.shape[0]]), w.alias_tensor([1, g.shape[0]])
        w.int(0)
    elif kind == "relu":
        w.string("relu_")
    elif kind == "add_prev":
        w.string("add_prev_")
    elif kind in ("max_pool", "avg_pool", "avg_pool_everything"):
        w.string("max_pool_2" if kind == "max_pool" else "avg_pool_2")
        geom = dict(max_pool=[3, 3, 2, 2, 0, 0], avg_pool=[2, 2, 2, 2, 0, 0], avg_pool_everything=[0, 0, 1, 1, 0, 0])[kind]
        for v in geom:
            w.int(v)
    elif kind == "fc_no_bias":
        w.string("fc_2")
        fc = np.asarray(d["fc"], np.float32)                    # [out, in]
        w.int(fc.shape[0]), w.int(fc.shape[1])
        w.tensor(fc.T.reshape(-1))
        w.alias_tensor([fc.shape[1], fc.shape[0]]), w.alias_tensor([0, 0])
        w.int(1)
        for _ in range(4):
            w.double(1.0)


def _dump_net(w, arch, dump_input, details):
    for spec in arch:
        w.int(1 if spec[0] in ("tag", "skip") else 2)
    dump_input(w)
    it = iter(details)
    for spec in reversed(arch):
        if spec[0] in ("tag", "skip"):
            continue
        _dump_layer_details(w, spec, next(it) if spec[0] in ("con", "affine", "fc_no_bias") else None)
        w.bool(True), w.bool(False), w.bool(False)
        w.tensor(np.zeros(0, np.float32)), w.tensor(np.zeros(0, np.float32)), w.tensor(np.zeros(0, np.float32))


def dump_embedder(m):
    w = Writer()
    w.int(1)
    w.string("loss_metric_2")
    w.float(0.04), w.float(0.6)

    def dump_input(w):
        w.string("input_rgb_image_sized")
        for v in W.PIXEL_MEAN:
            w.float(v)
        w.int(W.EMB_CHIP), w.int(W.EMB_CHIP)
    det = []
    for c in [m["conv1"]] + [c for blk in m["blocks"] for c in (blk["a"], blk["b"])]:
        det += [dict(w=c["w"], b=c["b"]), dict(gamma=c["gamma"], beta=c["beta"])]
    det.append(dict(fc=m["fc"]))
    _dump_net(w, anet_arch(), dump_input, det)
    return w.bytes()


def dump_detector(m):
    w = Writer()
    w.int(1)
    w.string("loss_mmod_2")
    w.vector([(m["window"], m["window"], "")], lambda t: (w.int(t[0]), w.int(t[1]), w.string(t[2])))
    w.double(1.0), w.double(1.0), w.double(0.5)
    w.double(m["iou_thresh"]), w.double(m["covered_thresh"])
    w.double(0.5), w.double(1.0)

    def dump_input(w):
        w.string("input_rgb_image_pyramid2")
        for v in W.PIXEL_MEAN:
            w.float(v)
        w.int(10), w.int(11)
    det = []
    n = len(m["convs"])
    for i, c in enumerate(m["convs"]):
        det.append(dict(w=c["w"], b=c["b"]))
        if i < n - 1:
            det.append(dict(gamma=c["gamma"], beta=c["beta"]))
    _dump_net(w, mmod_arch(), dump_input, det)
    return w.bytes()


_READERS = dict(ert_shape_predictor=read_shape_predictor, resnet_v1_embedder=read_embedder, mmod_detector=read_detector)
_WRITERS = dict(ert_shape_predictor=dump_shape_predictor, resnet_v1_embedder=dump_embedder, mmod_detector=dump_detector)


def load(path, kind):
    """read a dlib `.dat` model of the given kind ('ert_shape_predictor', 'resnet_v1_embedder', 'mmod_detector')"""
    if kind not in _READERS:
        raise DatError("unknown model kind %r" % kind)
    with open(path, "rb") as f:
        data = f.read()
    if data[:3] == b"BZh":
        raise DatError("%s is bzip2-compressed: decompress it first (dlib ships .dat.bz2)" % path)
    return _READERS[kind](Reader(data))


def loads(data, kind):
    return _READERS[kind](Reader(data))


def dumps(model):
    return _WRITERS[model["kind"]](model)
