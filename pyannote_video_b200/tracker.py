"""Correlation trackers on the GPU.

  TrackerBank          all live trackers of a frame advance in ONE kernel launch (csrc/tracker.cu)
  CorrelationTracker   dlib.correlation_tracker duck type (start_track / update / get_position,
                       pyannote/video/tracking.py:203,231,250-251) on top of a shared bank

Frames are uint8 [H,W,3] tensors resident in HBM (`prepare_frame` uploads a numpy frame once per
shot; the reference keeps the shot's frames in host RAM, tracking.py:361,420).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .geometry import DRect

FILTER_SIZE = 64
PADDING = 1.4
REGULARIZER_SPACE = 0.001
NU_SPACE = 0.025
N_SCALES = 32
SCALE_ALPHA = 1.020
REGULARIZER_SCALE = 0.001
NU_SCALE = 0.025


def _scale_tables():
    f32 = np.float32
    n = N_SCALES
    hann = (f32(0.5) - f32(0.5) * np.cos(2 * np.pi * np.arange(n) / (n - 1))).astype(f32)
    factor = np.asarray([f32(SCALE_ALPHA) ** f32(k - n // 2) for k in range(n)], f32)
    m = np.arange(n)
    return hann, factor, np.cos(-2 * np.pi * m / n).astype(f32), np.sin(-2 * np.pi * m / n).astype(f32)


def _tables():
    f32 = np.float32
    n = FILTER_SIZE
    hann = (f32(0.5) - f32(0.5) * np.cos(2 * np.pi * np.arange(n) / (n - 1))).astype(f32)
    uu = np.cos(np.arange(9) * np.pi / 9).astype(f32)
    vv = np.sin(np.arange(9) * np.pi / 9).astype(f32)
    k = np.arange(32)
    tw_re = np.cos(-2 * np.pi * k / 64).astype(f32)
    tw_im = np.sin(-2 * np.pi * k / 64).astype(f32)
    return hann, uu, vv, tw_re, tw_im


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class TrackerBank(object):
    def __init__(self, capacity=256, device=None, use_scale=True):
        if not torch.cuda.is_available():
            raise RuntimeError("TrackerBank needs a CUDA device; there is no CPU fallback")
        self.dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.capacity = int(capacity)
        self._tables = _tables()
        h = C.c_void_p()
        hann, uu, vv, tr, ti = self._tables
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().pv_tracker_create(self.capacity, _fp(hann), _fp(uu), _fp(vv), _fp(tr), _fp(ti),
                                                    C.c_float(PADDING), C.c_float(REGULARIZER_SPACE), C.c_float(NU_SPACE),
                                                    C.byref(h)), "pv_tracker_create")
        self.h = h
        self.use_scale = bool(use_scale)
        if self.use_scale:
            sh, sf, sr, si = self._scale_tabs = _scale_tables()
            with torch.cuda.device(self.dev):
                _lib.check(_lib.lib().pv_tracker_enable_scale(self.h, _fp(sh), _fp(sf), _fp(sr), _fp(si),
                                                              C.c_double(SCALE_ALPHA), C.c_float(REGULARIZER_SCALE),
                                                              C.c_float(NU_SCALE)), "pv_tracker_enable_scale")
        pos, psr = C.POINTER(C.c_float)(), C.POINTER(C.c_float)()
        _lib.check(_lib.lib().pv_tracker_state(self.h, C.byref(pos), C.byref(psr)), "pv_tracker_state")
        self._pos_ptr, self._psr_ptr = pos, psr
        self._state = torch.empty(self.capacity, 5, dtype=torch.float32, device=self.dev)
        self.reset()

    # ---- bank interface used by TrackingByDetection ----
    def reset(self):
        self._free = list(range(self.capacity - 1, -1, -1))
        self._pending = []          # (slot, rect) started on self._pending_frame
        self._pending_frame = None
        self._cache = {}            # slot -> (l,t,r,b) host copy of positions

    def prepare_frame(self, frame):
        if isinstance(frame, torch.Tensor):
            t = frame
        else:
            t = torch.from_numpy(np.ascontiguousarray(frame))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
            raise RuntimeError("frames must be uint8 [H,W,3]")
        return t.to(self.dev, non_blocking=True).contiguous()

    def start(self, frame, rect):
        if not self._free:
            raise RuntimeError("TrackerBank: capacity %d exhausted (construct the bank with a larger capacity)"
                               % self.capacity)
        frame = self.prepare_frame(frame)
        if self._pending and self._pending_frame.data_ptr() != frame.data_ptr():
            self._flush()
        slot = self._free.pop()
        self._pending.append((slot, (rect.left(), rect.top(), rect.right(), rect.bottom())))
        self._pending_frame = frame
        self._cache[slot] = tuple(float(v) for v in self._pending[-1][1])
        return slot

    def _flush(self):
        if not self._pending:
            return
        ids = torch.tensor([s for s, _ in self._pending], dtype=torch.int32, device=self.dev)
        rects = torch.tensor([r for _, r in self._pending], dtype=torch.float32, device=self.dev)
        f = self._pending_frame
        _lib.check(_lib.lib().pv_tracker_start(self.h, _lib.ptr(f), f.shape[0], f.shape[1], _lib.ptr(ids), _lib.ptr(rects),
                                               len(self._pending), _lib.stream_ptr()), "pv_tracker_start")
        self._keep = (ids, rects, f)
        self._pending, self._pending_frame = [], None

    def update(self, frame, handles):
        """advance `handles` to `frame`; returns their confidences (PSR) as floats"""
        self._flush()
        if not handles:
            return []
        frame = self.prepare_frame(frame)
        ids = torch.tensor(list(handles), dtype=torch.int32, device=self.dev)
        _lib.check(_lib.lib().pv_tracker_update(self.h, _lib.ptr(frame), frame.shape[0], frame.shape[1], _lib.ptr(ids),
                                                len(handles), _lib.stream_ptr()), "pv_tracker_update")
        state = self.read_state(ids).cpu().numpy()
        for s, row in zip(handles, state):
            self._cache[s] = tuple(float(v) for v in row[:4])
        return [float(v) for v in state[:, 4]]

    # ---- batched device-side interface (no host synchronisation): tracks of many shots in one launch ----
    def start_batch(self, frames, frame_idx, ids, rects):
        """frames u8 [F,H,W,3] (device); frame_idx i32 [n], ids i32 [n], rects f32 [n,4] device tensors"""
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
        _lib.check(_lib.lib().pv_tracker_start_frames(self.h, _lib.ptr(frames), frames.shape[1], frames.shape[2],
                                                      _lib.ptr(frame_idx), _lib.ptr(ids), _lib.ptr(rects), int(ids.shape[0]),
                                                      _lib.stream_ptr()), "pv_tracker_start_frames")

    def update_batch(self, frames, frame_idx, ids):
        """advance tracks `ids` (i32 [n], device) to frames[frame_idx[i]]; positions / PSR stay in the bank
        (`read_state`)"""
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
        _lib.check(_lib.lib().pv_tracker_update_frames(self.h, _lib.ptr(frames), frames.shape[1], frames.shape[2],
                                                       _lib.ptr(frame_idx), _lib.ptr(ids), int(ids.shape[0]),
                                                       _lib.stream_ptr()), "pv_tracker_update_frames")

    def read_state(self, ids):
        """device tensor [n,5]: l,t,r,b,psr of the given slots"""
        pos = _from_ptr(self._pos_ptr, (self.capacity, 4), self.dev)
        psr = _from_ptr(self._psr_ptr, (self.capacity,), self.dev)
        idx = ids.long()
        return torch.cat([pos[idx], psr[idx][:, None]], dim=1)

    def position(self, handle):
        return DRect(*self._cache[handle])

    def release(self, handle):
        self._pending = [(s, r) for s, r in self._pending if s != handle]
        self._cache.pop(handle, None)
        self._free.append(handle)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().pv_tracker_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _from_ptr(ptr, shape, device):
    """zero-copy torch view of a device float buffer owned by the library"""
    n = int(np.prod(shape))
    addr = C.cast(ptr, C.c_void_p).value

    class _Holder(object):
        pass
    h = _Holder()
    h.__cuda_array_interface__ = dict(shape=(n,), typestr="<f4", data=(addr, False), version=2)
    return torch.as_tensor(h, device=device).view(*shape)


_shared_bank = None


class CorrelationTracker(object):
    """dlib.correlation_tracker duck type"""

    def __init__(self, bank=None):
        global _shared_bank
        if bank is None:
            if _shared_bank is None:
                _shared_bank = TrackerBank()
            bank = _shared_bank
        self.bank = bank
        self.handle = None

    def start_track(self, frame, rect):
        if self.handle is not None:
            self.bank.release(self.handle)
        self.handle = self.bank.start(frame, rect)

    def update(self, frame):
        return self.bank.update(frame, [self.handle])[0]

    def get_position(self):
        return self.bank.position(self.handle)

    def __del__(self):
        try:
            if self.handle is not None:
                self.bank.release(self.handle)
        except Exception:
            pass
