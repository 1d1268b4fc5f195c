"""ctypes binding of libpvb200.so (the sm_100a kernels behind this package).

There is deliberately no CPU fallback: if the shared library is missing or a launch fails,
the caller gets an exception (`PvError`).  The ABI is declared in include/pv_b200.h.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvb200.so")

PV_SR_MAX_TAPS = 9
PV_SR_MAX_ENTRIES = 192


class PvError(RuntimeError):
    pass


class PvSrEntry(C.Structure):
    _fields_ = [
        ("a_row_off", C.c_int32),
        ("b_row", C.c_int32),
        ("a_smem_off", C.c_int32),
        ("b_smem_off", C.c_int32),
        ("slot_tx_bytes", C.c_uint32),
        ("a_col", C.c_int16),
        ("cls", C.c_int16),
        ("n_taps", C.c_int16),
        ("use_tail", C.c_int16),
        ("flags", C.c_int16),
        ("tap_rel", C.c_int16 * (PV_SR_MAX_TAPS + 2)),
    ]


class PvRowMap(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("cols", C.c_int32),
        ("w", C.c_int32),
        ("py", C.c_int32),
        ("px", C.c_int32),
        ("img", C.c_int64),
        ("plane_rows", C.c_int64),
    ]


class PvSrgemmDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("x_rows", C.c_int64),
        ("x_cols", C.c_int32),
        ("n_out", C.c_int32),
        ("n_classes", C.c_int32),
        ("class_width", C.c_int32 * 2),
        ("w_packed", C.c_void_p * 2),
        ("w_rows", C.c_int64 * 2),
        ("tail_rows", C.c_int32),
        ("n_entries", C.c_int32),
        ("entries", C.POINTER(PvSrEntry)),
        ("x_row_stride_bytes", C.c_int64),
        ("weights_resident_max_bytes", C.c_int64),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("hq", C.c_int32),
        ("wq", C.c_int32),
        ("oh", C.c_int32),
        ("ow", C.c_int32),
        ("relu", C.c_int32),
        ("out_mode", C.c_int32),
        ("out", C.c_void_p),
        ("dst", PvRowMap),
        ("resid", C.c_void_p),
        ("res", PvRowMap),
        ("max_ctas", C.c_int32),
        ("acc_split", C.c_int32),
        ("ctas_per_sm", C.c_int32),
        ("mma_warps", C.c_int32),
    ]


class PvDetconvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("pitch", C.c_int32),
        ("c_in", C.c_int32), ("n_out", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32), ("out_f32", C.c_int32),
        ("w_img", C.c_void_p),
        ("w_bytes", C.c_int64),
        ("scale", C.c_void_p),
        ("shift", C.c_void_p),
        ("relu", C.c_int32),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int32), ("out_cs", C.c_int32),
        ("resid", C.c_void_p),
        ("gap_period", C.c_int32), ("gap_pos", C.c_int32),
    ]


class PvC12Desc(C.Structure):
    _fields_ = [
        ("plane", C.c_void_p),
        ("B", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32),
        ("w1_img", C.c_void_p), ("w1_bytes", C.c_int64),
        ("w2_img", C.c_void_p), ("w2_bytes", C.c_int64),
        ("scale1", C.c_void_p), ("shift1", C.c_void_p),
        ("scale2", C.c_void_p), ("shift2", C.c_void_p),
        ("out", C.c_void_p),
        ("out_pitch", C.c_int32),
        ("mean_host", C.POINTER(C.c_float)),
    ]


PV_HOG_MAX_LEVELS = 24


class PvHogLevel(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("x0", "y0", "w", "h", "cx", "cy", "fx0", "fy0", "px_off", "cell_off", "feat_off")]


class PvHogGeo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_levels", "Hp", "Wp", "total_px", "total_cells", "total_feat", "FH", "FW", "fpitch")] \
        + [("lv", PvHogLevel * PV_HOG_MAX_LEVELS)]


_lib = None


def lib():
    """Load libpvb200.so (once).  Raises PvError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PvError(
            "libpvb200.so not found at %s — run `python -c 'import __graft_entry__ as g; g.build()'`"
            % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    l.pv_last_error.restype = C.c_char_p
    l.pv_launch_count.restype = C.c_int64
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        msg = lib().pv_last_error().decode("utf-8", "replace")
        raise PvError("%s failed (rc=%d): %s" % (what or "pv call", rc, msg))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def launch_count():
    return int(lib().pv_launch_count())
