"""ORACLE - TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu legs).

ctypes front end of the native checkers for the reference's "fast inference" grouping
(nano_demo/fast_utils/parse/{find_peaks,assign}.cpp, called from nano_demo/fast_utils/group.py:38-47):

  which="port"  oracle/native/fast_utils_port.c - the C restatement (defined up to 32 candidates/persons)
  which="ref"   the reference's own .cpp files compiled where they lie (only where oracle/_ref/libfastutils_ref.so
                exists; defined while counts and persons stay <= 10, the size of the reference's stack arrays)
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_c_i = ctypes.c_int
_c_f = ctypes.c_float
_p = ctypes.c_void_p


def available(which):
    return os.path.exists(os.path.join(_HERE, "_ref", "libfastutils_%s.so" % which))


def _lib(which):
    if which not in _LIBS:
        path = os.path.join(_HERE, "_ref", "libfastutils_%s.so" % which)
        if not os.path.exists(path):
            from .native import build_native
            build_native.build()
        lib = ctypes.CDLL(path)
        pre = "port" if which == "port" else "ref"
        fp = getattr(lib, pre + "_find_peaks_nchw")
        fp.argtypes = [_p, _p, _p, _p, _p, _p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_f, _c_i]
        fp.restype = None
        asg = getattr(lib, pre + "_assign")
        asg.argtypes = [_p, _p, _p, _p, _p, _p, _p, _c_i, _c_i, _c_f]
        asg.restype = _c_i if which == "port" else None
        _LIBS[which] = (fp, asg)
    return _LIBS[which]


def find_peaks(det, tmap, threshold, window_size, max_count, which="port"):
    """det, tmap: float32 [N,C,H,W] -> (count [N,C] i32, val [N,C,M] f32, tag [N,C,M] f32, ind [N,C,M,2] i32),
    outputs zero initialised as in plugins.cpp:52-56."""
    det = np.ascontiguousarray(det, np.float32)
    tmap = np.ascontiguousarray(tmap, np.float32)
    n, c, h, w = det.shape
    m = int(max_count)
    count = np.zeros((n, c), np.int32)
    val = np.zeros((n, c, m), np.float32)
    tag = np.zeros((n, c, m), np.float32)
    ind = np.zeros((n, c, m, 2), np.int32)
    _lib(which)[0](count.ctypes.data, val.ctypes.data, tag.ctypes.data, ind.ctypes.data, det.ctypes.data,
                   tmap.ctypes.data, n, c, h, w, m, float(threshold), int(window_size))
    return count, val, tag, ind


def assign(count, val, tag, ind, joint_order, threshold, max_count, which="port"):
    """One image: count [C], val/tag [C,M], ind [C,M,2] -> (num, ans [M,C,4] f32 zero initialised, status)."""
    count = np.ascontiguousarray(count, np.int32)
    val = np.ascontiguousarray(val, np.float32)
    tag = np.ascontiguousarray(tag, np.float32)
    ind = np.ascontiguousarray(ind, np.int32)
    jo = np.ascontiguousarray(joint_order, np.int32)
    c, m = val.shape[0], int(max_count)
    assert val.shape[1] == m
    if which == "ref":
        assert count.max(initial=0) <= 10, "the reference's arrays hold 10 entries"
    num = np.zeros(1, np.int32)
    ans = np.zeros((m, c, 4), np.float32)
    st = _lib(which)[1](num.ctypes.data, ans.ctypes.data, count.ctypes.data, val.ctypes.data, tag.ctypes.data,
                        ind.ctypes.data, jo.ctypes.data, c, m, float(threshold))
    return int(num[0]), ans, int(st or 0)


def parse(det, tmap, params, which="port"):
    """nano_demo/fast_utils/group.py:38-47 for a batch: returns a list over images of ans[:num] (scale 1)."""
    tm = np.asarray(tmap)
    if tm.ndim == 5:
        tm = tm[..., 0]
    count, val, tag, ind = find_peaks(det, tm, params["detection_threshold"], params["window_size"],
                                      params["max_num_people"], which)
    out = []
    for i in range(count.shape[0]):
        num, ans, st = assign(count[i], val[i], tag[i], ind[i], params["joint_order"], params["tag_threshold"],
                              params["max_num_people"], which)
        out.append((num, ans, st))
    return out
