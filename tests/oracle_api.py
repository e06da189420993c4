"""ctypes signatures + numpy wrappers for the CPU oracle (oracle/libicg_oracle.so).  Tests/bench only."""
import ctypes as C

import numpy as np

vp = C.c_void_p


def _p(a):
    return C.c_void_p(a.ctypes.data)


def declare(lib):
    lib.icgo_pyr_down.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    lib.icgo_pyr_down.restype = None
    lib.icgo_calc_optical_flow_pyr_lk.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    lib.icgo_calc_optical_flow_pyr_lk.restype = C.c_int
    lib.icgo_track_fb.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_double, C.c_double, C.c_double]
    lib.icgo_track_fb.restype = None


def pyr_down(lib, img):
    H, W = img.shape
    out = np.zeros(((H + 1) // 2, (W + 1) // 2), np.uint8)
    img = np.ascontiguousarray(img)
    lib.icgo_pyr_down(_p(img), W, H, W, _p(out), out.shape[1])
    return out


def lk(lib, a, b, p, init, max_level=3, max_iter=30, eps=0.01, flags=4, want_err=True):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    H, W = a.shape
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 2)
    q = np.array(init, np.float32).reshape(-1, 2).copy()
    n = p.shape[0]
    st = np.zeros(n, np.uint8)
    err = np.zeros(n, np.float32)
    lib.icgo_calc_optical_flow_pyr_lk(_p(a), _p(b), W, H, W, _p(p), _p(q), _p(st), _p(err) if want_err else None, n, 21,
                                      max_level, max_iter, eps, flags, 1e-4)
    return q, st, err


def track_fb(lib, a, b, p, init):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    H, W = a.shape
    p = np.ascontiguousarray(p, np.float32).reshape(-1, 2)
    q = np.array(init, np.float32).reshape(-1, 2).copy()
    n = p.shape[0]
    back = np.zeros((n, 2), np.float32)
    st = np.zeros(n, np.uint8)
    lib.icgo_track_fb(_p(a), _p(b), W, H, W, _p(p), _p(q), _p(back), _p(st), n, 21, 3, 30, 0.01, 0.5, 5.0)
    return q, back, st
