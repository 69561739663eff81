"""Oracle (test infrastructure): ctypes binding of oracle/c/hydro_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhydro_oracle.so")
_lib = None


def build():
    """Compile the C oracle (idempotent)."""
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.ho_majority.argtypes = [p, p, i64, i64, ctypes.c_int, ctypes.c_int]
        _lib.ho_expand.argtypes = [p, p, i64, i64, ctypes.c_int]
        _lib.ho_median.argtypes = [p, p, i64, i64, ctypes.c_int, ctypes.c_int]
        _lib.ho_priority_flood.argtypes = [p, p, i64, i64]
        _lib.ho_d8.argtypes = [p, p, i64, i64]
        _lib.ho_route_rivers.argtypes = [p, p, p, i64, i64]
        _lib.ho_route_rivers.restype = None
        for f in (_lib.ho_majority, _lib.ho_expand, _lib.ho_median, _lib.ho_priority_flood, _lib.ho_d8):
            f.restype = None
    return _lib


def _f32c(a):
    return np.ascontiguousarray(np.asarray(a).astype('float32'))


def majority(image, ws, min_count):
    g = _f32c(image)
    out = np.zeros(g.shape)
    lib().ho_majority(g.ctypes.data, out.ctypes.data, g.shape[0], g.shape[1], ws, min_count)
    return out


def expand(image, ws):
    g = _f32c(image)
    out = np.zeros(g.shape)
    lib().ho_expand(g.ctypes.data, out.ctypes.data, g.shape[0], g.shape[1], ws)
    return out


def median(image, ws, circular=False):
    g = _f32c(image)
    out = np.empty_like(g)
    lib().ho_median(g.ctypes.data, out.ctypes.data, g.shape[0], g.shape[1], ws, int(circular))
    return out


def priority_flood(z):
    g = _f32c(z)
    out = np.empty_like(g)
    lib().ho_priority_flood(g.ctypes.data, out.ctypes.data, g.shape[0], g.shape[1])
    return out


def d8(w):
    g = _f32c(w)
    out = np.zeros(g.shape, dtype=np.uint8)
    lib().ho_d8(g.ctypes.data, out.ctypes.data, g.shape[0], g.shape[1])
    return out


def route_rivers(mask, dem):
    """RouteRivers(window_size=3, dem=dem).apply(mask) (custom_filters.py:165-199): float64 zeros / ones."""
    m, g = _f32c(mask), _f32c(dem).copy()
    out = np.zeros(m.shape, dtype=np.uint8)
    lib().ho_route_rivers(m.ctypes.data, g.ctypes.data, out.ctypes.data, m.shape[0], m.shape[1])
    return out.astype(np.float64)
