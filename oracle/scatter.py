"""ctypes front end of the C oracle (oracle/sbs_scatter.c).  TEST INFRASTRUCTURE ONLY.

Tables (bounds, offsets) come from oracle/sbs_layered.py's python-double restatement of
PredictAndGenerate.py:101-126; this module only runs the per-pixel work in C so that 1080p / 4K
parity cases finish in seconds.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import sbs_layered as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsbs_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "sbs_scatter.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        l = ctypes.CDLL(_SO)
        u16p = ctypes.POINTER(ctypes.c_uint16)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        l.sbs_oracle_smooth_f16.argtypes = [u16p, u16p, u16p, u16p, ctypes.c_size_t,
                                            ctypes.c_float, ctypes.c_float, ctypes.c_float]
        l.sbs_oracle_smooth_f16.restype = None
        l.sbs_oracle_max_f16.argtypes = [u16p, ctypes.c_size_t]
        l.sbs_oracle_max_f16.restype = ctypes.c_float
        l.sbs_oracle_warp_frame.argtypes = [u8p, u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            u16p, u16p, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_float), u8p,
                                            ctypes.POINTER(ctypes.c_int16), u8p]
        l.sbs_oracle_warp_frame.restype = ctypes.c_long
        f32p = ctypes.POINTER(ctypes.c_float)
        l.sbs_oracle_smooth_f32.argtypes = [f32p, f32p, f32p, f32p, ctypes.c_size_t, ctypes.c_float, ctypes.c_float, ctypes.c_float]
        l.sbs_oracle_smooth_f32.restype = None
        l.sbs_oracle_max_f32.argtypes = [f32p, ctypes.c_size_t]
        l.sbs_oracle_max_f32.restype = ctypes.c_float
        l.sbs_oracle_warp_frame_f32.argtypes = [u8p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, f32p,
                                                ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, f32p, u8p, ctypes.POINTER(ctypes.c_int16), u8p]
        l.sbs_oracle_warp_frame_f32.restype = ctypes.c_long
        _lib = l
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def smooth_depth(state, raw):
    """Same contract as sbs_layered.smooth_depth (fp16 or fp32 depth)."""
    assert raw.dtype in (np.float16, np.float32)
    raw = np.ascontiguousarray(raw)
    hist = state.history
    while len(hist) < 2:
        hist.append(raw.copy())
    out = np.empty_like(raw)
    if raw.dtype == np.float32:
        lib().sbs_oracle_smooth_f32(_p(raw, ctypes.c_float), _p(hist[1], ctypes.c_float), _p(hist[0], ctypes.c_float),
                                    _p(out, ctypes.c_float), raw.size, np.float32(state.w_now), np.float32(state.taps[0]),
                                    np.float32(state.taps[1]))
        del hist[0]
        hist.append(raw.copy())
        return out
    lib().sbs_oracle_smooth_f16(_p(raw.view(np.uint16), ctypes.c_uint16),
                                _p(hist[1].view(np.uint16), ctypes.c_uint16),
                                _p(hist[0].view(np.uint16), ctypes.c_uint16),
                                _p(out.view(np.uint16), ctypes.c_uint16), raw.size,
                                np.float32(state.w_now), np.float32(state.taps[0]),
                                np.float32(state.taps[1]))
    del hist[0]
    hist.append(raw.copy())
    return out


def depth_max(depth):
    depth = np.ascontiguousarray(depth)
    if depth.dtype == np.float32:
        return float(lib().sbs_oracle_max_f32(_p(depth, ctypes.c_float), depth.size))
    return float(lib().sbs_oracle_max_f16(_p(depth.view(np.uint16), ctypes.c_uint16), depth.size))


def warp_frame(img, depth, marks, steps, offsets, weights=None, stages=None):
    H, W, _ = img.shape
    n = len(steps)
    f32 = depth.dtype == np.float32
    lo, hi = L.layer_bounds(marks, steps, np.float32 if f32 else np.float16)
    off = np.asarray(offsets, dtype=np.int32)
    kx, ky = L.blur_kernel_shape(H)
    if weights is None:
        weights = L.gaussian_weights(kx, ky)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    ky, kx = weights.shape                      # caller-supplied kernels keep their own shape
    img = np.ascontiguousarray(img)
    depth = np.ascontiguousarray(depth)
    sbs = np.empty((H, 2 * W, 3), dtype=np.uint8)
    winner = np.empty((H, W), dtype=np.int16)
    pre = np.empty((H, W, 3), dtype=np.uint8)
    fill = int(n * 3 / 5)
    sw = L.strip_width(offsets[-1], W)
    if f32:
        holes = lib().sbs_oracle_warp_frame_f32(
            _p(img, ctypes.c_uint8), _p(depth, ctypes.c_float), H, W, n, _p(lo, ctypes.c_float), _p(hi, ctypes.c_float),
            _p(off, ctypes.c_int), fill, sw, kx, ky, _p(weights, ctypes.c_float),
            _p(sbs, ctypes.c_uint8), _p(winner, ctypes.c_int16), _p(pre, ctypes.c_uint8))
    else:
        holes = lib().sbs_oracle_warp_frame(
            _p(img, ctypes.c_uint8), _p(depth.view(np.uint16), ctypes.c_uint16), H, W, n,
            _p(lo.view(np.uint16), ctypes.c_uint16), _p(hi.view(np.uint16), ctypes.c_uint16),
            _p(off, ctypes.c_int), fill, sw, kx, ky, _p(weights, ctypes.c_float),
            _p(sbs, ctypes.c_uint8), _p(winner, ctypes.c_int16), _p(pre, ctypes.c_uint8))
    if holes < 0:
        raise ValueError("sbs_oracle_warp_frame: bad arguments")
    if stages is not None:
        stages.update(lo=lo, hi=hi, winner=winner, holes=winner < 0, pre_blur=pre, strip=sw,
                      fill_layer=fill, weights=weights, n_holes=holes)
    return sbs


def process_frame(state, img, raw_depth, weights=None, stages=None):
    depth = smooth_depth(state, raw_depth)
    marks, steps, offsets, limit, rng = L.layer_tables(state, depth_max(depth), depth.shape[0])
    if stages is not None:
        stages.update(depth=depth, marks=marks, steps=steps, offsets=offsets, limit=limit, range=rng)
    return warp_frame(img, depth, marks, steps, offsets, weights, stages)
