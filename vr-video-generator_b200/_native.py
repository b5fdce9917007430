"""ctypes binding of the C ABI in include/vrsbs.h (libvrsbs.so, built in-tree by build.py).

There is no fallback: if the shared library is missing or no sm_100 device is present the import
of the product path fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvrsbs.so")

ABI_VERSION = 2

# every symbol include/vrsbs.h declares (tests check that the library exports exactly these)
SYMBOLS = [
    "vrsbs_abi_version", "vrsbs_create", "vrsbs_destroy", "vrsbs_last_error", "vrsbs_reset",
    "vrsbs_get_range_state", "vrsbs_set_range_state", "vrsbs_set_blur_weights",
    "vrsbs_depth_from_lowres", "vrsbs_depth_from_full", "vrsbs_build_tables", "vrsbs_warp_batch",
    "vrsbs_process_batch", "vrsbs_process_host", "vrsbs_submit_host", "vrsbs_collect", "vrsbs_host_depends_on",
    "vrsbs_get_frame_info", "vrsbs_get_tables", "vrsbs_get_bounds",
    "vrsbs_get_hole_mask", "vrsbs_get_stage_times", "vrsbs_launch_count", "vrsbs_set_option",
]

FRAME_NAN, FRAME_OVERFLOW, FRAME_GENERIC = 1, 2, 4
DEPTH_F16, DEPTH_F32 = 0, 1
HOST_RIGHT_IN_PLACE = 1


class VrsbsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"vrsbs error {code}: {message}")
        self.code = code


class Params(ctypes.Structure):
    _fields_ = [("offset_fg", ctypes.c_double), ("offset_bg", ctypes.c_double),
                ("offset_step_size", ctypes.c_int), ("blur", ctypes.c_int), ("depth_dtype", ctypes.c_int)]


class FrameInfo(ctypes.Structure):
    _fields_ = [("status", ctypes.c_uint32), ("layers", ctypes.c_int32), ("limit_step", ctypes.c_int32),
                ("fill_layer", ctypes.c_int32), ("strip", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("depth_max", ctypes.c_float), ("reserved2", ctypes.c_float),
                ("offset_range", ctypes.c_double * 2), ("holes", ctypes.c_uint64)]


_lib = None


def load():
    """Load libvrsbs.so; raises ImportError with the build command if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` in the repo root "
            "(there is no CPU fallback for this path).")
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.vrsbs_abi_version.restype = ci
    lib.vrsbs_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci]
    lib.vrsbs_destroy.argtypes = [vp]
    lib.vrsbs_last_error.argtypes = [vp]
    lib.vrsbs_last_error.restype = ctypes.c_char_p
    lib.vrsbs_reset.argtypes = [vp, ctypes.POINTER(Params)]
    lib.vrsbs_get_range_state.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ctypes.c_double)]
    lib.vrsbs_set_range_state.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_double)]
    lib.vrsbs_set_blur_weights.argtypes = [vp, ctypes.POINTER(cf), ci, ci]
    lib.vrsbs_depth_from_lowres.argtypes = [vp, vp, ci, ci, ci, cf, ci, ci, vp, vp]
    lib.vrsbs_depth_from_full.argtypes = [vp, vp, ci, ci, ci, vp, vp]
    lib.vrsbs_build_tables.argtypes = [vp, ci, ci, ci, vp]
    lib.vrsbs_warp_batch.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp]
    lib.vrsbs_process_batch.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
    lib.vrsbs_process_host.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, cf, vp]
    lib.vrsbs_submit_host.argtypes = [vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, ci, ci, ci, ci, ci, cf, vp, ctypes.c_uint,
                                      ctypes.POINTER(ctypes.c_uint64)]
    lib.vrsbs_collect.argtypes = [vp, ctypes.c_uint64]
    lib.vrsbs_host_depends_on.argtypes = [vp, vp]
    lib.vrsbs_get_frame_info.argtypes = [vp, ci, ctypes.POINTER(FrameInfo), vp]
    lib.vrsbs_get_tables.argtypes = [vp, ci, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32),
                                     ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint16), vp]
    lib.vrsbs_get_bounds.argtypes = [vp, ci, ci, ctypes.POINTER(cf), ctypes.POINTER(cf), vp]
    lib.vrsbs_get_hole_mask.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ctypes.c_uint32), vp]
    lib.vrsbs_get_stage_times.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    lib.vrsbs_launch_count.argtypes = [vp]
    lib.vrsbs_launch_count.restype = ctypes.c_uint64
    lib.vrsbs_set_option.argtypes = [vp, ctypes.c_char_p, ci]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ("vrsbs_abi_version",):
            fn.restype = ci
    if lib.vrsbs_abi_version() != ABI_VERSION:
        raise ImportError(f"libvrsbs.so ABI {lib.vrsbs_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


class Context:
    """Thin RAII wrapper around vrsbs_ctx*: turns negative return codes into VrsbsError."""

    def __init__(self, device, max_h, max_w, max_batch, max_layers=512):
        self.lib = load()
        self.handle = ctypes.c_void_p()
        rc = self.lib.vrsbs_create(ctypes.byref(self.handle), device, max_h, max_w, max_batch, max_layers)
        if rc != 0:
            msg = self.lib.vrsbs_last_error(None).decode()
            self.handle = None
            raise VrsbsError(rc, msg)
        self.device, self.max_h, self.max_w = device, max_h, max_w
        self.max_batch, self.max_layers = max_batch, max_layers

    def check(self, rc):
        if rc < 0:
            raise VrsbsError(rc, self.lib.vrsbs_last_error(self.handle).decode())
        return rc

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vrsbs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- state -------------------------------------------------------------------------------
    def reset(self, offset_fg, offset_bg, offset_step_size, blur=True, depth_dtype=DEPTH_F16):
        p = Params(offset_fg, offset_bg, offset_step_size, 1 if blur else 0, depth_dtype)
        self.check(self.lib.vrsbs_reset(self.handle, ctypes.byref(p)))

    def get_range_state(self):
        has = ctypes.c_int()
        rng = (ctypes.c_double * 2)()
        self.check(self.lib.vrsbs_get_range_state(self.handle, ctypes.byref(has), rng))
        return [rng[0], rng[1]] if has.value else None

    def set_range_state(self, rng):
        arr = (ctypes.c_double * 2)(*(rng if rng is not None else (0.0, 0.0)))
        self.check(self.lib.vrsbs_set_range_state(self.handle, 0 if rng is None else 1, arr))

    def set_blur_weights(self, weights):
        import numpy as np
        w = np.ascontiguousarray(weights, dtype=np.float32)
        ky, kx = w.shape
        self.check(self.lib.vrsbs_set_blur_weights(self.handle, w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), kx, ky))

    def set_option(self, name, value):
        self.check(self.lib.vrsbs_set_option(self.handle, name.encode(), int(value)))

    # --- stages (device pointers as ints) ------------------------------------------------------
    def depth_from_lowres(self, lo_ptr, B, h, w, scaler, H, W, out_ptr, stream=0):
        self.check(self.lib.vrsbs_depth_from_lowres(self.handle, lo_ptr, B, h, w, scaler, H, W, out_ptr, stream))

    def depth_from_full(self, raw_ptr, B, H, W, out_ptr, stream=0):
        self.check(self.lib.vrsbs_depth_from_full(self.handle, raw_ptr, B, H, W, out_ptr, stream))

    def build_tables(self, B, H, W, stream=0):
        self.check(self.lib.vrsbs_build_tables(self.handle, B, H, W, stream))

    def warp_batch(self, frames_ptr, depth_ptr, B, H, W, sbs_ptr, stream=0):
        self.check(self.lib.vrsbs_warp_batch(self.handle, frames_ptr, depth_ptr, B, H, W, sbs_ptr, stream))

    def process_batch(self, frames_ptr, raw_ptr, B, H, W, scratch_ptr, sbs_ptr, stream=0):
        self.check(self.lib.vrsbs_process_batch(self.handle, frames_ptr, raw_ptr, B, H, W, scratch_ptr, sbs_ptr, stream))

    def process_host(self, frames_ptr, depth_ptr, B, H, W, lh, lw, scaler, sbs_ptr):
        self.check(self.lib.vrsbs_process_host(self.handle, frames_ptr, depth_ptr, B, H, W, lh, lw, scaler, sbs_ptr))

    def submit_host(self, frames_ptr, row_pitch, frame_pitch, depth_ptr, B, H, W, lh, lw, scaler, sbs_ptr, flags=0):
        """Asynchronous vrsbs_process_host: returns a ticket for collect()."""
        t = ctypes.c_uint64()
        self.check(self.lib.vrsbs_submit_host(self.handle, frames_ptr, row_pitch, frame_pitch, depth_ptr, B, H, W, lh, lw,
                                              scaler, sbs_ptr, flags, ctypes.byref(t)))
        return int(t.value)

    def collect(self, ticket):
        self.check(self.lib.vrsbs_collect(self.handle, ticket))

    def host_depends_on(self, stream):
        self.check(self.lib.vrsbs_host_depends_on(self.handle, stream))

    # --- introspection -------------------------------------------------------------------------
    def frame_info(self, B, stream=0):
        arr = (FrameInfo * B)()
        self.check(self.lib.vrsbs_get_frame_info(self.handle, B, arr, stream))
        return list(arr)

    def tables(self, frame, stream=0):
        import numpy as np
        cap = self.max_layers + 1
        cut = np.empty(cap, dtype=np.float64)
        off = np.empty(cap, dtype=np.int32)
        lo = np.empty(cap, dtype=np.uint16)
        hi = np.empty(cap, dtype=np.uint16)
        L = self.check(self.lib.vrsbs_get_tables(
            self.handle, frame, cap, cut.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), lo.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)),
            hi.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), stream))
        return cut[:L + 1].copy(), off[:L].copy(), lo[:L].view(np.float16).copy(), hi[:L].view(np.float16).copy()

    def bounds(self, frame, stream=0):
        """(lo, hi) float32 arrays: the bounds the device compares against (exact for fp16 and fp32 depth)."""
        import numpy as np
        cap = self.max_layers
        lo = np.empty(cap, dtype=np.float32)
        hi = np.empty(cap, dtype=np.float32)
        L = self.check(self.lib.vrsbs_get_bounds(self.handle, frame, cap, lo.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                                 hi.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), stream))
        return lo[:L].copy(), hi[:L].copy()

    def hole_mask(self, B, H, W, stream=0):
        import numpy as np
        words = (W + 31) // 32
        m = np.empty((B, H, words), dtype=np.uint32)
        self.check(self.lib.vrsbs_get_hole_mask(self.handle, B, H, W, m.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), stream))
        bits = np.unpackbits(m.view(np.uint8).reshape(B, H, words * 4), axis=-1, bitorder="little")
        return bits[:, :, :W].astype(bool)

    STAGES = ("depth", "tables", "warp", "blur", "strip")

    def stage_times(self):
        """{stage: (total ms, launches)} since the previous call (needs set_option('stage_timing', 1))."""
        ms = (ctypes.c_double * 5)()
        n = (ctypes.c_uint64 * 5)()
        self.check(self.lib.vrsbs_get_stage_times(self.handle, ms, n))
        return {s: (ms[i], int(n[i])) for i, s in enumerate(self.STAGES)}

    def launch_count(self):
        return int(self.lib.vrsbs_launch_count(self.handle))
