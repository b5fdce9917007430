"""Drive the UNMODIFIED reference warp (`/root/reference/PredictAndGenerate.py`) on the CPU.

TEST INFRASTRUCTURE ONLY.  This module exists to (a) pin the oracle restatements in this
directory against the reference itself and (b) generate the golden fixtures under
`tests/golden/` (see `tests/golden/make_golden.py`).  It only works where `/root/reference`
exists (the build container); nothing in the product, the `-m gpu` tests, `smoke()` or
`bench.py` imports it.

The reference hard-codes `torch.device('cuda')` (PredictAndGenerate.py:133,148,158,161-163).
Instead of editing it, the module-level name `torch` inside the imported reference module is
replaced by a forwarding proxy whose `.device(...)` always answers the CPU device; every other
attribute is the real torch.  Depth maps are handed over through a plain `queue.Queue`
standing in for the inference worker (`left_side_sbs` only ever calls `result_queue.get()`).
"""
import argparse
import contextlib
import io
import os
import queue
import sys

REFERENCE_ROOT = os.environ.get("VRSBS_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "PredictAndGenerate.py"))


class _CpuTorch:
    """Forward everything to torch, except that every device is the CPU."""

    def __init__(self, real):
        object.__setattr__(self, "_real", real)

    def device(self, *a, **k):
        return self._real.device("cpu")

    def __getattr__(self, name):
        return getattr(self._real, name)


_PAG = None


def load_reference():
    global _PAG
    if _PAG is not None:
        return _PAG
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    import torch
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import PredictAndGenerate as PAG  # noqa: N814  (prints "Import done")
    PAG.torch = _CpuTorch(torch)
    _PAG = PAG
    return PAG


def make_args(offset_fg=0.025, offset_bg=-0.01, offset_step_size=1):
    return argparse.Namespace(offset_fg=offset_fg, offset_bg=offset_bg,
                              offset_step_size=offset_step_size)


class ReferenceWarp:
    """One reference `SbsProcessor` (= one clip range: depth history + range EMA state)."""

    def __init__(self, offset_fg=0.025, offset_bg=-0.01, offset_step_size=1):
        PAG = load_reference()
        self.proc = PAG.SbsProcessor(None, 0, make_args(offset_fg, offset_bg, offset_step_size))
        self.q = queue.Queue()

    def left_side_sbs(self, img_u8, depth_tensor):
        """img_u8: numpy [H,W,3] uint8; depth_tensor: CPU torch tensor [H,W] (fp16 or fp32), raw."""
        self.q.put(depth_tensor.clone())
        return self.proc.left_side_sbs(img_u8, None, self.q)

    def get_cutoff(self, depth_tensor):
        return self.proc.get_cutoff(depth_tensor)
