"""Drive the UNMODIFIED reference warp (`PredictAndGenerate.py`) on the CPU or, as it ships, on CUDA.

TEST / BASELINE INFRASTRUCTURE ONLY.  This module exists to (a) pin the oracle restatements in this
directory against the reference itself, (b) generate the golden fixtures under `tests/golden/`
(see `tests/golden/make_golden.py`), (c) run the reference's own CUDA path on the GPU box for parity
tier T5-ii (`tests/test_gpu_reference_cuda.py`) and (d) time the reference for `bench.py`'s baseline legs.
The reference is loaded from `/root/reference` where that exists (the build container) and otherwise from
the byte-for-byte staged copy `oracle/_ref/` (`oracle/stage_ref.py`; git-ignored, shipped by gpurun).
Nothing in the product path imports it.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("VRSBS_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "PredictAndGenerate.py")):
            return cand
    return os.environ.get("VRSBS_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def reference_available():
    if os.environ.get("VRSBS_NO_REFERENCE"):               # tests: exercise the fallbacks for a box without any copy
        return False
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


def load_reference(device="cpu"):
    """Import the reference module (once) and point its `torch.device('cuda')` calls at `device`:
    "cpu" installs the forwarding proxy, "cuda" leaves the real torch in place (the code runs as shipped)."""
    global _PAG
    import torch
    if _PAG is None:
        if not reference_available():
            raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (run oracle/stage_ref.py in the build container)")
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        with contextlib.redirect_stdout(io.StringIO()):
            import PredictAndGenerate as PAG  # noqa: N814  (prints "Import done")
        _PAG = PAG
    _PAG.torch = _CpuTorch(torch) if device == "cpu" else torch
    return _PAG


def make_args(offset_fg=0.025, offset_bg=-0.01, offset_step_size=1):
    return argparse.Namespace(offset_fg=offset_fg, offset_bg=offset_bg,
                              offset_step_size=offset_step_size)


class ReferenceWarp:
    """One reference `SbsProcessor` (= one clip range: depth history + range EMA state)."""

    def __init__(self, offset_fg=0.025, offset_bg=-0.01, offset_step_size=1, device="cpu"):
        self.device = device
        PAG = load_reference(device)
        self.proc = PAG.SbsProcessor(None, 0, make_args(offset_fg, offset_bg, offset_step_size))
        self.q = queue.Queue()

    def left_side_sbs(self, img_u8, depth_tensor):
        """img_u8: numpy [H,W,3] uint8; depth_tensor: CPU torch tensor [H,W] (fp16 or fp32), raw - what the
        inference worker puts on the result queue (PredictAndGenerate.py:55-56)."""
        load_reference(self.device)            # several instances on different devices may alternate
        self.q.put(depth_tensor.clone())
        return self.proc.left_side_sbs(img_u8, None, self.q)

    def get_cutoff(self, depth_tensor):
        return self.proc.get_cutoff(depth_tensor)


def depth_model(encoder, device):
    """The reference's Depth-Anything-V2 module with random-initialised weights (no checkpoints exist offline),
    built like SupportFunction.load_model (:158-168) minus the `load_state_dict`."""
    import torch
    load_reference("cuda" if str(device).startswith("cuda") else "cpu")
    from depth_anything_v2.dpt import DepthAnythingV2
    cfg = {
        'vits': {'encoder': 'vits', 'features': 64, 'out_channels': [48, 96, 192, 384]},
        'vitb': {'encoder': 'vitb', 'features': 128, 'out_channels': [96, 192, 384, 768]},
        'vitl': {'encoder': 'vitl', 'features': 256, 'out_channels': [256, 512, 1024, 1024]},
    }[encoder]
    torch.manual_seed(0)
    return DepthAnythingV2(device=torch.device(device), **cfg).to(device).eval()
